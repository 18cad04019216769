"""Synchronous-schedule CPU restatement of the `Simulation` step (TEST INFRASTRUCTURE ONLY).

Same per-agent rules as `seq_oracle.SeqSimulation` (abs.py:141-230), restated
for the schedule a data-parallel step can honour:

* every agent does move -> update with its own counter-based draws
  (oracle/philox.py) instead of consuming a shared sequential stream
  (abs.py:240-242);
* the contact set is enumerated on the post-move positions (abs.py:249-259) and
  a Susceptible agent becomes Infected iff at least one of its Infected
  contacts passes `draw(pair, step) <= contagion_rate` (abs.py:149-153), where
  "Infected" means infected *before* this step's contact phase -- the reference
  applies contacts one after the other so an agent infected early in the pair
  list can infect others in the same step; that in-step chain is
  the one documented divergence of the synchronous schedule;
* the critical-limit test (abs.py:214-217) and the declarative simulation
  triggers (abs.py:267-271) read the statistics of the end of the previous step,
  which is what the reference does when driven by `batch_experiment`
  (experiments.py:193-194: the memo filled by `get_statistics` after step t-1 is
  still valid during step t, abs.py:273,298).

Q1..Q5 are accumulated in fixed point (wealth * 2**scale_bits rounded to
nearest-even, summed as int64) so that the sum does not depend on the order of
the agents -- the CUDA path does the same and must agree bit for bit.

Contact enumeration here uses a k-d tree (scipy) or brute force, never a cell
list: it is meant to be algorithmically independent of the CUDA path.
"""
import numpy as np

from oracle import philox
from oracle.seq_oracle import (S, I, R, D, SEV_ASYM, SEV_HOSP, SEV_SEVERE, STATUS_NAMES, SEVERITY_NAMES,
                               AGE_HOSP, AGE_SEVERE, AGE_DEATH, BASIC_INCOME, LORENZ, STAT_KEYS)

OPS = {'>=': np.greater_equal, '<=': np.less_equal, '>': np.greater, '<': np.less}


def distance_threshold(r):
    """largest integer n with sqrt_f64(n) <= r (integer lattice positions); -1 if none."""
    r = float(r)
    if not (r >= 0.0):
        return -1
    n = int(np.floor(r * r))
    while np.sqrt(np.float64(n + 1)) <= r:
        n += 1
    while n >= 0 and np.sqrt(np.float64(n)) > r:
        n -= 1
    return n


def contact_pairs(x, y, r, brute_limit=3000):
    """All (i, j), i < j, with sqrt(dx^2 + dy^2) <= r (abs.py:253-259), sorted lexicographically."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    n = len(x)
    T = distance_threshold(r)
    if T < 0 or n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    if n <= brute_limit:
        dx = x[:, None] - x[None, :]
        dy = y[:, None] - y[None, :]
        ii, jj = np.nonzero(np.triu(dx * dx + dy * dy <= T, k=1))
        return np.stack([ii, jj], axis=1).astype(np.int64)
    from scipy.spatial import cKDTree
    pts = np.stack([x, y], axis=1).astype(np.float64)
    tree = cKDTree(pts)
    cand = tree.query_pairs(np.sqrt(T) + 0.25, output_type='ndarray').astype(np.int64)
    if len(cand) == 0:
        return np.zeros((0, 2), dtype=np.int64)
    dx = x[cand[:, 0]] - x[cand[:, 1]]
    dy = y[cand[:, 0]] - y[cand[:, 1]]
    cand = cand[dx * dx + dy * dy <= T]
    lo = np.minimum(cand[:, 0], cand[:, 1])
    hi = np.maximum(cand[:, 0], cand[:, 1])
    order = np.lexsort((hi, lo))
    return np.stack([lo[order], hi[order]], axis=1)


class Trigger(object):
    """Declarative form of an abs.py:74-84 simulation trigger: `if stats[metric] <op> threshold: attribute = value`."""

    def __init__(self, metric, op, threshold, attribute, value):
        assert metric in STAT_KEYS and op in OPS
        assert attribute in ('amplitudes', 'contagion_rate', 'contagion_distance', 'critical_limit')
        self.metric, self.op, self.threshold, self.attribute, self.value = metric, op, float(threshold), attribute, value


def contact_pairs_float(x, y, r, brute_limit=3000):
    """Off-lattice positions: all (i, j), i < j, with the reference's own binary64 predicate
    sqrt((xi - xj)**2 + (yi - yj)**2) <= r (abs.py:9-10,258), every operation rounded once."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    r = np.float64(r)
    if not (r >= 0.0) or n < 2:
        return np.zeros((0, 2), dtype=np.int64)
    if n <= brute_limit:
        dx = x[:, None] - x[None, :]
        dy = y[:, None] - y[None, :]
        ii, jj = np.nonzero(np.triu(np.sqrt(dx * dx + dy * dy) <= r, k=1))
        return np.stack([ii, jj], axis=1).astype(np.int64)
    from scipy.spatial import cKDTree
    tree = cKDTree(np.stack([x, y], axis=1))
    cand = tree.query_pairs(float(r) * (1 + 1e-9) + 1e-9, output_type='ndarray').astype(np.int64)
    if len(cand) == 0:
        return np.zeros((0, 2), dtype=np.int64)
    lo = np.minimum(cand[:, 0], cand[:, 1])
    hi = np.maximum(cand[:, 0], cand[:, 1])
    dx = x[lo] - x[hi]
    dy = y[lo] - y[hi]
    keep = np.sqrt(dx * dx + dy * dy) <= r
    lo, hi = lo[keep], hi[keep]
    order = np.lexsort((hi, lo))
    return np.stack([lo[order], hi[order]], axis=1)


class _AgentView(object):
    """What a population-trigger condition sees (abs.py:86-95)."""

    def __init__(self, status, infected_status, age, social_stratum):
        self.status, self.infected_status, self.age, self.social_stratum = status, infected_status, age, social_stratum


class SyncSimulation(object):
    def __init__(self, seed=0, wealth_scale_bits=None, **kwargs):
        self.seed = int(seed)
        self.population_size = kwargs.get("population_size", 20)
        self.length = kwargs.get("length", 10)
        self.height = kwargs.get("height", 10)
        self.contagion_distance = kwargs.get("contagion_distance", 1.)
        self.contagion_rate = kwargs.get("contagion_rate", 0.9)
        self.critical_limit = kwargs.get("critical_limit", 0.6)
        self.amplitudes = dict(kwargs.get("amplitudes", {S: 5, R: 5, I: 5}))
        self.minimum_expense = kwargs.get("minimum_expense", 1.0)
        self.total_wealth = kwargs.get("total_wealth", 10 ** 4)
        self.triggers = list(kwargs.get("triggers", []))
        # "synchronous": only agents Infected before the contact phase transmit (one data-parallel pass);
        # "sequential": the reference's own order -- pairs walked lexicographically, statuses mutating on the way,
        # so an agent infected by an earlier pair transmits along later ones (abs.py:261-265)
        self.schedule = kwargs.get("schedule", "synchronous")
        if wealth_scale_bits is None:
            wealth_scale_bits = default_wealth_scale_bits(self.total_wealth)
        self.wealth_scale_bits = int(wealth_scale_bits)
        self.step_index = 0
        self.prev_stats = None
        self.last_pairs = None
        # population triggers (abs.py:86-95), declarative: dicts
        #   {'kind': 'move', 'when': callable(view) -> bool, 'target': 'stay' | (x, y), 'sigma': s}
        #   {'kind': 'attr', 'when': ..., 'attribute': name, 'p': p, 'q': q}   (discrete: := int(q); wealth: p * w + q)
        # `when` receives status / infected_status codes of this module, age, social_stratum
        self.rules = list(kwargs.get("rules", []))
        self.float_positions = bool(kwargs.get("float_positions", False)) or len(self.rules) > 0

    def _matches(self, when, st, sev, k):
        return bool(when(_AgentView(int(st[k]), int(sev[k]), int(self.age[k]), int(self.stratum[k]))))

    def set_population(self, x, y, status, age, stratum, wealth, severity=None, infected_time=None):
        n = len(x)
        self.population_size = n
        ptype = np.float64 if self.float_positions else np.int64
        self.x = np.array(x, dtype=ptype)
        self.y = np.array(y, dtype=ptype)
        self.status = np.array(status, dtype=np.int64)
        self.age = np.array(age, dtype=np.int64)
        self.stratum = np.array(stratum, dtype=np.int64)
        self.wealth = np.array(wealth, dtype=np.float64)
        self.severity = (np.full(n, SEV_ASYM, dtype=np.int64) if severity is None
                         else np.array(severity, dtype=np.int64))
        self.infected_time = (np.zeros(n, dtype=np.int64) if infected_time is None
                              else np.array(infected_time, dtype=np.int64))
        self.prev_stats = self.get_statistics()

    # ---------------------------------------------------------------- step
    def amplitude_table(self):
        return np.array([self.amplitudes.get(S, 0), self.amplitudes.get(I, 0), self.amplitudes.get(R, 0), 0],
                        dtype=np.float32)

    def execute(self, need_pairs=True):
        n = self.population_size
        ids = np.arange(n, dtype=np.uint64)
        step = self.step_index
        w0, w1 = philox.agent_draws(self.seed, ids, step, philox.STREAM_MOVE)
        w2, w3 = philox.agent_draws(self.seed, ids, step, philox.STREAM_ECONOMY)
        v0, _ = philox.agent_draws(self.seed, ids, step, philox.STREAM_DEATH)
        st, sev = self.status, self.severity
        dead = st == D
        inf = st == I
        # ---- move (abs.py:156-189)
        mobile = ~(dead | (inf & ((sev == SEV_HOSP) | (sev == SEV_SEVERE))))
        z0, z1 = philox.normal_pair(w0, w1)
        amp = self.amplitude_table()[st]
        ix = np.trunc(z0 * amp).astype(np.int64)          # float32 product, int() truncates toward zero
        iy = np.trunc(z1 * amp).astype(np.int64)
        ix = np.where(mobile, ix, 0)
        iy = np.where(mobile, iy, 0)
        # population triggers with attribute 'move' (abs.py:169-172): the first whose condition holds replaces the walk
        moved_by_rule = np.zeros(n, dtype=bool)
        move_rules = [r for r in self.rules if r['kind'] == 'move']
        if move_rules:
            rx, ry = self.x.copy(), self.y.copy()
            for k in np.nonzero(mobile)[0]:
                for r in move_rules:
                    if self._matches(r['when'], st, sev, k):
                        tx, ty = (self.x[k], self.y[k]) if r['target'] == 'stay' else r['target']
                        rx[k] = np.float64(tx) + np.float64(z0[k]) * np.float64(r['sigma'])
                        ry[k] = np.float64(ty) + np.float64(z1[k]) * np.float64(r['sigma'])
                        moved_by_rule[k] = True
                        break
        if self.float_positions:                       # abs.py:177-185 on binary64 positions
            fx, fy = ix.astype(np.float64), iy.astype(np.float64)
            nx = self.x + fx
            ny = self.y + fy
            bx = (nx <= 0) | (nx >= np.float64(self.length))
            by = (ny <= 0) | (ny >= np.float64(self.height))
            wx = np.where(bx, self.x - fx, nx)
            wy = np.where(by, self.y - fy, ny)
            if move_rules:
                wx = np.where(moved_by_rule, rx, wx)
                wy = np.where(moved_by_rule, ry, wy)
            self.x, self.y = wx, wy
        else:
            nx = self.x + ix
            ny = self.y + iy
            bx = (nx <= 0) | (nx >= self.length)
            by = (ny <= 0) | (ny >= self.height)
            self.x = np.where(bx, self.x - ix, nx)
            self.y = np.where(by, self.y - iy, ny)
        dist = np.sqrt((ix * ix + iy * iy).astype(np.float64))
        inc = ((dist * philox.u01(w2)) * np.float64(self.minimum_expense)) * BASIC_INCOME[self.stratum]
        self.wealth = np.where(mobile & ~moved_by_rule, self.wealth + inc, self.wealth)
        # ---- update (abs.py:191-230)
        idx = np.where(self.age > 10, self.age // 10 - 1, 0)
        idx = np.minimum(idx, 8)
        hosp = np.array(AGE_HOSP)[idx]
        severe = np.array(AGE_SEVERE)[idx]
        death = np.array(AGE_DEATH)[idx]
        u_sub = philox.u01(w3)
        u_death = philox.u01(v0)
        prev = self.prev_stats
        crit = (prev['Severe'] + prev['Hospitalization']) >= self.critical_limit
        time = np.where(inf, self.infected_time + 1, self.infected_time)
        to_hosp = inf & (sev == SEV_ASYM) & (hosp > u_sub)
        to_severe = inf & (sev == SEV_HOSP) & (severe > u_sub)
        sev = np.where(to_hosp, SEV_HOSP, sev)
        sev = np.where(to_severe, SEV_SEVERE, sev)
        crit_death = to_severe & crit
        st = np.where(crit_death, D, st)
        sev = np.where(crit_death, SEV_ASYM, sev)
        die = inf & (death > u_death)
        st = np.where(die, D, st)
        sev = np.where(die, SEV_ASYM, sev)
        recover = inf & ~die & (time > 20)                 # also "resurrects" a critical-limit death (A.7-3)
        time = np.where(recover, 0, time)
        st = np.where(recover, R, st)
        sev = np.where(recover, SEV_ASYM, sev)
        pays = ~dead & ~die
        expense = np.float64(self.minimum_expense) * BASIC_INCOME[self.stratum]
        self.wealth = np.where(pays, self.wealth - expense, self.wealth)
        self.infected_time = time
        self.severity = sev
        # ---- population triggers on other attributes (abs.py:244-247), in order, on the updated agent
        attr_rules = [r for r in self.rules if r['kind'] == 'attr']
        if attr_rules:
            st = st.copy()
            sev = sev.copy()
            for k in range(n):                            # every agent, the dead included (the loop is in execute())
                for r in attr_rules:
                    if not self._matches(r['when'], st, sev, k):
                        continue
                    a, v = r['attribute'], int(r['q'])
                    if a == 'status':
                        st[k] = v & 3
                    elif a == 'infected_status':
                        sev[k] = v & 3
                    elif a == 'infected_time':
                        self.infected_time[k] = v & 0xFF
                    elif a == 'age':
                        self.age[k] = v & 0xFF
                    elif a == 'social_stratum':
                        self.stratum[k] = min(v, 4)
                    else:
                        self.wealth[k] = np.float64(r['p']) * self.wealth[k] + np.float64(r['q'])
            self.severity = sev
        # ---- contacts (abs.py:249-265), synchronous
        pairs = (contact_pairs_float(self.x, self.y, self.contagion_distance) if self.float_positions
                 else contact_pairs(self.x, self.y, self.contagion_distance))
        self.last_pairs = pairs
        if len(pairs):
            a, b = pairs[:, 0], pairs[:, 1]
            si = (st[a] == S) & (st[b] == I)
            is_ = (st[a] == I) & (st[b] == S)
            enc = si | is_
            a, b, si = a[enc], b[enc], si[enc]
            c0, _, _, _ = philox.pair_draws(self.seed, a.astype(np.uint64), b.astype(np.uint64), step)
            ok = philox.u01(c0) <= np.float64(self.contagion_rate)
            if self.schedule == "sequential":
                c_all, _, _, _ = philox.pair_draws(self.seed, pairs[:, 0].astype(np.uint64),
                                                   pairs[:, 1].astype(np.uint64), step)
                ok_all = philox.u01(c_all) <= np.float64(self.contagion_rate)
                st = st.copy()
                for k in range(len(pairs)):                      # lexicographic (i, j), abs.py:253-265
                    i, j = pairs[k]
                    if st[i] == S and st[j] == I:
                        if ok_all[k]:
                            st[i] = I
                    elif st[j] == S and st[i] == I:
                        if ok_all[k]:
                            st[j] = I
            else:
                target = np.where(si, a, b)[ok]
                st = st.copy()
                st[target] = I
        self.status = st
        # ---- statistics + declarative triggers
        stats = self.get_statistics()
        for trig in self.triggers:
            if OPS[trig.op](prev[trig.metric], trig.threshold):
                if trig.attribute == 'amplitudes':
                    self.amplitudes = dict(trig.value)
                else:
                    setattr(self, trig.attribute, trig.value)
        self.prev_stats = stats
        self.step_index += 1
        return stats

    def get_statistics(self):
        n = self.population_size
        stats = {}
        for code, name in enumerate(STATUS_NAMES):
            stats[name] = np.float64(np.sum(self.status == code)) / n
        alive = self.status != D
        for code in (SEV_ASYM, SEV_HOSP, SEV_SEVERE):
            stats[SEVERITY_NAMES[code]] = np.float64(np.sum((self.severity == code) & alive)) / n
        scale = 2.0 ** self.wealth_scale_bits
        fixed = np.rint(self.wealth * scale).astype(np.int64)
        for q in range(5):
            sel = (self.stratum == q) & (self.age >= 18) & alive
            stats['Q{}'.format(q + 1)] = np.float64(int(np.sum(fixed[sel]))) / scale
        return stats


def default_wealth_scale_bits(total_wealth):
    """Largest s <= 40 with 1024 * total_wealth * 2**s < 2**62 (headroom for wealth growth)."""
    tw = max(1.0, float(total_wealth))
    s = int(np.floor(62 - np.log2(1024.0 * tw)))
    return max(0, min(40, s))


def synthetic_population(n, length, height, seed, total_wealth=None, infected=0.01, immune=0.01):
    """Config-3 style placement: x,y ~ U{0..L}, age=int(Beta(2,5)*100), stratum ~ U{0..4},
    wealth per abs.py:134-139, first int(n*infected) agents Infected, next int(n*immune) immune (abs.py:121-131)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, int(length) + 1, size=n)
    y = rng.integers(0, int(height) + 1, size=n)
    age = (rng.beta(2, 5, size=n) * 100).astype(np.int64)
    stratum = rng.integers(0, 5, size=n)
    status = np.full(n, S, dtype=np.int64)
    ni, nm = int(n * infected), int(n * immune)
    status[:ni] = I
    status[ni:ni + nm] = R
    if total_wealth is None:
        total_wealth = 1e4 * n / 20
    wealth = np.zeros(n)
    for q in range(5):
        sel = (stratum == q) & (age >= 18)
        wealth[sel] = LORENZ[q] * total_wealth / max(1.0, float(np.sum(sel)))
    return dict(x=x, y=y, status=status, age=age, stratum=stratum, wealth=wealth), total_wealth
