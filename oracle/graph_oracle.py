"""CPU restatement of `covid_abs.network.graph_abs.GraphSimulation` (TEST INFRASTRUCTURE ONLY).

The hourly routine model (graph_abs.py:190-300, network/agents.py:14-415, network/util.py:3-63) restated over
structure-of-arrays state with counter-based draws (oracle/philox.py), in the form a data-parallel step can
honour:

* every draw is a pure function of (seed; entity, iteration, stream, aux) -- `GraphDraws` below is the single
  definition of that map.  `oracle/graph_shim.py` feeds the *same* draws to the UNMODIFIED reference through a
  frame-inspecting `np.random` replacement, which is how this file is pinned: reference and restatement must then
  walk the same trajectory (positions, statuses, employment exactly; money to fixed-point rounding);
* the per-person loop of graph_abs.py:210-233 is order-free except for business stocks (a sale is capped by the
  stock left, network/agents.py:82-84, and every employee check-in adds one unit, 101-102).  Stock is therefore
  resolved exactly in the reference's person order with the max-plus scan  s -> max(s + a, b)
  (check-in: a=+1, b=-inf; request q: a=-q, b=0), which is associative and so parallelisable;
* the statistics that conditional lockdown (run_batch.py:22-26) and the critical-limit test
  (network/agents.py:389-390) read are the memo of the end of the previous executed hour -- what the reference
  holds when `batch_experiment` drives it (experiments.py:193-194);
* money that several entities add to concurrently (house / business / government / healthcare accounts) is held
  in fixed point (`money_bits` fractional bits, each flow rounded to nearest-even once), so sums do not depend
  on the order of the additions; the CUDA path does the same and must agree bit for bit.

Where the reference would raise (hiring with nobody unemployed, firing with no employee: `np.random.randint(0, 0)`
at network/agents.py:127-131, which aborts the experiment in batch_experiment) this restatement skips the action.

World state is a plain dict of numpy arrays (`World`), built either by converting a reference simulation
(`world_from_reference`) or by the product's host-side builder; the oracle never imports product code.
"""
import numpy as np

from oracle import philox

S, I, R, D = 0, 1, 2, 3
SEV_E, SEV_A, SEV_H, SEV_G = 0, 1, 2, 3
AGE_HOSP = [0.001, 0.003, 0.012, 0.032, 0.049, 0.102, 0.166, 0.243, 0.273]      # common.py:13
AGE_SEVERE = [0.05, 0.05, 0.05, 0.05, 0.063, 0.122, 0.274, 0.432, 0.709]        # common.py:14
AGE_DEATH = [0.002, 0.00006, 0.0003, 0.0008, 0.0015, 0.006, 0.022, 0.051, 0.093]  # common.py:15
LORENZ = [.04, .08, .13, .2, .55]                                               # common.py:25
BASIC_INCOME = np.array(LORENZ) / np.min(LORENZ)                                # common.py:27

GRAPH_STAT_KEYS = ("Q1", "Q2", "Q3", "Q4", "Q5", "Business", "Government", "Susceptible", "Infected",
                   "Recovered_Immune", "Death", "Asymptomatic", "Hospitalization", "Severe")  # graph_abs.py:305-322

# scenario modes = declarative form of the run_batch.py callbacks
MODE_NONE, MODE_LOCKDOWN, MODE_CONDITIONAL_LOCKDOWN, MODE_VERTICAL = 0, 1, 2, 3

# draw streams (shared contract with csrc/graph_kernels.cuh)
ST_MOVE, ST_UPDATE, ST_HOSP_MOVE, ST_DEATH_MOVE, ST_CONTACT, ST_GOV, ST_ACCOUNT, ST_SHOP = 0, 1, 2, 3, 4, 5, 6, 16
GOV_ID = 0xFFFFFFFF
BUSINESS_ID0 = 0xFFFF0000


class GraphDraws(object):
    """The draw map.  counter = (entity, iteration, stream, aux), key = seed."""

    def __init__(self, seed):
        self.k0, self.k1 = philox.split_seed(seed)

    def words(self, c0, it, stream, c3=0):
        return philox.philox4x32_10(c0, it, stream, c3, self.k0, self.k1)

    def normal2(self, pid, it, stream):
        """The `np.random.normal(0, sigma, 2)` of a move (network/agents.py:303,322,338,348): two float32 standard
        normals; the caller scales in binary64."""
        w = self.words(pid, it, stream)
        return philox.normal_pair(w[0], w[1])

    def update_uniforms(self, pid, it):
        """test_sub, death_test of Person.update (network/agents.py:380,403)."""
        w = self.words(pid, it, ST_UPDATE)
        return w[0], w[1]

    @staticmethod
    def randint(word, n):
        """np.random.randint(0, n) from 32 bits: floor(word * n / 2^32)."""
        return (np.asarray(word, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)

    def shop_qty(self, buyer, it, b):
        """np.random.randint(1, 10) of Business.supply (network/agents.py:82) for (buyer, business b)."""
        w = self.words(buyer, it, ST_SHOP + (b >> 2))
        return 1 + int(self.randint(np.asarray(w)[b & 3], 9))

    def contact(self, a, b, it):
        """low, up (np.random.randint(-1, 1), graph_abs.py:290-291) and the contagion word (294) of a pair."""
        lo, hi = (a, b) if a < b else (b, a)
        w = self.words(lo, it, ST_CONTACT, hi)
        return -1 + (int(w[0]) >> 31), -1 + (int(w[1]) >> 31), int(w[2])


def prob_bound_lt(p):
    """p > (w + .5) 2^-32  <=>  w < ceil(p 2^32 - .5)"""
    t = np.ceil(float(p) * 4294967296.0 - 0.5)
    return int(min(max(t, 0.0), 4294967296.0))


def prob_bound_le(r):
    """(w + .5) 2^-32 <= r  <=>  w <= floor(r 2^32 - .5)"""
    t = np.floor(float(r) * 4294967296.0 - 0.5)
    if not (t >= 0.0):
        return -1
    return int(min(t, 4294967295.0))


def distance_threshold(r):
    """Largest integer n with sqrt_f64(n) <= r, -1 if none (integer lattice positions, abs.py:9-10)."""
    r = float(r)
    if not (r >= 0.0):
        return -1
    n = int(np.floor(r * r))
    while np.sqrt(np.float64(n + 1)) <= r:
        n += 1
    while n >= 0 and np.sqrt(np.float64(n)) > r:
        n -= 1
    return n


def default_money_bits(total_wealth):
    tw = max(1.0, float(total_wealth))
    return int(max(8, min(24, np.floor(56 - np.log2(tw)))))


def trunc_add(base, z, sigma):
    """int(base + normal) of network/agents.py:304-305,323-324,339-340,349-350: binary64 add, truncation toward 0."""
    return np.trunc(np.asarray(base, dtype=np.float64) + np.asarray(z, dtype=np.float64) * float(sigma)).astype(np.int64)


class GraphOracle(object):
    """Synchronous restatement of GraphSimulation.execute / get_statistics over a `World` dict.

    world keys (numpy arrays unless noted):
      persons : px, py, status, severity, infected_time, age, stratum, pwealth (f64), house, employer (-1 = none),
                active (bool), isolated (bool), hire_seq (order inside the employer's list, network/agents.py:38)
      houses  : hx, hy, hstratum, hwealth (f64), hsize, hfixed (f64)
      business: bx, by, bstratum, bprice, bwealth (f64), bstocks, bnum_employees, bfixed (f64)
      scalars : gov = dict(x, y, wealth, price, fixed), health = dict(x, y, wealth, fixed, expenses)
      params  : dict with the GraphSimulation kwargs (graph_abs.py:13-32, abs.py:17-49) + mode, sleep, money_bits
    """

    def __init__(self, world, seed):
        w = world
        p = dict(w["params"])
        self.p = p
        self.draws = GraphDraws(seed)
        self.bits = int(p.get("money_bits", default_money_bits(p["total_wealth"])))
        self.scale = float(2 ** self.bits)
        fx = self.fx
        self.px = np.array(w["px"], dtype=np.int64)
        self.py = np.array(w["py"], dtype=np.int64)
        self.status = np.array(w["status"], dtype=np.int64)
        self.severity = np.array(w["severity"], dtype=np.int64)
        self.time = np.array(w["infected_time"], dtype=np.int64)
        self.age = np.array(w["age"], dtype=np.int64)
        self.stratum = np.array(w["stratum"], dtype=np.int64)
        self.pwealth = fx(np.array(w["pwealth"], dtype=np.float64))
        self.house = np.array(w["house"], dtype=np.int64)
        self.employer = np.array(w["employer"], dtype=np.int64)
        self.active = np.array(w["active"], dtype=bool)
        self.isolated = np.array(w["isolated"], dtype=bool)
        self.hire_seq = np.array(w["hire_seq"], dtype=np.int64)
        self.next_hire_seq = int(self.hire_seq.max(initial=-1)) + 1
        self.n = len(self.px)
        bi = BASIC_INCOME[np.clip(self.stratum, 0, 4)]
        self.incomes = np.where(self.active, bi * float(p["minimum_income"]), 0.0)   # graph_abs.py:156
        self.expenses = bi * float(p["minimum_expense"])                             # graph_abs.py:171
        self.hx = np.array(w["hx"], dtype=np.int64)
        self.hy = np.array(w["hy"], dtype=np.int64)
        self.hstratum = np.array(w["hstratum"], dtype=np.int64)
        self.hwealth = fx(np.array(w["hwealth"], dtype=np.float64))
        self.hsize = np.array(w["hsize"], dtype=np.int64)
        self.hfixed = fx(np.array(w["hfixed"], dtype=np.float64))
        self.hincomes = np.zeros(len(self.hx), dtype=np.int64)
        self.hexpenses = np.zeros(len(self.hx), dtype=np.int64)
        self.bx = np.array(w["bx"], dtype=np.int64)
        self.by = np.array(w["by"], dtype=np.int64)
        self.bstratum = np.array(w["bstratum"], dtype=np.int64)
        self.bprice = np.array(w["bprice"], dtype=np.float64)
        self.bwealth = fx(np.array(w["bwealth"], dtype=np.float64))
        self.bstocks = np.array(w["bstocks"], dtype=np.int64)
        self.bsales = np.zeros(len(self.bx), dtype=np.int64)
        self.bnum = np.array(w["bnum_employees"], dtype=np.int64)
        self.bfixed = fx(np.array(w["bfixed"], dtype=np.float64))
        self.bincomes = np.zeros(len(self.bx), dtype=np.int64)
        g, h = w["gov"], w["health"]
        self.gov = dict(x=int(g["x"]), y=int(g["y"]), wealth=int(fx(g["wealth"])), price=float(g["price"]),
                        fixed=int(fx(g["fixed"])), stocks=10, incomes=0, sales=0)
        self.health = dict(x=int(h["x"]), y=int(h["y"]), wealth=int(fx(h["wealth"])), fixed=int(fx(h["fixed"])),
                           expenses=int(fx(h.get("expenses", 0.0))))
        self.iteration = -1                                                          # graph_abs.py:25
        self.thr_hosp = [prob_bound_lt(v) for v in AGE_HOSP]
        self.thr_sev = [prob_bound_lt(v) for v in AGE_SEVERE]
        self.thr_death = [prob_bound_lt(v) for v in AGE_DEATH]
        self.stats = self.compute_statistics()
        self.last_contacts = 0

    # ------------------------------------------------------------------ money
    def fx(self, v):
        return np.rint(np.asarray(v, dtype=np.float64) * self.scale).astype(np.int64)

    def fl(self, v):
        return np.asarray(v, dtype=np.float64) / self.scale

    # ------------------------------------------------------------------ accounts
    def person_demand(self, i, value_fx):    # network/agents.py:279-284
        if self.house[i] >= 0:
            self.hwealth[self.house[i]] -= value_fx
            self.hexpenses[self.house[i]] += value_fx
        else:
            self.pwealth[i] -= value_fx

    def person_supply(self, i, value_fx):    # network/agents.py:286-291
        if self.house[i] >= 0:
            self.hwealth[self.house[i]] += value_fx
            self.hincomes[self.house[i]] += value_fx
        else:
            self.pwealth[i] += value_fx

    def house_checkin(self, i):              # network/agents.py:202-207: house.demand(agent.expenses / 720)
        v = int(self.fx(self.expenses[i] / 720))
        self.hwealth[self.house[i]] -= v
        self.hexpenses[self.house[i]] += v

    def fire(self, b, i):                    # network/agents.py:44-53
        self.employer[i] = -1
        self.hire_seq[i] = -1
        v = int(self.fx(self.incomes[i]))
        self.person_supply(i, v)
        self.bwealth[b] -= v
        self.bnum[b] -= 1
        self.bfixed[b] -= int(self.fx(self.expenses[i] / 720 * 24))

    def hire(self, b, i):                    # network/agents.py:36-42
        if self.status[i] != D and self.severity[i] == SEV_A:
            self.employer[i] = b
            self.hire_seq[i] = self.next_hire_seq
            self.next_hire_seq += 1
            self.bnum[b] += 1
            self.bfixed[b] += int(self.fx(self.expenses[i] / 720 * 24))

    # ------------------------------------------------------------------ one hour
    def clock(self, it):                     # network/util.py:25-63
        h = it % 24
        day = it // 24
        return dict(bed=0 <= h < 8, work=8 <= h < 16, free=17 <= h < 24, lunch=h == 12, new_day=h == 0,
                    work_day=(day % 7 + 1) not in (1, 7), new_month=(day % 30 + 1) == 1 and h == 0)

    def execute(self):
        """graph_abs.py:190-276.  Returns the statistics row batch_experiment would append."""
        p = self.p
        self.iteration += 1
        it = self.iteration
        c = self.clock(it)
        if p.get("sleep", False) and (not c["new_day"]) and c["bed"]:            # run_batch.py:8-13 (on_execute)
            return self.stats
        prev = self.stats
        amp = p["amplitudes"]
        mode = int(p.get("mode", MODE_NONE))
        lock_all = mode == MODE_LOCKDOWN or (mode == MODE_CONDITIONAL_LOCKDOWN and prev["Infected"] > .05)
        crit_active = (prev["Severe"] + prev["Hospitalization"]) >= float(p["critical_limit"])
        T_bus = distance_threshold(p["business_distance"])
        nb = len(self.bx)
        events = [[] for _ in range(nb)]      # per business, in person order: (+1 check-in) or (person, qty)

        for i in range(self.n):
            if self.status[i] == D:                                              # graph_abs.py:210
                continue
            # ---- on_person_move (run_batch.py:16-50) or the default routine (graph_abs.py:212-220)
            if lock_all or (mode == MODE_VERTICAL and not self.active[i]):
                if self.house[i] >= 0:
                    self.house_checkin(i)
            elif self.isolated[i]:
                self.move_to_home(i, it, ST_MOVE, amp)
            elif c["bed"]:
                self.move_to_home(i, it, ST_MOVE, amp)
            elif c["lunch"] or c["free"] or not c["work_day"]:
                self.move_freely(i, it, ST_MOVE, amp)
            elif c["work_day"] and c["work"]:
                b = self.move_to_work(i, it, amp)
                if b >= 0:
                    events[b].append((-1, 1))
            # ---- daily health update (graph_abs.py:227-228 -> network/agents.py:362-415)
            if c["new_day"]:
                self.update_person(i, it, crit_active, amp)
            # ---- purchases at every business in reach other than the employer (graph_abs.py:230-233)
            if self.severity[i] == SEV_A:
                dx = self.bx - self.px[i]
                dy = self.by - self.py[i]
                for b in np.nonzero(dx * dx + dy * dy <= T_bus)[0]:
                    if b != self.employer[i]:
                        events[b].append((i, self.draws.shop_qty(i, it, int(b))))

        # ---- stocks in the reference's person order (network/agents.py:82-93, 101-102)
        for b in range(nb):
            stock = int(self.bstocks[b])
            for who, q in events[b]:
                if who < 0:
                    stock += 1
                    continue
                taken = min(q, stock)
                stock -= taken
                value = int(self.fx(self.bprice[b] * float(self.stratum[who]) * taken))
                self.person_demand(who, value)
                self.bwealth[b] += value
                self.bincomes[b] += value
                self.bsales[b] += taken
            self.bstocks[b] = stock

        monthly = it > 1 and c["new_month"]
        # ---- businesses (graph_abs.py:235-240)
        for b in range(nb):
            if c["new_day"]:                                                     # network/agents.py:147-153
                self.gov["wealth"] += int(self.fx(self.fl(self.bfixed[b]) / 3))
                self.bwealth[b] -= self.bfixed[b]
            if monthly:
                self.business_accounting(b, it)
        # ---- houses (graph_abs.py:242-247)
        occupied = self.hsize > 0
        if c["new_day"]:                                                         # network/agents.py:240-245
            self.gov["wealth"] += int(np.sum(self.fx(self.fl(self.hfixed[occupied]) / 10)))
        if monthly:                                                              # network/agents.py:227-238
            tax = self.fx(self.gov["price"] * self.hsize[occupied] + self.fl(self.hexpenses[occupied]) / 10)
            self.gov["wealth"] += int(np.sum(tax))
            self.hwealth[occupied] -= tax
            self.hincomes[occupied] = 0
            self.hexpenses[occupied] = 0
        # ---- government and healthcare (graph_abs.py:249-254)
        if c["new_day"]:
            w = self.draws.words(GOV_ID, it, ST_GOV)                              # network/agents.py:163-166
            b = int(GraphDraws.randint(w[0], nb))
            q = min(1 + int(GraphDraws.randint(w[1], 9)), int(self.bstocks[b]))
            value = int(self.fx(self.bprice[b] * 4.0 * q))                        # government buys as stratum 4
            self.gov["wealth"] -= value
            self.bwealth[b] += value
            self.bincomes[b] += value
            self.bstocks[b] -= q
            self.bsales[b] += q
            self.gov["wealth"] += int(self.fx(self.fl(self.health["fixed"]) / 3))  # healthcare.update()
            self.health["wealth"] -= self.health["fixed"]
        if monthly:                                                              # network/agents.py:132-139
            labor = self.health["expenses"]
            self.health["wealth"] += labor
            self.gov["wealth"] -= labor
            alive_mild = (self.status != D) & (self.severity == SEV_A)
            for i in np.nonzero(alive_mild & (self.house < 0))[0]:                # homeless
                v = int(self.fx(self.expenses[i]))
                self.person_supply(i, v)
                self.gov["wealth"] -= v
            for i in np.nonzero(alive_mild & (self.employer < 0) & self.active)[0]:  # unemployed
                v = int(self.fx(self.expenses[i]))
                self.person_supply(i, v)
                self.gov["wealth"] -= v
            self.gov["incomes"] = 0
            self.gov["sales"] = 0

        self.contacts(it)
        self.stats = self.compute_statistics()
        return self.stats

    # ------------------------------------------------------------------ moves
    def move_freely(self, i, it, stream, amp):       # network/agents.py:332-342
        if self.severity[i] != SEV_A:
            return
        z0, z1 = self.draws.normal2(i, it, stream)
        a = float(amp[int(self.status[i])])
        self.px[i] = trunc_add(self.px[i], z0, a)
        self.py[i] = trunc_add(self.py[i], z1, a)

    def move_to_home(self, i, it, stream, amp):      # network/agents.py:312-330
        if self.severity[i] != SEV_A:
            return
        if self.house[i] >= 0:
            self.house_checkin(i)
            z0, z1 = self.draws.normal2(i, it, stream)
            self.px[i] = trunc_add(self.hx[self.house[i]], z0, 0.25)
            self.py[i] = trunc_add(self.hy[self.house[i]], z1, 0.25)
        else:
            self.pwealth[i] -= int(self.fx(self.incomes[i] / 720))
            self.move_freely(i, it, stream, amp)

    def move_to_work(self, i, it, amp):              # network/agents.py:293-310; returns the business checked in at
        if self.severity[i] != SEV_A:
            return -1
        if self.active[i]:
            b = int(self.employer[i])
            if b >= 0:
                z0, z1 = self.draws.normal2(i, it, ST_MOVE)
                self.px[i] = trunc_add(self.bx[b], z0, 0.25)
                self.py[i] = trunc_add(self.by[b], z1, 0.25)
                self.bwealth[b] -= int(self.fx(self.expenses[i] / 720))           # checkin, network/agents.py:101-103
                return b
            self.move_freely(i, it, ST_MOVE, amp)
        return -1

    # ------------------------------------------------------------------ health
    def update_person(self, i, it, crit_active, amp):   # network/agents.py:362-415
        if self.status[i] != I:
            return
        self.time[i] += 1
        age = int(self.age[i])
        ix = min(age // 10 - 1 if age > 10 else 0, 8)
        w_sub, w_death = self.draws.update_uniforms(i, it)
        if self.severity[i] == SEV_A:
            if int(w_sub) < self.thr_hosp[ix]:
                self.severity[i] = SEV_H
                z0, z1 = self.draws.normal2(i, it, ST_HOSP_MOVE)                  # move_to(healthcare), 344-354
                self.px[i] = trunc_add(self.health["x"], z0, 0.25)
                self.py[i] = trunc_add(self.health["y"], z1, 0.25)
                self.health["expenses"] += int(self.fx(self.expenses[i]))         # checkin, network/agents.py:104-105
        elif self.severity[i] == SEV_H:
            if int(w_sub) < self.thr_sev[ix]:
                self.severity[i] = SEV_G
                if crit_active:                                                   # 388-401
                    self.status[i] = D
                    self.severity[i] = SEV_A
                    if self.house[i] >= 0:                                        # remove_mate, 196-200
                        h = self.house[i]
                        self.hwealth[h] -= int(self.fx(self.fl(self.pwealth[i]) / 2))
                        self.hsize[h] -= 1
                        self.hfixed[h] -= int(self.fx(self.expenses[i] / 720 * 24))
                    else:
                        self.gov["wealth"] -= int(self.fx(self.expenses[i]))
                    if self.employer[i] >= 0:
                        self.fire(int(self.employer[i]), i)
                    else:
                        self.gov["wealth"] -= int(self.fx(self.expenses[i]))
        if int(w_death) < self.thr_death[ix]:                                      # 403-408
            self.status[i] = D
            self.severity[i] = SEV_A
            self.move_to_home(i, it, ST_DEATH_MOVE, amp)
            return
        if self.time[i] > int(self.p["recovering_time"]):                          # 410-413
            self.time[i] = 0
            self.status[i] = R
            self.severity[i] = SEV_A

    # ------------------------------------------------------------------ monthly accounting of a business
    def business_accounting(self, b, it):            # network/agents.py:115-145
        members = np.nonzero(self.employer == b)[0]
        labor = 0
        for i in members:                            # demand(), 55-76
            if self.status[i] != D and self.severity[i] == SEV_A:
                v = int(self.fx(self.incomes[i]))
                self.person_supply(i, v)
                self.bwealth[b] -= v
                labor += v
        tax = int(self.fx(self.gov["price"] * float(self.bnum[b]) + self.fl(self.bincomes[b]) / 20))   # taxes(), 108-113
        self.gov["wealth"] += tax
        self.bwealth[b] -= tax
        w = self.draws.words(BUSINESS_ID0 + b, it, ST_ACCOUNT)
        if 2 * (self.fl(labor) + self.fl(tax)) < self.fl(self.bincomes[b]):
            unemployed = np.nonzero((self.employer < 0) & self.active & (self.status != D) &
                                    (self.severity == SEV_A))[0]                  # graph_abs.py:43-45, population order
            if len(unemployed) > 0:
                self.hire(b, int(unemployed[int(GraphDraws.randint(w[0], len(unemployed)))]))
        elif (self.fl(labor) + self.fl(tax)) > self.fl(self.bincomes[b]):
            if self.bnum[b] > 0 and len(members) > 0:
                order = members[np.argsort(self.hire_seq[members], kind="stable")]   # list order = hire order
                ix = int(GraphDraws.randint(w[0], int(self.bnum[b])))
                if ix < len(order):
                    self.fire(b, int(order[ix]))
        self.bincomes[b] = 0
        self.bsales[b] = 0

    # ------------------------------------------------------------------ contacts
    def contacts(self, it):                          # graph_abs.py:256-300
        p = self.p
        T = distance_threshold(p["contagion_distance"])
        self.last_contacts = 0
        if T < 0:
            return
        inf = np.nonzero((self.status == I))[0]
        sus = np.nonzero((self.status == S))[0]
        if len(inf) == 0 or len(sus) == 0:
            return
        thr_rate = prob_bound_le(p["contagion_rate"])
        inc, con = int(p["incubation_time"]), int(p["contagion_time"])
        newly = []
        for j in inf:
            d2 = (self.px[sus] - self.px[j]) ** 2 + (self.py[sus] - self.py[j]) ** 2
            for s in sus[d2 <= T]:
                low, up, word = self.draws.contact(int(s), int(j), it)
                self.last_contacts += 1
                if inc + low <= self.time[j] <= con + up and word <= thr_rate:   # graph_abs.py:292-296
                    newly.append(s)
        # contact() only rewrites `status` (the reference assigns a misspelt attribute, graph_abs.py:298)
        self.status[np.array(newly, dtype=np.int64)] = I

    # ------------------------------------------------------------------ statistics
    def compute_statistics(self):                    # graph_abs.py:302-324
        tw = float(self.p["total_wealth"])
        n = float(self.p["population_size"])
        out = {}
        for q in range(5):
            out["Q%d" % (q + 1)] = float(self.fl(int(np.sum(self.hwealth[self.hstratum == q])))) / tw
        out["Business"] = float(self.fl(int(np.sum(self.bwealth)))) / tw
        out["Government"] = float(self.fl(self.gov["wealth"])) / tw
        for k, name in enumerate(("Susceptible", "Infected", "Recovered_Immune", "Death")):
            out[name] = float(np.sum(self.status == k)) / n
        alive = self.status != D
        for k, name in ((SEV_A, "Asymptomatic"), (SEV_H, "Hospitalization"), (SEV_G, "Severe")):
            out[name] = float(np.sum((self.severity == k) & alive)) / n
        return out


# ---------------------------------------------------------------------------------------------------------
def world_from_reference(sim, mode=MODE_NONE, sleep=False, isolated_ids=(), money_bits=None):
    """Convert an initialised reference GraphSimulation (unmodified covid_abs objects) to a `World` dict."""
    pop = sim.population
    hindex = {id(h): k for k, h in enumerate(sim.houses)}
    bindex = {id(b): k for k, b in enumerate(sim.business)}
    stc = {"s": S, "i": I, "c": R, "m": D}
    sevc = {"e": SEV_E, "a": SEV_A, "h": SEV_H, "g": SEV_G}
    hire_seq = np.full(len(pop), -1, dtype=np.int64)
    seq = 0
    pindex = {id(a): k for k, a in enumerate(pop)}
    # hire order: the reference appends to business.employees while walking quintiles (graph_abs.py:149-169);
    # a global sequence that is increasing inside every employees list reproduces each list's order
    for b in sim.business:
        for a in b.employees:
            hire_seq[pindex[id(a)]] = seq
            seq += 1
    iso = set(isolated_ids)
    f = lambda v: float(np.asarray(v, dtype=np.float64).ravel()[0])  # noqa: E731
    world = dict(
        px=[int(a.x) for a in pop], py=[int(a.y) for a in pop],
        status=[stc[a.status.value] for a in pop], severity=[sevc[a.infected_status.value] for a in pop],
        infected_time=[int(a.infected_time) for a in pop], age=[int(a.age) for a in pop],
        stratum=[int(a.social_stratum) for a in pop], pwealth=[f(a.wealth) for a in pop],
        house=[hindex[id(a.house)] if a.house is not None else -1 for a in pop],
        employer=[bindex[id(a.employer)] if a.employer is not None else -1 for a in pop],
        active=[a.economical_status.value == 1 for a in pop], isolated=[a.id in iso for a in pop], hire_seq=hire_seq,
        hx=[int(h.x) for h in sim.houses], hy=[int(h.y) for h in sim.houses],
        hstratum=[int(h.social_stratum) for h in sim.houses], hwealth=[f(h.wealth) for h in sim.houses],
        hsize=[int(h.size) for h in sim.houses], hfixed=[f(h.fixed_expenses) for h in sim.houses],
        bx=[int(b.x) for b in sim.business], by=[int(b.y) for b in sim.business],
        bstratum=[int(b.social_stratum) for b in sim.business], bprice=[f(b.price) for b in sim.business],
        bwealth=[f(b.wealth) for b in sim.business], bstocks=[int(b.stocks) for b in sim.business],
        bnum_employees=[int(b.num_employees) for b in sim.business], bfixed=[f(b.fixed_expenses) for b in sim.business],
        gov=dict(x=int(sim.government.x), y=int(sim.government.y), wealth=f(sim.government.wealth),
                 price=f(sim.government.price), fixed=f(sim.government.fixed_expenses)),
        health=dict(x=int(sim.healthcare.x), y=int(sim.healthcare.y), wealth=f(sim.healthcare.wealth),
                    fixed=f(sim.healthcare.fixed_expenses), expenses=f(sim.healthcare.expenses)),
    )
    amp = np.zeros(4)
    for k, v in sim.amplitudes.items():
        amp[stc[k.value]] = float(v)
    world["params"] = dict(
        population_size=int(sim.population_size), total_wealth=float(sim.total_wealth),
        contagion_distance=float(sim.contagion_distance), contagion_rate=float(sim.contagion_rate),
        critical_limit=float(sim.critical_limit), amplitudes=amp.tolist(),
        minimum_income=float(sim.minimum_income), minimum_expense=float(sim.minimum_expense),
        business_distance=float(sim.business_distance), incubation_time=int(sim.incubation_time),
        contagion_time=int(sim.contagion_time), recovering_time=int(sim.recovering_time),
        mode=int(mode), sleep=bool(sleep),
        money_bits=int(money_bits if money_bits is not None else default_money_bits(sim.total_wealth)))
    return world
