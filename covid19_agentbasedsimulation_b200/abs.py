"""Host-side mirror of covid_abs/abs.py: the `Simulation` engine with its per-iteration step on a B200.

Same constructor kwargs, attributes and methods as the reference class (abs.py:13-322); the population
lives on the GPU as structure of arrays and `execute()` runs the hand-written sm_100a kernels of
`csrc/` through the C ABI of `include/covidabs.h`.  There is no CPU fallback.

Differences a user can observe (all documented in DESIGN.md):
  * draws come from Philox4x32-10 keyed (seed, agent, step) instead of numpy's sequential MT19937
    stream; `seed=None` (default) takes the seed from `np.random`, so `np.random.seed(s)` still makes a
    run reproducible;
  * contacts are applied synchronously (an agent infected in a step starts transmitting in the next);
  * the critical-limit test (abs.py:214-217) reads the statistics of the end of the previous step,
    which is what the reference's memo holds when `batch_experiment` drives it;
  * `triggers_population` (abs.py:86-95) run on the device as declarative rules (`rules.MoveRule`, `AttributeRule`);
    the reference's own dict-of-lambdas entries are recognised by probing when they have the shipped shape (condition on
    status / infected_status / age / social_stratum; move action = point + np.random.normal(0, sigma); attribute
    action = constant, or scale * wealth + shift), anything else raises NotImplementedError.  With a population trigger
    the positions are binary64, as in the reference once a move action has assigned a float (`positions="float64"`);
  * `get_population()` / `sim.population` return `Agent` objects materialised from a device download: they are
    detached copies, so mutating `sim.population[i]` in place does not reach the device -- pass the edited list
    back through `set_population()` (or use `set_population_arrays`).
"""
import numpy as np

from covid19_agentbasedsimulation_b200 import _native
from covid19_agentbasedsimulation_b200.agents import (Status, InfectionSeverity, Agent, STATUS_LIST, STATUS_CODE,
                                                      SEVERITY_LIST, SEVERITY_CODE)
from covid19_agentbasedsimulation_b200.common import *  # noqa: F401,F403  (the reference does the same, abs.py:6)
from covid19_agentbasedsimulation_b200.common import (age_hospitalization_probs, age_severe_probs, age_death_probs,
                                                      lorenz_curve, basic_income)
from covid19_agentbasedsimulation_b200.rules import MoveRule, AttributeRule, When, lower_trigger  # noqa: F401

STAT_KEYS = ("Susceptible", "Infected", "Recovered_Immune", "Death", "Asymptomatic", "Hospitalization", "Severe",
             "Q1", "Q2", "Q3", "Q4", "Q5")
_OPS = {'>=': 0, '<=': 1, '>': 2, '<': 3}
_ATTRS = {'amplitudes': 0, 'contagion_rate': 1, 'contagion_distance': 2, 'critical_limit': 3}


def distance(a, b):
    """abs.py:9-10."""
    return np.sqrt((a.x - b.x) ** 2 + (a.y - b.y) ** 2)


class Trigger(object):
    """Declarative simulation trigger, evaluated on the device at the end of every step.

    Equivalent of `{'condition': lambda s: s.get_statistics()[metric] <op> threshold, 'attribute': attribute,
    'action': lambda _: value}` (abs.py:74-84, applied at abs.py:267-271), e.g. the conditional lockdown of
    run_batch.py:628-643.  Like the reference under `batch_experiment`, the condition sees the statistics
    memoised after the previous step.
    """

    def __init__(self, metric, op, threshold, attribute, value):
        if metric not in STAT_KEYS or op not in _OPS or attribute not in _ATTRS:
            raise ValueError("unsupported trigger: {} {} {}".format(metric, op, attribute))
        self.metric, self.op, self.threshold, self.attribute, self.value = metric, op, float(threshold), attribute, value

    def __getitem__(self, key):  # lets code that inspects trigger dicts keep working (abs.py:237-238)
        return {'attribute': self.attribute, 'condition': None, 'action': None}[key]


def amplitude_array(amplitudes):
    """abs.py:33-36: dict keyed by Status -> [S, I, R, D] array (Death never moves, abs.py:164)."""
    out = np.zeros(4, dtype=np.float64)
    for status, value in amplitudes.items():
        out[STATUS_CODE[status] if isinstance(status, Status) else int(status)] = float(value)
    out[3] = 0.0
    return out


def wealth_scale_bits(total_wealth):
    """Fixed-point scale of the Q1..Q5 sums: largest s <= 40 with 1024 * total_wealth * 2**s < 2**62."""
    tw = max(1.0, float(total_wealth))
    s = int(np.floor(62 - np.log2(1024.0 * tw)))
    return max(0, min(40, s))


# constructor keyword -> default (abs.py:17-49); values the reference's scenario dictionaries may pass
_REFERENCE_KWARGS = (
    ("population_size", 20),          # number of agents
    ("length", 10), ("height", 10),   # extent of the shared environment
    ("initial_infected_perc", 0.05), ("initial_immune_perc", 0.05),
    ("contagion_distance", 1.),       # contact = distance at most this
    ("contagion_rate", 0.9),          # probability of contagion per contact
    ("critical_limit", 0.6),          # fraction of the population the health system can treat
    ("minimum_income", 1.0), ("minimum_expense", 1.0),   # daily units of the poorest quintile
    ("total_wealth", 10 ** 4),
)
# keyword -> default of what this implementation adds (nothing the reference passes is reinterpreted)
_EXTRA_KWARGS = (
    ("seed", None),                   # Philox seed; None = drawn from np.random at initialize(), so np.random.seed() governs
    ("device", 0),                    # CUDA device ordinal
    ("placement", "reference"),       # "reference": abs.py:97-114 from np.random in the reference's order; "uniform": x, y ~ U{0..L}
    ("schedule", "synchronous"),      # or "sequential": the reference's pair order with in-step chains (abs.py:261-265), exact, slower
    ("max_cells", 0), ("history_capacity", 4096),
    ("positions", "lattice"),         # "float64": binary64 positions and the reference's floating-point contact test
                                      # (chosen automatically as soon as a population trigger or a float position appears)
)


class Simulation(object):
    def __init__(self, **kwargs):
        for name, default in _REFERENCE_KWARGS + _EXTRA_KWARGS:
            setattr(self, name, kwargs.get(name, default))
        # mobility (standard deviation of a move) by status; Death never moves
        self.amplitudes = kwargs.get('amplitudes', {st: 5 for st in (Status.Susceptible, Status.Recovered_Immune,
                                                                     Status.Infected)})
        self.triggers_simulation = kwargs.get("triggers_simulation", [])   # conditional rewrites of simulation attributes
        self.triggers_population = kwargs.get("triggers_population", [])   # conditional rewrites of agent attributes
        self.statistics = None        # memo of get_statistics(), dropped at the end of execute()

        self._handle = None
        self._pushed = None          # last parameter tuple sent to the device
        self._pushed_triggers = None
        self._pushed_rules = None
        self._trigger_epoch = 0      # device-side trigger firings already mirrored into the attributes
        self._stats_values = None    # statistics fetched together with the trigger state, valid until the next step
        self._host = None            # cached download of the population (dict of columns)
        self._agents = None          # cached list of Agent objects
        self._iteration = 0

    # ------------------------------------------------------------------ reference helpers
    def _xclip(self, x):
        return np.clip(int(x), 0, self.length)

    def _yclip(self, y):
        return np.clip(int(y), 0, self.height)

    def set_amplitudes(self, amp):
        self.amplitudes = amp

    def append_trigger_simulation(self, condition, attribute, action):
        """abs.py:74-84.  Python lambdas are evaluated on the host after each step."""
        self.triggers_simulation.append({'condition': condition, 'attribute': attribute, 'action': action})

    def append_trigger_population(self, condition, attribute, action):
        """abs.py:86-95.  Accepted for API compatibility; see execute()."""
        self.triggers_population.append({'condition': condition, 'attribute': attribute, 'action': action})

    def random_position(self):
        # abs.py:97-101
        x = self._xclip(self.length / 2 + (np.random.randn(1)[0] * (self.length / 3)))
        y = self._yclip(self.height / 2 + (np.random.randn(1)[0] * (self.height / 3)))
        return x, y

    # ------------------------------------------------------------------ population on the device
    def _ensure_handle(self, n):
        if self._handle is None or self._handle.capacity < n:
            if self._handle is not None:
                self._handle.close()
            if self.seed is None:
                self.seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
            self._handle = _native.Handle(capacity=max(1, n), seed=self.seed, device=self.device,
                                          max_cells=self.max_cells, history_capacity=self.history_capacity,
                                          f64=self._wants_f64())
            self._pushed = None
            self._pushed_triggers = None
            self._pushed_rules = None
        return self._handle

    def _wants_f64(self):
        if self.positions not in ("lattice", "float64"):
            raise ValueError("positions must be 'lattice' or 'float64'")
        return self.positions == "float64" or len(self.triggers_population) > 0

    def _push_rules(self):
        """Population triggers (abs.py:86-95) as device rules; switches the handle to binary64 positions when the
        first one appears on a lattice handle."""
        if not self.triggers_population:
            if self._pushed_rules:
                self._handle.set_rules([], [])
                self._pushed_rules = None
            return
        rules = [lower_trigger(t) for t in self.triggers_population]
        key = tuple(r.key() for r in rules)
        if not self._handle.f64:
            self._to_float64()
        if key != self._pushed_rules:
            self._handle.set_rules([r.native() for r in rules if isinstance(r, MoveRule)],
                                   [r.native() for r in rules if isinstance(r, AttributeRule)])
            self._pushed_rules = key

    def _to_float64(self):
        """Re-create the handle in off-lattice mode with the same population and the same draw clock."""
        cols = self._handle.download(int(self.population_size))
        step = self._handle.info()["step"]
        self._handle.close()
        self._handle = None
        self.positions = "float64"
        self._ensure_handle(len(cols["x"]))
        self._push_params()
        xf, yf = cols.pop("x").astype(np.float64), cols.pop("y").astype(np.float64)
        self._handle.upload_f64(cols, xf, yf)
        self._handle.set_step(step)
        self._invalidate_host()

    def _targets(self):
        """Handles that hold a part of this population (several in slab mode)."""
        return [self._handle]

    def _param_tuple(self):
        amp = amplitude_array(self.amplitudes)
        return (float(self.length), float(self.height), float(self.contagion_distance), float(self.contagion_rate),
                float(self.critical_limit), tuple(amp.tolist()), float(self.minimum_income),
                float(self.minimum_expense), int(self.population_size), float(self.total_wealth), self.schedule)

    def _push_params(self):
        """Send the attributes of abs.py:17-49 to the device if user code changed them (abs.py:267-271)."""
        t = self._param_tuple()
        if t != self._pushed:
            p = _native.Params()
            (p.length, p.height, p.contagion_distance, p.contagion_rate, p.critical_limit, amp, p.minimum_income,
             p.minimum_expense, pop, tw, _schedule) = t
            p.amplitudes[:] = amp
            p.basic_income[:] = [float(v) for v in basic_income]
            p.age_hospitalization_probs[:] = age_hospitalization_probs
            p.age_severe_probs[:] = age_severe_probs
            p.age_death_probs[:] = age_death_probs
            p.population_size = max(1, pop)
            p.recovering_time = 20  # abs.py:225
            p.wealth_scale_bits = wealth_scale_bits(tw)
            if self.schedule not in ("synchronous", "sequential"):
                raise ValueError("schedule must be 'synchronous' or 'sequential'")
            p.schedule = _native.SEQUENTIAL if self.schedule == "sequential" else _native.SYNCHRONOUS
            for h in self._targets():
                h.set_params(p)
            self._pushed = t
        decl = [t for t in self.triggers_simulation if isinstance(t, Trigger)]
        key = tuple((t.metric, t.op, t.threshold, t.attribute, repr(t.value)) for t in decl)
        if key != self._pushed_triggers:
            native = []
            for t in decl:
                n = _native.Trigger()
                n.metric, n.op, n.threshold, n.attribute = STAT_KEYS.index(t.metric), _OPS[t.op], t.threshold, _ATTRS[
                    t.attribute]
                if t.attribute == 'amplitudes':
                    n.value[:] = amplitude_array(t.value).tolist()
                else:
                    n.value[:] = [float(t.value), 0.0, 0.0, 0.0]
                native.append(n)
            for h in self._targets():
                h.set_triggers(native)
            self._pushed_triggers = key

    def _refresh_from_device(self):
        """A fired declarative trigger rewrote simulation attributes on the device (abs.py:267-271: the change is
        permanent).  Mirror them into the Python attributes, and into the record of what was last pushed, so that a
        later change of any other attribute does not send stale values back.  One copy and one synchronisation fetch
        the statistics of the step as well (kept for the get_statistics() that follows), and the attributes are only
        touched when the device says a trigger has fired since the last look."""
        if self._handle is None or self._pushed is None:
            return
        if not any(isinstance(t, Trigger) for t in self.triggers_simulation):
            return
        if self._param_tuple() != self._pushed:
            return          # user code changed attributes since the last push: they win at the next _push_params()
        values, _, p, epoch = self._targets()[0].stats_ex()
        self._stats_values = values
        if epoch == self._trigger_epoch:
            return
        self._trigger_epoch = epoch
        amp = [float(v) for v in p.amplitudes]
        for status, code in STATUS_CODE.items():
            if status in self.amplitudes:
                keep_int = isinstance(self.amplitudes[status], int) and float(int(amp[code])) == amp[code]
                self.amplitudes[status] = int(amp[code]) if keep_int else amp[code]
        self.contagion_rate = float(p.contagion_rate)
        self.contagion_distance = float(p.contagion_distance)
        self.critical_limit = float(p.critical_limit)
        self._pushed = self._param_tuple()

    def set_population_arrays(self, x, y, status, age, stratum, wealth, severity=None, infected_time=None):
        """Upload a population given as columns (codes of include/covidabs.h); ids = positions."""
        n = len(x)
        xa, ya = np.asarray(x), np.asarray(y)
        fractional = (xa.dtype.kind == 'f' and bool(np.any(xa != np.floor(xa)))) or \
                     (ya.dtype.kind == 'f' and bool(np.any(ya != np.floor(ya))))
        if fractional:
            self.positions = "float64"
        cols = {
            "x": np.ascontiguousarray(np.floor(xa), dtype=np.int32), "y": np.ascontiguousarray(np.floor(ya), dtype=np.int32),
            "status": np.ascontiguousarray(status, dtype=np.uint8),
            "severity": (np.full(n, 1, dtype=np.uint8) if severity is None
                         else np.ascontiguousarray(severity, dtype=np.uint8)),
            "infected_time": (np.zeros(n, dtype=np.uint8) if infected_time is None
                              else np.ascontiguousarray(infected_time, dtype=np.uint8)),
            "age": np.ascontiguousarray(np.minimum(age, 255), dtype=np.uint8),
            "stratum": np.ascontiguousarray(stratum, dtype=np.uint8),
            "wealth": np.ascontiguousarray(wealth, dtype=np.float64),
        }
        self.population_size = n
        if self._handle is not None and self._handle.f64 != self._wants_f64():
            self._handle.close()
            self._handle = None
        self._ensure_handle(n)
        self._push_params()
        if self._handle.f64:
            self._handle.upload_f64(cols, np.asarray(xa, dtype=np.float64), np.asarray(ya, dtype=np.float64))
        else:
            self._handle.upload(cols)
        self._invalidate_host()
        self.statistics = None

    def set_population_tensors(self, x, y, status, age, stratum, wealth, severity=None, infected_time=None):
        """Like set_population_arrays for columns that already live on the GPU: torch CUDA tensors (int32 x / y, uint8
        status / age / stratum / severity / infected_time, float64 wealth) on this simulation's device.  They are
        packed straight into the device records (cabs_bind_device): nothing crosses the PCIe bus."""
        import torch
        n = int(x.numel())
        dev = torch.device("cuda", int(self.device))
        want = (("x", x, torch.int32), ("y", y, torch.int32), ("status", status, torch.uint8), ("age", age, torch.uint8),
                ("stratum", stratum, torch.uint8), ("wealth", wealth, torch.float64),
                ("severity", torch.ones(n, dtype=torch.uint8, device=dev) if severity is None else severity, torch.uint8),
                ("infected_time", torch.zeros(n, dtype=torch.uint8, device=dev) if infected_time is None else infected_time,
                 torch.uint8))
        cols = {}
        for name, t, dt in want:
            if t.device != dev or t.dtype != dt or t.numel() != n:
                raise ValueError("{}: expected a {} tensor of {} elements on {}".format(name, dt, n, dev))
            cols[name] = t.contiguous()
        self.population_size = n
        if self._handle is not None and self._handle.f64:
            raise NotImplementedError("device columns are lattice positions; use set_population_arrays for float64")
        self._ensure_handle(n)
        self._push_params()
        torch.cuda.current_stream(dev).synchronize()      # the columns are complete before the library's stream reads them
        self._handle.bind_device(n, {k: t.data_ptr() for k, t in cols.items()})
        self._handle.sync()                               # ... and consumed before the caller may free them
        self._invalidate_host()
        self.statistics = None

    def _invalidate_host(self):
        self._host = None
        self._agents = None
        self._stats_values = None

    def _download(self):
        if self._host is None:
            if self._handle is None:
                self._host = {name: np.zeros(0, dtype=dt) for name, dt in _native.SOA_DTYPES}
            else:
                self._host = self._handle.download(int(self.population_size))
                if self._handle.f64:   # the exact positions (the int columns hold their floors)
                    self._host["x"], self._host["y"] = self._handle.download_positions(len(self._host["x"]))
        return self._host

    def get_population(self):
        """Current population as `Agent` objects, materialised from a device download (abs.py:57-63)."""
        if self._agents is None:
            h = self._download()
            num = float if h["x"].dtype.kind == 'f' else int
            self._agents = [Agent(id=int(h["id"][k]), x=num(h["x"][k]), y=num(h["y"][k]),
                                  status=STATUS_LIST[h["status"][k]],
                                  infected_status=SEVERITY_LIST[h["severity"][k]],
                                  infected_time=int(h["infected_time"][k]), age=int(h["age"][k]),
                                  social_stratum=int(h["stratum"][k]), wealth=float(h["wealth"][k]),
                                  environment=self) for k in range(len(h["x"]))]
        return self._agents

    def set_population(self, pop):
        """Replace the population with a list of Agent-like objects (abs.py:65-69)."""
        pop = list(pop)
        self.set_population_arrays(
            [float(np.asarray(a.x).ravel()[0]) for a in pop], [float(np.asarray(a.y).ravel()[0]) for a in pop],
            [STATUS_CODE[a.status] for a in pop],
            [int(a.age) for a in pop], [int(a.social_stratum) for a in pop],
            [float(np.asarray(a.wealth).ravel()[0]) for a in pop],
            severity=[SEVERITY_CODE[a.infected_status] for a in pop],
            infected_time=[int(a.infected_time) for a in pop])

    population = property(get_population, set_population)

    # ------------------------------------------------------------------ initialize
    def create_agent(self, status):
        """abs.py:103-114 (draw order: x, y, age, stratum)."""
        x, y = self.random_position()
        age = int(np.random.beta(2, 5, 1)[0] * 100)
        social_stratum = int(np.random.rand(1)[0] * 100 // 20)
        return int(x), int(y), age, social_stratum

    def initialize(self):
        """Create the population (abs.py:116-139) on the host and upload it."""
        n = int(self.population_size)
        n_inf = int(n * self.initial_infected_perc)
        n_imm = int(n * self.initial_immune_perc)
        status = np.zeros(n, dtype=np.uint8)
        status[:n_inf] = STATUS_CODE[Status.Infected]
        status[n_inf:n_inf + n_imm] = STATUS_CODE[Status.Recovered_Immune]
        if self.placement == "reference":
            rows = [self.create_agent(None) for _ in range(n)]
            cols = np.array(rows, dtype=np.int64).reshape(n, 4)
            x, y, age, stratum = cols[:, 0], cols[:, 1], cols[:, 2], cols[:, 3]
        elif self.placement == "uniform":
            # synthetic uniform lattice placement (large N, no Python loop)
            if self.seed is None:
                self.seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
            rng = np.random.default_rng(self.seed)
            x = rng.integers(0, int(self.length) + 1, size=n)
            y = rng.integers(0, int(self.height) + 1, size=n)
            age = (rng.beta(2, 5, size=n) * 100).astype(np.int64)
            stratum = rng.integers(0, 5, size=n)
        else:
            raise ValueError("placement must be 'reference' or 'uniform'")
        wealth = np.zeros(n, dtype=np.float64)
        for quintile in [0, 1, 2, 3, 4]:  # abs.py:134-139
            total = lorenz_curve[quintile] * self.total_wealth
            sel = (stratum == quintile) & (age >= 18)
            qty = max(1.0, float(np.sum(sel)))
            wealth[sel] = total / qty
        self.set_population_arrays(x, y, status, age, stratum, wealth)
        self._iteration = 0

    # ------------------------------------------------------------------ step
    def _host_triggers(self):
        return [t for t in self.triggers_simulation if not isinstance(t, Trigger)]

    def execute(self):
        """One iteration (abs.py:232-273): move, update, contacts on the GPU; simulation triggers after."""
        if self._handle is None:
            raise RuntimeError("initialize() or set_population() first")
        self._push_params()
        self._push_rules()
        self._handle.step(1)
        self._iteration += 1
        self._invalidate_host()
        self._refresh_from_device()
        for trigger in self._host_triggers():  # abs.py:267-271
            if trigger['condition'](self):
                attr = trigger['attribute']
                self.__dict__[attr] = trigger['action'](self.__dict__[attr])
        self.statistics = None

    def run(self, iterations):
        """`iterations` x (execute(); get_statistics('all')) -> float64 array [iterations, 12] (key order STAT_KEYS).

        With only declarative triggers the whole run is enqueued at once and the statistics rows come
        back in one copy; Python-lambda triggers force one host round trip per iteration.
        """
        iterations = int(iterations)
        if self._host_triggers() or iterations > self.history_capacity:
            rows = np.empty((iterations, len(STAT_KEYS)))
            for it in range(iterations):
                self.execute()
                s = self.get_statistics(kind='all')
                rows[it] = [s[k] for k in STAT_KEYS]
            return rows
        self._push_params()
        self._push_rules()
        first = self._handle.info()["step"]     # the handle's clock is not reset by a re-upload
        self._handle.step(iterations)
        self._iteration += iterations
        self._invalidate_host()
        self.statistics = None
        rows = self._handle.stats_history(first, iterations)
        self._refresh_from_device()
        return rows

    # ------------------------------------------------------------------ accessors
    def get_positions(self, stride=1):
        """abs.py:275-277.  stride > 1: every stride-th agent only (what an animation of millions of agents can draw:
        `cabs_snapshot` moves 10 bytes per sampled agent instead of the whole population)."""
        if stride > 1:
            s = self.snapshot(stride)
            return [[int(a), int(b)] for a, b in zip(s["x"], s["y"])]
        h = self._download()
        num = float if h["x"].dtype.kind == 'f' else int
        return [[num(a), num(b)] for a, b in zip(h["x"], h["y"])]

    def snapshot(self, stride=1):
        """Columns x, y, status, severity (codes of include/covidabs.h) of every stride-th agent of the population list:
        the per-frame feed of graphics.py:113-142,241-276 (`get_positions()` + `get_description()`), strided."""
        if self._handle is None:
            raise RuntimeError("initialize() or set_population() first")
        return self._handle.snapshot(int(self.population_size), int(stride))

    def remove_trigger_population(self, attribute):
        """Drop the population triggers on `attribute` (used by run_batch.py's commented scenarios)."""
        def attr_of(t):
            return t['attribute'] if isinstance(t, dict) else ('move' if isinstance(t, MoveRule) else t.attribute)
        self.triggers_population = [t for t in self.triggers_population if attr_of(t) != attribute]

    def get_description(self, complete=False, stride=1):
        """abs.py:279-289 (stride > 1: every stride-th agent, see get_positions)."""
        if stride > 1:
            s = self.snapshot(stride)
            if complete:
                return [STATUS_LIST[a].name if STATUS_LIST[a] != Status.Infected else
                        "{}({})".format(STATUS_LIST[a].name, SEVERITY_LIST[b].name) for a, b in zip(s["status"], s["severity"])]
            return [STATUS_LIST[a].name for a in s["status"]]
        if complete:
            return [a.get_description() for a in self.get_population()]
        h = self._download()
        return [STATUS_LIST[s].name for s in h["status"]]

    def get_contact_pairs(self):
        """The pair list of abs.py:253-259 for the current positions, sorted lexicographically (test hook)."""
        self._push_params()
        pairs = self._handle.contact_pairs()
        order = np.lexsort((pairs[:, 1], pairs[:, 0]))
        return pairs[order]

    def get_statistics(self, kind='info'):
        """abs.py:291-314: memoised until the end of the next execute()."""
        if self.statistics is None:
            if self._handle is None:
                raise RuntimeError("initialize() or set_population() first")
            self._push_params()
            values = self._stats_values
            if values is None:
                values, _ = self._handle.stats()
            self.statistics = {k: np.float64(v) for k, v in zip(STAT_KEYS, values)}
        return self.filter_stats(kind)

    def filter_stats(self, kind):
        # abs.py:316-322
        if kind == 'info':
            return {k: v for k, v in self.statistics.items()
                    if not k.startswith('Q') and k not in ('Business', 'Government')}
        elif kind == 'ecom':
            return {k: v for k, v in self.statistics.items() if k.startswith('Q') or k in ('Business', 'Government')}
        else:
            return self.statistics

    def device_info(self):
        return self._handle.info()

    def __str__(self):
        return str(self.get_description())


class MultiPopulationSimulation(Simulation):
    """Several populations on one plane (abs.py:328-415): every `Simulation` of `simulations` steps on its own, then the
    agents of different populations meet across the boxes placed at `positions`.

    The cross contacts run on the device (`cabs_cross_contact`) under the synchronous rule of the step itself: an agent
    infected by a cross contact starts transmitting at the next phase, not later inside the same double loop.  Offsets
    must be integers (lattice).  The quirk of abs.py:389-410 is kept: every statistic is the LAST population's count
    over `total_population` (each population overwrites the previous one's entry), wealth sums are the last one's.
    """

    def __init__(self, **kwargs):
        super(MultiPopulationSimulation, self).__init__(**kwargs)
        self.simulations = kwargs.get('simulations', [])
        self.positions = kwargs.get('positions', [])
        self.total_population = kwargs.get('total_population', 0)

    def get_population(self):
        population = []
        for simulation in self.simulations:
            population.extend(simulation.get_population())
        return population

    def append(self, simulation, position):
        self.simulations.append(simulation)
        self.positions.append(position)
        self.total_population += simulation.population_size

    def initialize(self):
        if self.seed is None:
            self.seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        for simulation in self.simulations:
            simulation.initialize()
        self._iteration = 0
        self.statistics = None

    def _offset(self, k):
        px, py = self.positions[k]
        if int(px) != px or int(py) != py:
            raise NotImplementedError("populations are placed on the integer lattice: offsets must be integers")
        return int(px), int(py)

    def execute(self, **kwargs):
        if self.seed is None:
            raise RuntimeError("initialize() first")
        for simulation in self.simulations:
            simulation.execute()
        self._iteration += 1
        count = len(self.simulations)
        for m in range(count):
            for n in range(m + 1, count):
                a, b = self.simulations[m], self.simulations[n]
                a._push_params()
                b._push_params()
                a._handle.cross_contact(b._handle, self._offset(m), self._offset(n), self.contagion_distance,
                                        self.seed, m * count + n)
                for sim in (a, b):
                    sim._invalidate_host()
                    sim.statistics = None
        self.statistics = None

    def run(self, iterations):
        """`iterations` x (execute(); get_statistics('all')) -> float64 [iterations, 13], columns in the key order of
        get_statistics (the reference's dictionary has an extra, always zero, 'Exposed' entry here)."""
        rows = []
        for it in range(int(iterations)):
            self.execute()
            rows.append(list(self.get_statistics(kind='all').values()))
        return np.array(rows, dtype=np.float64).reshape(int(iterations), -1)

    def get_positions(self):
        positions = []
        for ct, simulation in enumerate(self.simulations):
            h = simulation._download()
            px, py = self.positions[ct]
            positions.extend([[int(x) + px, int(y) + py] for x, y in zip(h["x"], h["y"])])
        return positions

    def get_description(self, complete=False):
        out = []
        for simulation in self.simulations:
            out.extend(simulation.get_description(complete))
        return out

    def get_statistics(self, kind='info'):
        if self.statistics is None:
            if not self.simulations:
                raise RuntimeError("no populations")
            last = self.simulations[-1]
            last._push_params()
            values, counts = last._handle.stats()
            stats = {}
            for k, key in enumerate(STAT_KEYS):
                if k == 4:     # abs.py:396 iterates over every InfectionSeverity, Exposed included (never assigned)
                    stats['Exposed'] = np.float64(0.0) / self.total_population
                if k < 7:      # S, I, R, D, Asymptomatic, Hospitalization, Severe: counts of the last population
                    stats[key] = np.float64(counts[k]) / self.total_population
                else:          # Q1..Q5: the last population's sums, not normalised (abs.py:403-408)
                    stats[key] = np.float64(values[k])
            self.statistics = stats
        return self.filter_stats(kind)
