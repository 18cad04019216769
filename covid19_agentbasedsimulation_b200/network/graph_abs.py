"""Host-side mirror of covid_abs/network/graph_abs.py: `GraphSimulation` with its hourly step on a B200.

Same constructor kwargs, attributes and methods as the reference class (graph_abs.py:12-324).  `initialize()` builds
the world on the host exactly as the reference does -- the same `np.random` calls in the same order
(graph_abs.py:89-188, network/agents.py:36-42,186-194), so `np.random.seed(s)` gives the reference's world bit for bit
(tests/test_graph_host.py) -- and uploads it as structure of arrays; `execute()` runs the per-hour kernel of
csrc/graph_kernels.cuh through the C ABI (`cabs_graph_*`, include/covidabs.h).  There is no CPU fallback.

Callbacks: the reference takes Python lambdas (run_batch.py:87-213); the behaviours it ships are named in
`network/scenarios.py` and lowered to kernel flags.  The reference's own callables (`lambda x: sleep(x)`,
`lambda x: lockdown(x)`, `lambda x: pset(x, 'contagion_rate', 0.1)`, ...) are recognised by probing them with
instrumented stand-ins (`scenarios.recognise_callbacks`), so unmodified run_batch.py scenario dictionaries work; a
callable that matches no shipped behaviour raises NotImplementedError -- arbitrary Python cannot run on the device.

Differences a user can observe: draws are counter-based (Philox keyed (seed; entity,
iteration, stream)), money that several agents add to is fixed point, statistics read inside a step are those of
the previous executed hour.  Fed the same draws, the unmodified reference walks the same trajectory
(tests/test_graph_oracle.py).
"""
import numpy as np

from covid19_agentbasedsimulation_b200 import _native
from covid19_agentbasedsimulation_b200.abs import Simulation, amplitude_array
from covid19_agentbasedsimulation_b200.agents import (Status, InfectionSeverity, AgentType, STATUS_LIST, SEVERITY_LIST,
                                                      STATUS_CODE)
from covid19_agentbasedsimulation_b200.common import (age_hospitalization_probs, age_severe_probs, age_death_probs,
                                                      lorenz_curve, basic_income)
from covid19_agentbasedsimulation_b200.network.agents import EconomicalStatus, Business, House, Person
from covid19_agentbasedsimulation_b200.network.scenarios import Declarative, recognise_callbacks
from covid19_agentbasedsimulation_b200.network.util import (new_day, work_day, new_month, bed_time, work_time,  # noqa: F401
                                                            lunch_time, free_time)

GRAPH_STAT_KEYS = ("Q1", "Q2", "Q3", "Q4", "Q5", "Business", "Government", "Susceptible", "Infected",
                   "Recovered_Immune", "Death", "Asymptomatic", "Hospitalization", "Severe")   # graph_abs.py:305-322


_RECOGNISED = {}   # callbacks dicts whose Python callables have been probed already


def default_money_bits(total_wealth):
    """Fractional bits of the fixed-point accounts: 2^-24 units for the shipped total_wealth of 1e7."""
    tw = max(1.0, float(total_wealth))
    return int(max(8, min(24, np.floor(56 - np.log2(tw)))))


def lower_callbacks(callbacks, sim=None):
    """run_batch.py-style callbacks dict -> (mode, sleep, isolation_rate, overrides).  Entries are the named
    behaviours of network/scenarios.py, or the reference's own Python callables (run_batch.py:8-50,173-175), which are
    identified by probing (scenarios.recognise_callbacks); any other callable raises NotImplementedError."""
    mode, sleep, rate, overrides, isolation_move = _native.GRAPH_MODE_NONE, False, None, {}, False
    if any(not isinstance(a, Declarative) for a in (callbacks or {}).values()):
        key = tuple(sorted((e, id(a)) for e, a in callbacks.items()))
        cached = _RECOGNISED.get(key)
        if cached is None or cached[0] is not callbacks:      # probe once per callbacks dict, not once per call
            cached = (callbacks, recognise_callbacks(callbacks, sim))
            _RECOGNISED[key] = cached
        callbacks = cached[1]
    for event, action in (callbacks or {}).items():
        kind = action.kind
        if event == 'on_execute' and kind == 'sleep':
            sleep = True
        elif event == 'on_person_move' and kind == 'lockdown':
            mode = _native.GRAPH_MODE_LOCKDOWN
        elif event == 'on_person_move' and kind == 'conditional_lockdown':
            mode = _native.GRAPH_MODE_CONDITIONAL_LOCKDOWN
        elif event == 'on_person_move' and kind == 'vertical_isolation':
            mode = _native.GRAPH_MODE_VERTICAL_ISOLATION
        elif event == 'on_person_move' and kind == 'check_isolation':
            isolation_move = True
        elif event == 'post_initialize' and kind == 'sample_isolated':
            rate = action.args['isolation_rate']
        elif event == 'on_initialize' and kind == 'pset':
            if 'overrides' in action.args:
                overrides.update(action.args['overrides'])
            else:
                overrides[action.args['attribute']] = action.args['value']
        else:
            raise NotImplementedError("unsupported callback {} -> {}".format(event, action))
    if (rate is None) != (not isolation_move):
        raise NotImplementedError("partial isolation needs both post_initialize=sample_isolated(rate) and "
                                  "on_person_move=check_isolation (run_batch.py:136-146)")
    return mode, sleep, rate, overrides


# constructor keyword -> default of what GraphSimulation adds to Simulation (graph_abs.py:15-32)
_GRAPH_KWARGS = (
    ("total_population", 0), ("total_business", 10), ("business_distance", 10),
    ("homeless_rate", 0.01), ("unemployment_rate", 0.09), ("homemates_avg", 3), ("homemates_std", 1),
    ("public_gdp_share", 0.1), ("business_gdp_share", 0.5),
    ("incubation_time", 5), ("contagion_time", 10), ("recovering_time", 20),
    ("money_bits", None),             # new, optional: fractional bits of the fixed-point accounts
)


class GraphSimulation(Simulation):
    def __init__(self, **kwargs):
        super(GraphSimulation, self).__init__(**kwargs)
        for name, default in _GRAPH_KWARGS:
            setattr(self, name, kwargs.get(name, default))
        self.callbacks = kwargs.get('callbacks', {})
        self.iteration = -1           # the first execute() is iteration 0 (graph_abs.py:25,192)
        self._ghandle = None
        self._gsim = 0                # index of this simulation inside the handle (ensembles share one handle)
        self._world = None            # host copy of the current world (dict of columns), refreshed lazily
        self._entities = None         # (persons, houses, businesses, government, healthcare) materialised on request

    def register_callback(self, event, action):
        self.callbacks[event] = action

    # ------------------------------------------------------------------ world building (graph_abs.py:89-188)
    def build_world(self):
        """The lists `initialize` builds, as columns.  Consumes np.random exactly like the reference."""
        mode, sleep, iso_rate, overrides = lower_callbacks(self.callbacks, self)
        for k, v in overrides.items():                      # on_initialize (graph_abs.py:94)
            setattr(self, k, v)
        N = int(self.population_size)
        hx, hy, hstratum = [], [], []
        bx, by, bstratum = [], [], []

        health_x, health_y = self.random_position()          # graph_abs.py:96-98
        health_fixed = 0.0 + self.minimum_expense * 3
        gov_x, gov_y = self.random_position()                # graph_abs.py:99-102
        gov_fixed = 0.0 + self.population_size * (self.minimum_expense * 0.05)

        def create_house(stratum):
            x, y = self.random_position()
            hx.append(int(x)); hy.append(int(y)); hstratum.append(int(stratum))

        for i in np.arange(0, int(self.population_size // self.homemates_avg)):   # 105-106
            create_house(i % 5)
        for i in np.arange(0, self.total_business):                                # 109-110
            x, y = self.random_position()
            bx.append(int(x)); by.append(int(y)); bstratum.append(int(5 - (i % 5)))

        status, age, stratum, itime = [], [], [], []

        def create_agent(st, social_stratum=None, infected_time=0):               # 72-87
            a = int(np.random.beta(2, 5, 1)[0] * 100)
            if social_stratum is None:
                social_stratum = int(np.random.rand(1)[0] * 100 // 20)
            status.append(STATUS_CODE[st]); age.append(a); stratum.append(int(social_stratum))
            itime.append(int(infected_time))

        for i in np.arange(0, int(self.population_size * self.initial_infected_perc)):   # 113-114
            create_agent(Status.Infected, infected_time=5)
        for i in np.arange(0, int(self.population_size * self.initial_immune_perc)):     # 117-118
            create_agent(Status.Recovered_Immune)
        for i in np.arange(0, self.population_size - len(status)):                        # 121-122
            create_agent(Status.Susceptible, social_stratum=i % 5)

        n = len(status)
        age_a, stratum_a = np.array(age, dtype=np.int64), np.array(stratum, dtype=np.int64)
        active = (age_a > 16) & (age_a <= 65)                                             # network/agents.py:270-271
        pwealth = np.zeros(n)
        px, py = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        house = np.full(n, -1, dtype=np.int64)
        employer = np.full(n, -1, dtype=np.int64)
        hire_seq = np.full(n, -1, dtype=np.int64)
        expenses = np.zeros(n)
        nb = len(bx)
        bwealth = np.zeros(nb)
        bnum = np.zeros(nb, dtype=np.int64)
        bfixed = np.zeros(nb)
        employees = [[] for _ in range(nb)]
        hwealth, hsize, hfixed = [0.0] * len(hx), [0] * len(hx), [0.0] * len(hx)
        gov_wealth = self.total_wealth * self.public_gdp_share                           # 126

        for quintile in [0, 1, 2, 3, 4]:                                                  # 128-186
            _houses = [k for k in range(len(hx)) if hstratum[k] == quintile]
            if len(_houses) == 0:
                create_house(quintile)
                hwealth.append(0.0); hsize.append(0); hfixed.append(0.0)
                _houses = [len(hx) - 1]
            nhouses = len(_houses)
            if self.total_business > 5:
                btotal = lorenz_curve[quintile] * (self.total_wealth * self.business_gdp_share)
                bqty = max(1.0, np.sum([1.0 for s in bstratum if s == quintile]))
            else:
                btotal = self.total_wealth * self.business_gdp_share
                bqty = self.total_business
            ag_share = btotal / bqty
            for k in range(nb):
                if bstratum[k] == quintile:
                    bwealth[k] = ag_share
            ptotal = lorenz_curve[quintile] * self.total_wealth * (1 - (self.public_gdp_share + self.business_gdp_share))
            pqty = max(1.0, np.sum([1 for k in range(n) if stratum_a[k] == quintile and active[k]]))
            ag_share = ptotal / pqty
            for k in range(n):
                if stratum_a[k] != quintile:
                    continue
                if active[k]:
                    pwealth[k] = ag_share
                    unemployed_test = np.random.rand()
                    if unemployed_test >= self.unemployment_rate:
                        ix = np.random.randint(0, self.total_business)
                        # Business.hire (network/agents.py:36-42): agent.expenses is still 0.0 here, so the
                        # employer's fixed_expenses do not grow (the reference sets expenses afterwards, 171)
                        employees[ix].append(k)
                        employer[k] = ix
                        bnum[ix] += 1
                        bfixed[ix] += (expenses[k] / 720) * 24
                expenses[k] = basic_income[stratum_a[k]] * self.minimum_expense
                homeless_test = np.random.rand()
                if not (quintile == 0 and homeless_test <= self.homeless_rate):
                    def append_mate(h):                                                    # network/agents.py:186-194
                        hwealth[h] += pwealth[k]
                        hsize[h] += 1
                        house[k] = h
                        dx, dy = np.random.normal(0.0, 0.25, 2)
                        px[k] = int(hx[h] + dx)
                        py[k] = int(hy[h] + dy)
                        hfixed[h] += (expenses[k] / 720) * 24
                    for kp in range(0, 5):                                                 # 177-182 (`continue`, not `break`)
                        ix = np.random.randint(0, nhouses)
                        h = _houses[ix]
                        if hsize[h] < self.homemates_avg + self.homemates_std:
                            append_mate(h)
                            continue
                    if house[k] < 0:                                                       # 183-186: global list, local count
                        ix = np.random.randint(0, nhouses)
                        append_mate(ix)
        seq = 0
        for b in range(nb):
            for k in employees[b]:
                hire_seq[k] = seq
                seq += 1
        isolated = np.zeros(n, dtype=bool)
        if iso_rate is not None:                                                          # run_batch.py:31-35
            for k in range(n):
                isolated[k] = np.random.rand() <= iso_rate
        bprice = [(s + 1) * 12.0 for s in bstratum]                                       # network/agents.py:30
        world = dict(
            px=px, py=py, status=np.array(status), severity=np.full(n, 1), infected_time=np.array(itime), age=age_a,
            stratum=stratum_a, active=active, isolated=isolated, house=house, employer=employer, hire_seq=hire_seq,
            pwealth=pwealth,
            hx=np.array(hx), hy=np.array(hy), hstratum=np.array(hstratum), hsize=np.array(hsize),
            hwealth=np.array(hwealth, dtype=np.float64), hfixed=np.array(hfixed, dtype=np.float64),
            bx=np.array(bx), by=np.array(by), bstratum=np.array(bstratum), bstocks=np.full(nb, 10),
            bnum_employees=bnum, bprice=np.array(bprice), bwealth=bwealth, bfixed=bfixed,
            gov_x=int(gov_x), gov_y=int(gov_y), health_x=int(health_x), health_y=int(health_y),
            gov_wealth=float(gov_wealth), gov_price=1.0, gov_fixed=float(gov_fixed), health_wealth=0.0,
            health_fixed=float(health_fixed), health_expenses=0.0, iteration=-1)
        return world, mode, sleep

    def build_world_fast(self, seed):
        """The same construction in native code with counter-based draws keyed by `seed` (cabs_graph_build_world):
        a world of 10^5 persons in milliseconds, thread-safe.  Same distribution as build_world(), not the same draws;
        does not touch np.random."""
        mode, sleep, iso_rate, overrides = lower_callbacks(self.callbacks, self)
        for k, v in overrides.items():                      # on_initialize (graph_abs.py:94)
            setattr(self, k, v)
        spec = _native.GraphWorldSpec()
        spec.population_size, spec.length, spec.height = int(self.population_size), int(self.length), int(self.height)
        spec.total_business, spec.homemates_avg = int(self.total_business), int(self.homemates_avg)
        spec.homemates_std = int(self.homemates_std)
        spec.initial_infected_perc, spec.initial_immune_perc = float(self.initial_infected_perc), float(self.initial_immune_perc)
        spec.homeless_rate, spec.unemployment_rate = float(self.homeless_rate), float(self.unemployment_rate)
        spec.total_wealth = float(self.total_wealth)
        spec.public_gdp_share, spec.business_gdp_share = float(self.public_gdp_share), float(self.business_gdp_share)
        spec.minimum_income, spec.minimum_expense = float(self.minimum_income), float(self.minimum_expense)
        spec.basic_income[:] = [float(v) for v in basic_income]
        spec.lorenz_curve[:] = [float(v) for v in lorenz_curve]
        spec.isolation_rate = -1.0 if iso_rate is None else float(iso_rate)
        return _native.build_world_native(spec, seed), mode, sleep

    def graph_params(self, mode, sleep):
        p = _native.GraphParams()
        p.amplitudes[:] = amplitude_array(self.amplitudes).tolist()
        p.basic_income[:] = [float(v) for v in basic_income]
        p.minimum_income, p.minimum_expense = float(self.minimum_income), float(self.minimum_expense)
        p.total_wealth, p.critical_limit = float(self.total_wealth), float(self.critical_limit)
        p.contagion_distance, p.contagion_rate = float(self.contagion_distance), float(self.contagion_rate)
        p.business_distance = float(self.business_distance)
        p.age_hospitalization_probs[:] = age_hospitalization_probs
        p.age_severe_probs[:] = age_severe_probs
        p.age_death_probs[:] = age_death_probs
        p.population_size = max(1, int(self.population_size))
        p.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        p.incubation_time, p.contagion_time = int(self.incubation_time), int(self.contagion_time)
        p.recovering_time = int(self.recovering_time)
        p.mode, p.sleep = int(mode), 1 if sleep else 0
        p.money_bits = int(self.money_bits if self.money_bits is not None else default_money_bits(self.total_wealth))
        return p

    def _take_seed(self):
        if self.seed is None:
            self.seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))

    def initialize(self):
        """graph_abs.py:89-188 on the host, then upload."""
        world, mode, sleep = self.build_world()
        self.set_world(world, mode, sleep)

    def set_world(self, world, mode=_native.GRAPH_MODE_NONE, sleep=False, handle=None, index=0):
        """Upload a world given as columns (keys of cabs_graph_world).  `handle`/`index`: slot of a shared handle."""
        self._take_seed()
        for k, v in lower_callbacks(self.callbacks, self)[3].items():   # on_initialize overrides (graph_abs.py:94), idempotent
            setattr(self, k, v)
        n, nh, nb = len(world["px"]), len(world["hx"]), len(world["bx"])
        if handle is None:
            if self._ghandle is not None:
                self._ghandle.close()
            handle = _native.GraphHandle(1, max(1, n), max(1, nh), nb, device=self.device,
                                         history_capacity=self.history_capacity)
            index = 0
        self._ghandle, self._gsim = handle, int(index)
        self._mode, self._sleep = mode, sleep
        handle.upload(self._gsim, world, self.graph_params(mode, sleep))
        self.iteration = int(world.get("iteration", -1))
        self._world = None
        self.statistics = None
        self._entities = None

    # ------------------------------------------------------------------ step
    def execute(self):
        """One hourly iteration (graph_abs.py:190-276)."""
        if self._ghandle is None:
            raise RuntimeError("initialize() first")
        self._ghandle.run(1)
        self.iteration += 1
        self._world = None
        self._entities = None
        self.statistics = None

    def run(self, iterations):
        """`iterations` x (execute(); get_statistics('all')) -> float64 [iterations, 14] (key order GRAPH_STAT_KEYS),
        enqueued as ONE launch: the block that owns the simulation loops over the hours on the device."""
        iterations = int(iterations)
        if iterations > self.history_capacity:
            out = [self.run(min(self.history_capacity, iterations - k))
                   for k in range(0, iterations, self.history_capacity)]
            return np.concatenate(out, axis=0)
        first = self.iteration + 1
        self._ghandle.run(iterations)
        self.iteration += iterations
        self._world = None
        self._entities = None
        self.statistics = None
        return self._ghandle.stats(self._gsim, first, iterations)

    def get_statistics(self, kind='all'):
        """graph_abs.py:302-324."""
        if self.statistics is None:
            if self._ghandle is None:
                raise RuntimeError("initialize() first")
            row = self._ghandle.stats(self._gsim, self.iteration, 1)[0]
            self.statistics = {k: np.float64(v) for k, v in zip(GRAPH_STAT_KEYS, row)}
        return self.filter_stats(kind)

    # ------------------------------------------------------------------ accessors
    def world(self):
        """Current world as columns (device download)."""
        if self._world is None:
            self._world = self._ghandle.download(self._gsim)
        return self._world

    def _materialise(self):
        if getattr(self, "_entities", None) is None:
            w = self.world()
            houses = [House(id=k, x=int(w["hx"][k]), y=int(w["hy"][k]), social_stratum=int(w["hstratum"][k]),
                            wealth=float(w["hwealth"][k]), size=int(w["hsize"][k]), incomes=float(w["hincomes"][k]),
                            expenses=float(w["hexpenses"][k]), fixed_expenses=float(w["hfixed"][k]), environment=self)
                      for k in range(len(w["hx"]))]
            business = [Business(id=k, x=int(w["bx"][k]), y=int(w["by"][k]), social_stratum=int(w["bstratum"][k]),
                                 wealth=float(w["bwealth"][k]), price=float(w["bprice"][k]), stocks=int(w["bstocks"][k]),
                                 sales=int(w["bsales"][k]), num_employees=int(w["bnum_employees"][k]),
                                 incomes=float(w["bincomes"][k]), fixed_expenses=float(w["bfixed"][k]), environment=self)
                        for k in range(len(w["bx"]))]
            exp = np.asarray(basic_income)[np.clip(w["stratum"], 0, 4)] * self.minimum_expense
            inc = np.where(w["active"] != 0, np.asarray(basic_income)[np.clip(w["stratum"], 0, 4)] * self.minimum_income, 0.0)
            people = []
            for k in range(len(w["px"])):
                p = Person(id=k, x=int(w["px"][k]), y=int(w["py"][k]), status=STATUS_LIST[w["status"][k]],
                           infected_status=SEVERITY_LIST[w["severity"][k]], infected_time=int(w["infected_time"][k]),
                           age=int(w["age"][k]), social_stratum=int(w["stratum"][k]), wealth=float(w["pwealth"][k]),
                           income=float(inc[k]), expense=float(exp[k]), environment=self)
                if w["house"][k] >= 0:
                    p.house = houses[w["house"][k]]
                    p.house.homemates.append(p)
                if w["employer"][k] >= 0:
                    p.employer = business[w["employer"][k]]
                people.append(p)
            for b, bus in enumerate(business):   # employees in hire order (network/agents.py:38)
                members = [k for k in range(len(people)) if w["employer"][k] == b]
                bus.employees = [people[k] for k in sorted(members, key=lambda k: w["hire_seq"][k])]
            gov = Business(id=-1, x=w["gov_x"], y=w["gov_y"], social_stratum=4, price=w["gov_price"],
                           wealth=w["gov_wealth"], fixed_expenses=w["gov_fixed"], type=AgentType.Government,
                           environment=self)
            health = Business(id=-2, x=w["health_x"], y=w["health_y"], wealth=w["health_wealth"],
                              fixed_expenses=w["health_fixed"], expenses=w["health_expenses"], type=AgentType.Healthcare,
                              environment=self)
            self._entities = (people, houses, business, gov, health)
        return self._entities

    def get_population(self):
        return self._materialise()[0]

    def set_population(self, pop):
        raise NotImplementedError("GraphSimulation worlds are uploaded with set_world()")

    population = property(get_population, set_population)

    def get_positions(self):
        w = self.world()
        return [[int(a), int(b)] for a, b in zip(w["px"], w["py"])]

    def get_description(self, complete=False):
        if complete:
            return [a.get_description() for a in self.get_population()]
        return [STATUS_LIST[s].name for s in self.world()["status"]]

    def get_unemployed(self):
        """graph_abs.py:43-45."""
        return [p for p in self.get_population() if p.is_unemployed() and p.status != Status.Death
                and p.infected_status == InfectionSeverity.Asymptomatic]

    def get_homeless(self):
        """graph_abs.py:47-49."""
        return [p for p in self.get_population() if p.is_homeless() and p.status != Status.Death
                and p.infected_status == InfectionSeverity.Asymptomatic]

    def last_run_ms(self):
        return self._ghandle.last_run_ms()

    @classmethod
    def run_ensemble(cls, experiments, iterations, **kwargs):
        """What batch_experiment (experiments.py:181-196) does with `experiments` simulations, as one launch.
        Returns (statistics keys, float64 [experiments, iterations, 14])."""
        from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble
        kwargs = dict(kwargs)
        kwargs.pop('verbose', None)
        seed = kwargs.pop('seed', None)
        sims = [cls(seed=None if seed is None else int(seed) + k, **kwargs) for k in range(int(experiments))]
        ens = GraphEnsemble(sims, device=kwargs.get('device', 0),
                            history_capacity=max(2048, min(int(iterations), 1 << 16)))
        try:
            ens.initialize()
            table = ens.run(iterations)
        finally:
            ens.close()
        return list(GRAPH_STAT_KEYS), table

    @classmethod
    def run_ensemble_aggregated(cls, experiments, iterations, **kwargs):
        """run_ensemble followed by the Min/Avg/Std/Max loop of batch_experiment (experiments.py:200-208) on the device.
        Returns (statistics keys, float64 [iterations, 14, 4])."""
        from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble
        kwargs = dict(kwargs)
        kwargs.pop('verbose', None)
        seed = kwargs.pop('seed', None)
        sims = [cls(seed=None if seed is None else int(seed) + k, **kwargs) for k in range(int(experiments))]
        ens = GraphEnsemble(sims, device=kwargs.get('device', 0),
                            history_capacity=max(2048, min(int(iterations), 1 << 16)))
        try:
            ens.initialize()
            agg = ens.run_aggregated(iterations)
        finally:
            ens.close()
        return list(GRAPH_STAT_KEYS), agg


def _entity_property(index):
    def getter(self):
        if self._ghandle is None:
            return [] if index in (1, 2) else None
        return self._materialise()[index]

    def setter(self, value):      # Simulation.__init__ / GraphSimulation.__init__ assign the empty defaults
        pass
    return property(getter, setter)


GraphSimulation.houses = _entity_property(1)
GraphSimulation.business = _entity_property(2)
GraphSimulation.government = _entity_property(3)
GraphSimulation.healthcare = _entity_property(4)
