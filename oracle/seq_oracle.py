"""Sequential CPU restatement of `covid_abs.abs.Simulation` (TEST INFRASTRUCTURE ONLY).

Follows the reference line by line but holds the population as structure of
arrays and takes its random draws from a pluggable source, so that

* with `LegacyStream` (the numpy global legacy RandomState, consumed in the
  reference's own call order) it reproduces the unmodified reference bit for
  bit -- this is what pins the oracle (tests/test_oracle_vs_reference.py and
  tests/golden/simulation_kat.json);
* with `FrozenDraws` it gives the deterministic-schedule fixtures of
  SURVEY.md appendix C.3.

Reference lines restated here (all under /root/reference/covid_abs/):
  abs.py:51-55     _xclip/_yclip
  abs.py:97-101    random_position
  abs.py:103-114   create_agent
  abs.py:116-139   initialize
  abs.py:141-154   contact
  abs.py:156-189   move
  abs.py:191-230   update
  abs.py:232-273   execute
  abs.py:291-322   get_statistics / filter_stats
  agents.py:9-26   Status / InfectionSeverity encodings
  common.py:13-27  probability tables, lorenz curve, basic_income
"""
import numpy as np

# agents.py:13-16 -- enum declaration order defines the statistics key order
S, I, R, D = 0, 1, 2, 3
STATUS_NAMES = ("Susceptible", "Infected", "Recovered_Immune", "Death")
# agents.py:23-26
SEV_EXPOSED, SEV_ASYM, SEV_HOSP, SEV_SEVERE = 0, 1, 2, 3
SEVERITY_NAMES = ("Exposed", "Asymptomatic", "Hospitalization", "Severe")

# common.py:13-15
AGE_HOSP = [0.001, 0.003, 0.012, 0.032, 0.049, 0.102, 0.166, 0.243, 0.273]
AGE_SEVERE = [0.05, 0.05, 0.05, 0.05, 0.063, 0.122, 0.274, 0.432, 0.709]
AGE_DEATH = [0.002, 0.00006, 0.0003, 0.0008, 0.0015, 0.006, 0.022, 0.051, 0.093]
# common.py:25-27
LORENZ = [.04, .08, .13, .2, .55]
BASIC_INCOME = np.array(LORENZ) / np.min(LORENZ)

STAT_KEYS = ("Susceptible", "Infected", "Recovered_Immune", "Death",
             "Asymptomatic", "Hospitalization", "Severe",
             "Q1", "Q2", "Q3", "Q4", "Q5")


def age_index(age):
    """abs.py:204 -- `age // 10 - 1 if age > 10 else 0` (bins 0-19, 20-29, ...)."""
    return age // 10 - 1 if age > 10 else 0


class LegacyStream(object):
    """Draws from numpy's global legacy RandomState with the reference's own calls.

    abs.py:98-99,174-175 use `np.random.randn(1)`, abs.py:113,188 `np.random.rand(1)`,
    abs.py:112 `np.random.beta(2, 5, 1)`, abs.py:150,206,219 `np.random.random()`.
    """

    def randn(self):
        return np.random.randn(1)[0]

    def rand(self):
        return np.random.rand(1)[0]

    def random(self):
        return np.random.random()

    def beta(self, a, b):
        return np.random.beta(a, b, 1)[0]


class FrozenDraws(object):
    """Constant draws (SURVEY.md appendix C.3 patches `np.random.random`)."""

    def __init__(self, randn=0.0, rand=0.5, random=0.999999, beta=0.3):
        self._randn, self._rand, self._random, self._beta = randn, rand, random, beta

    def randn(self):
        return self._randn

    def rand(self):
        return self._rand

    def random(self):
        return self._random

    def beta(self, a, b):
        return self._beta


class SeqSimulation(object):
    """SoA restatement of `covid_abs.abs.Simulation`, sequential schedule."""

    def __init__(self, draws=None, **kwargs):
        # abs.py:17-49
        self.population_size = kwargs.get("population_size", 20)
        self.length = kwargs.get("length", 10)
        self.height = kwargs.get("height", 10)
        self.initial_infected_perc = kwargs.get("initial_infected_perc", 0.05)
        self.initial_immune_perc = kwargs.get("initial_immune_perc", 0.05)
        self.contagion_distance = kwargs.get("contagion_distance", 1.)
        self.contagion_rate = kwargs.get("contagion_rate", 0.9)
        self.critical_limit = kwargs.get("critical_limit", 0.6)
        # amplitudes indexed by status code; Death never moves (abs.py:164)
        self.amplitudes = dict(kwargs.get("amplitudes", {S: 5, R: 5, I: 5}))
        self.minimum_income = kwargs.get("minimum_income", 1.0)
        self.minimum_expense = kwargs.get("minimum_expense", 1.0)
        self.statistics = None
        self.triggers_simulation = kwargs.get("triggers_simulation", [])
        self.total_wealth = kwargs.get("total_wealth", 10 ** 4)
        self.draws = draws if draws is not None else LegacyStream()
        n = 0
        self.x = np.zeros(n, dtype=np.int64)
        self.y = np.zeros(n, dtype=np.int64)
        self.status = np.zeros(n, dtype=np.int64)
        self.severity = np.zeros(n, dtype=np.int64)
        self.infected_time = np.zeros(n, dtype=np.int64)
        self.age = np.zeros(n, dtype=np.int64)
        self.stratum = np.zeros(n, dtype=np.int64)
        self.wealth = np.zeros(n, dtype=np.float64)
        self.last_contacts = []

    # ------------------------------------------------------------------ init
    def set_population(self, x, y, status, age, stratum, wealth, severity=None, infected_time=None):
        n = len(x)
        self.population_size = n
        self.x = np.array(x, dtype=np.int64)
        self.y = np.array(y, dtype=np.int64)
        self.status = np.array(status, dtype=np.int64)
        self.age = np.array(age, dtype=np.int64)
        self.stratum = np.array(stratum, dtype=np.int64)
        self.wealth = np.array(wealth, dtype=np.float64)
        self.severity = (np.full(n, SEV_ASYM, dtype=np.int64) if severity is None
                         else np.array(severity, dtype=np.int64))
        self.infected_time = (np.zeros(n, dtype=np.int64) if infected_time is None
                              else np.array(infected_time, dtype=np.int64))
        self.statistics = None

    def _random_position(self):
        # abs.py:97-101 with the clips of abs.py:51-55
        x = int(np.clip(int(self.length / 2 + (self.draws.randn() * (self.length / 3))), 0, self.length))
        y = int(np.clip(int(self.height / 2 + (self.draws.randn() * (self.height / 3))), 0, self.height))
        return x, y

    def initialize(self):
        # abs.py:116-139; create_agent = abs.py:103-114
        xs, ys, st, ages, strata = [], [], [], [], []

        def create(status):
            x, y = self._random_position()
            age = int(self.draws.beta(2, 5) * 100)
            stratum = int(self.draws.rand() * 100 // 20)
            xs.append(x), ys.append(y), st.append(status), ages.append(age), strata.append(stratum)

        for _ in range(int(self.population_size * self.initial_infected_perc)):
            create(I)
        for _ in range(int(self.population_size * self.initial_immune_perc)):
            create(R)
        for _ in range(self.population_size - len(xs)):
            create(S)
        n = len(xs)
        wealth = np.zeros(n, dtype=np.float64)
        ages_a, strata_a = np.array(ages), np.array(strata)
        for q in range(5):
            total = LORENZ[q] * self.total_wealth
            sel = (strata_a == q) & (ages_a >= 18)
            qty = max(1.0, float(np.sum(sel)))
            wealth[sel] = total / qty
        self.set_population(xs, ys, st, ages, strata, wealth)
        self.population_size = n

    # ------------------------------------------------------------------ step
    def _move(self, k):
        # abs.py:156-189 (population 'move' triggers are not restated: out of scope rows)
        st = self.status[k]
        if st == D or (st == I and (self.severity[k] == SEV_HOSP or self.severity[k] == SEV_SEVERE)):
            return
        ix = int(self.draws.randn() * self.amplitudes[int(st)])
        iy = int(self.draws.randn() * self.amplitudes[int(st)])
        if (self.x[k] + ix) <= 0 or (self.x[k] + ix) >= self.length:
            self.x[k] -= ix
        else:
            self.x[k] += ix
        if (self.y[k] + iy) <= 0 or (self.y[k] + iy) >= self.height:
            self.y[k] -= iy
        else:
            self.y[k] += iy
        dist = np.sqrt(ix ** 2 + iy ** 2)
        result_ecom = self.draws.rand()
        self.wealth[k] += dist * result_ecom * self.minimum_expense * BASIC_INCOME[self.stratum[k]]

    def _update(self, k):
        # abs.py:191-230
        if self.status[k] == D:
            return
        if self.status[k] == I:
            self.infected_time[k] += 1
            idx = age_index(int(self.age[k]))
            test = self.draws.random()
            if self.severity[k] == SEV_ASYM:
                if AGE_HOSP[idx] > test:
                    self.severity[k] = SEV_HOSP
            elif self.severity[k] == SEV_HOSP:
                if AGE_SEVERE[idx] > test:
                    self.severity[k] = SEV_SEVERE
                    self.get_statistics()
                    if self.statistics['Severe'] + self.statistics['Hospitalization'] >= self.critical_limit:
                        self.status[k] = D
                        self.severity[k] = SEV_ASYM
            death_test = self.draws.random()
            if AGE_DEATH[idx] > death_test:
                self.status[k] = D
                self.severity[k] = SEV_ASYM
                return
            if self.infected_time[k] > 20:
                self.infected_time[k] = 0
                self.status[k] = R
                self.severity[k] = SEV_ASYM
        self.wealth[k] -= self.minimum_expense * BASIC_INCOME[self.stratum[k]]

    def contact_pairs(self):
        """abs.py:253-259 -- all i<j with sqrt(dx^2+dy^2) <= contagion_distance, lexicographic."""
        n = self.population_size
        pairs = []
        for i in range(n - 1):
            dx = self.x[i] - self.x[i + 1:]
            dy = self.y[i] - self.y[i + 1:]
            hit = np.nonzero(np.sqrt(dx ** 2 + dy ** 2) <= self.contagion_distance)[0]
            pairs.extend((i, int(i + 1 + j)) for j in hit)
        return pairs

    def _contact(self, a, b):
        # abs.py:141-154 (the severity writes go to a misspelt attribute: no effect, quirk A.7-1)
        if self.status[a] == S and self.status[b] == I:
            test = self.draws.random()
            if test <= self.contagion_rate:
                self.status[a] = I

    def execute(self):
        # abs.py:232-273
        for k in range(self.population_size):
            self._move(k)
            self._update(k)
        contacts = self.contact_pairs()
        self.last_contacts = contacts
        for (i, j) in contacts:
            self._contact(i, j)
            self._contact(j, i)
        for trig in self.triggers_simulation:
            if trig['condition'](self):
                attr = trig['attribute']
                self.__dict__[attr] = trig['action'](self.__dict__[attr])
        self.statistics = None

    # ------------------------------------------------------------ statistics
    def get_statistics(self, kind='info'):
        # abs.py:291-314; memoised until the end of execute (abs.py:273,298)
        if self.statistics is None:
            n = self.population_size
            stats = {}
            for code, name in enumerate(STATUS_NAMES):
                stats[name] = np.sum(self.status == code) / n
            alive = self.status != D
            for code in (SEV_ASYM, SEV_HOSP, SEV_SEVERE):
                stats[SEVERITY_NAMES[code]] = np.sum((self.severity == code) & alive) / n
            for q in range(5):
                sel = (self.stratum == q) & (self.age >= 18) & alive
                stats['Q{}'.format(q + 1)] = np.sum(self.wealth[sel])
            self.statistics = stats
        return self.filter_stats(kind)

    def filter_stats(self, kind):
        # abs.py:316-322
        if kind == 'info':
            return {k: v for k, v in self.statistics.items()
                    if not k.startswith('Q') and k not in ('Business', 'Government')}
        elif kind == 'ecom':
            return {k: v for k, v in self.statistics.items()
                    if k.startswith('Q') or k in ('Business', 'Government')}
        return self.statistics

    def status_string(self):
        return ''.join('sicm'[int(s)] for s in self.status)
