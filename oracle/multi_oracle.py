"""CPU restatement of MultiPopulationSimulation (covid_abs/abs.py:328-415) -- TEST INFRASTRUCTURE ONLY.

Nothing of the shipped package imports this module; it is the checker of `cabs_cross_contact` and of the host mirror
`covid19_agentbasedsimulation_b200.abs.MultiPopulationSimulation`.  Parity pinned: tests/golden/multi_population.json
holds trajectories of the UNMODIFIED reference class stepped under the keyed draws of oracle/multi_shim.py; the
`cross_schedule="sequential"` mode below must reproduce them exactly.

One step (abs.py:349-367): every population executes on its own (`SyncSimulation.execute`, abs.py:232-273), then for
every pair of populations m < n, for i in m, for j in n: if the distance of the shifted positions is <= the
MultiPopulation's own contagion_distance, `simulations[m].contact(ai, aj)` and `simulations[n].contact(aj, ai)`:
a Susceptible agent next to an Infected one is infected when the draw is <= the contagion_rate of ITS population.

cross_schedule:
  "sequential"  -- the reference's loop order with statuses mutating on the way (an agent infected at pair (i, j)
                   transmits along later pairs of the same double loop);
  "synchronous" -- only agents Infected when the double loop of a pair of populations starts transmit (what the
                   device computes; equal to "sequential" whenever no such chain occurs).
Draw of the encounter (a in m, b in n) at step t (1 after the first execute): Philox4x32-10 with counter
(id_a, t, 3 | (m * count + n) << 8, id_b), key = seed of the MultiPopulation; word 0 decides a <- b, word 1 b <- a.

Statistics (abs.py:389-410) keep the reference's quirk: each population overwrites the entry of the one before, so
every value is the LAST population's count over total_population, the Q sums are the last population's sums, and an
'Exposed' entry (always 0) sits between Death and Asymptomatic.
"""
import numpy as np

from oracle import philox
from oracle.sync_oracle import S, I, distance_threshold

STREAM_CROSS = 3
KEYS = ("Susceptible", "Infected", "Recovered_Immune", "Death", "Exposed", "Asymptomatic", "Hospitalization",
        "Severe", "Q1", "Q2", "Q3", "Q4", "Q5")


class MultiPopulationOracle(object):
    def __init__(self, simulations, positions, contagion_distance, seed, total_population=None,
                 cross_schedule="synchronous"):
        self.simulations = list(simulations)
        self.positions = [(int(px), int(py)) for px, py in positions]
        self.contagion_distance = float(contagion_distance)
        self.seed = int(seed)
        self.total_population = (sum(s.population_size for s in self.simulations) if total_population is None
                                 else total_population)
        assert cross_schedule in ("synchronous", "sequential")
        self.cross_schedule = cross_schedule
        self.step = 0
        self.cross_infections = 0

    def execute(self):
        for s in self.simulations:
            s.execute(need_pairs=False)
        self.step += 1
        count = len(self.simulations)
        for m in range(count):
            for n in range(m + 1, count):
                self._cross(m, n, m * count + n)

    def _cross(self, m, n, code):
        A, B = self.simulations[m], self.simulations[n]
        T = distance_threshold(self.contagion_distance)
        if T < 0:
            return
        xa, ya = A.x + self.positions[m][0], A.y + self.positions[m][1]
        xb, yb = B.x + self.positions[n][0], B.y + self.positions[n][1]
        d2 = (xa[:, None] - xb[None, :]) ** 2 + (ya[:, None] - yb[None, :]) ** 2
        ii, jj = np.nonzero(d2 <= T)                       # row-major: the reference's loop order
        if len(ii) == 0:
            return
        k0, k1 = philox.split_seed(self.seed)
        r = philox.philox4x32_10(ii.astype(np.uint64), self.step, STREAM_CROSS | (code << 8), jj.astype(np.uint64),
                                 k0, k1)
        ua, ub = philox.u01(r[0]), philox.u01(r[1])
        ra, rb = np.float64(A.contagion_rate), np.float64(B.contagion_rate)
        if self.cross_schedule == "synchronous":
            sa, sb = A.status.copy(), B.status.copy()
            a_hit = (sa[ii] == S) & (sb[jj] == I) & (ua <= ra)
            b_hit = (sb[jj] == S) & (sa[ii] == I) & (ub <= rb)
            before = int(np.sum(A.status == I) + np.sum(B.status == I))
            A.status[ii[a_hit]] = I
            B.status[jj[b_hit]] = I
            self.cross_infections += int(np.sum(A.status == I) + np.sum(B.status == I)) - before
        else:
            for p in range(len(ii)):
                i, j = int(ii[p]), int(jj[p])
                if A.status[i] == S and B.status[j] == I and ua[p] <= ra:     # simulations[m].contact(ai, aj)
                    A.status[i] = I
                    self.cross_infections += 1
                if B.status[j] == S and A.status[i] == I and ub[p] <= rb:     # simulations[n].contact(aj, ai)
                    B.status[j] = I
                    self.cross_infections += 1
        # a population's memo is recomputed at its next execute (abs.py:215 reads it there): after the cross contacts
        A.prev_stats = A.get_statistics()
        B.prev_stats = B.get_statistics()

    def get_statistics(self):
        last = self.simulations[-1]
        own = last.get_statistics()
        n = last.population_size
        out = {}
        for k in KEYS:
            if k == "Exposed":
                out[k] = np.float64(0.0) / self.total_population
            elif k.startswith("Q"):
                out[k] = own[k]
            else:
                out[k] = np.float64(round(float(own[k]) * n)) / self.total_population
        return out
