"""tests/golden/multi_population.json: the UNMODIFIED reference `MultiPopulationSimulation` (abs.py:328-415) stepped
under the keyed draws of oracle/multi_shim.py (TEST INFRASTRUCTURE ONLY).  Pins oracle/multi_oracle.py
(cross_schedule="sequential", populations on the sequential schedule) and, through it, `cabs_cross_contact`.

    python oracle/gen_multi_golden.py
"""
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("COVID_ABS_REFERENCE", "/root/reference"))

from covid_abs.abs import Simulation, MultiPopulationSimulation  # noqa: E402
from covid_abs.agents import Status, InfectionSeverity  # noqa: E402
from oracle.multi_shim import KeyedMultiRandom  # noqa: E402
from oracle.multi_oracle import KEYS  # noqa: E402

STC = {Status.Susceptible: 0, Status.Infected: 1, Status.Recovered_Immune: 2, Status.Death: 3}
SEVC = {InfectionSeverity.Exposed: 0, InfectionSeverity.Asymptomatic: 1, InfectionSeverity.Hospitalization: 2,
        InfectionSeverity.Severe: 3}


def amp(v):
    return {"s": v, "c": v, "i": v}     # by Status value (s, i, c = Susceptible, Infected, Recovered_Immune)


LEGACY = dict(initial_infected_perc=0.02, length=100, height=100, population_size=80, contagion_distance=5.,
              critical_limit=.1)
CASES = [
    # run_batch.py:645-678 ("scenario5" of the legacy article): a mobile and a nearly still population, corners touching
    dict(name="legacy_two_towns", np_seed=0, seed=900, seeds=[901, 902], iterations=80,
         sims=[dict(LEGACY, amplitudes=amp(5)), dict(LEGACY, amplitudes=amp(1))],
         positions=[(0, 0), (80, 80)], multi=dict(length=200, height=200, contagion_distance=5., critical_limit=.1)),
    # three overlapping populations with their own contagion rates: many cross contacts in both directions
    dict(name="three_overlapping", np_seed=1, seed=910, seeds=[911, 912, 913], iterations=40,
         sims=[dict(population_size=60, length=40, height=40, contagion_distance=2., contagion_rate=.9,
                    initial_infected_perc=0.05, amplitudes=amp(3)),
               dict(population_size=50, length=40, height=40, contagion_distance=2., contagion_rate=.5,
                    initial_infected_perc=0.0, amplitudes=amp(3)),
               dict(population_size=70, length=40, height=40, contagion_distance=2., contagion_rate=1.0,
                    initial_infected_perc=0.0, initial_immune_perc=0.1, amplitudes=amp(2))],
         positions=[(0, 0), (20, 10), (35, 35)], multi=dict(length=80, height=80, contagion_distance=3.)),
]


def enum_amp(d):
    return {Status(k): v for k, v in d.items()}


def snapshot(sim):
    pop = sim.get_population()
    f = lambda v: float(np.asarray(v, dtype=np.float64).ravel()[0])  # noqa: E731
    return dict(x=[int(a.x) for a in pop], y=[int(a.y) for a in pop], status=[STC[a.status] for a in pop],
                severity=[SEVC[a.infected_status] for a in pop], infected_time=[int(a.infected_time) for a in pop],
                age=[int(a.age) for a in pop], stratum=[int(a.social_stratum) for a in pop],
                wealth=[f(a.wealth) for a in pop])


def main():
    out = []
    for case in CASES:
        np.random.seed(case["np_seed"])
        sims = [Simulation(**dict(kw, amplitudes=enum_amp(kw["amplitudes"]))) for kw in case["sims"]]
        multi = MultiPopulationSimulation(simulations=sims, positions=list(case["positions"]),
                                          total_population=sum(s.population_size for s in sims), **case["multi"])
        multi.initialize()
        initial = [snapshot(s) for s in sims]
        for s in sims:
            s.get_statistics()
        rows, per_pop = [], []
        with KeyedMultiRandom(multi, case["seeds"], case["seed"]) as kr:
            for _ in range(case["iterations"]):
                multi.execute()
                st = multi.get_statistics('all')
                kr.next_step()
                rows.append([float(st[k]) for k in KEYS])
                per_pop.append([[sum(1 for a in s.get_population() if STC[a.status] == c) for c in range(4)]
                                for s in sims])
        assert list(multi.get_statistics('all').keys()) == list(KEYS)
        out.append(dict(name=case["name"], seed=case["seed"], seeds=case["seeds"], iterations=case["iterations"],
                        sims=case["sims"], total_wealth=[float(s.total_wealth) for s in sims],
                        positions=case["positions"], multi=case["multi"], keys=list(KEYS), initial=initial, rows=rows,
                        status_counts=per_pop, final=[snapshot(s) for s in sims]))
        print(case["name"], rows[-1][:4], per_pop[-1])
    with open(os.path.join(ROOT, "tests", "golden", "multi_population.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
