"""tests/golden/simulation_sequential.json: the UNMODIFIED reference `Simulation` stepped under the keyed draws of
oracle/sim_shim.py (TEST INFRASTRUCTURE ONLY).  Its sequential pair loop (abs.py:253-265) produces in-step infection
chains; the fixtures pin the sequential schedule of the oracle and of the GPU path.

    python oracle/gen_sequential_golden.py
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

from covid_abs.abs import Simulation  # noqa: E402
from covid_abs.agents import Status, InfectionSeverity  # noqa: E402
from oracle.sim_shim import KeyedSimulationRandom  # noqa: E402

STC = {Status.Susceptible: 0, Status.Infected: 1, Status.Recovered_Immune: 2, Status.Death: 3}
SEVC = {InfectionSeverity.Exposed: 0, InfectionSeverity.Asymptomatic: 1, InfectionSeverity.Hospitalization: 2,
        InfectionSeverity.Severe: 3}
KEYS = ("Susceptible", "Infected", "Recovered_Immune", "Death", "Asymptomatic", "Hospitalization", "Severe",
        "Q1", "Q2", "Q3", "Q4", "Q5")
CASES = [("defaults", 0, 500, 40, {}),
         ("dense_r2", 1, 501, 40, dict(population_size=80, length=30, height=30, contagion_distance=2.,
                                       initial_infected_perc=0.05, critical_limit=0.05)),
         ("rate1_r3", 2, 502, 25, dict(population_size=150, length=40, height=40, contagion_distance=3.,
                                       contagion_rate=1.0, initial_infected_perc=0.02)),
         ("legacy", 3, 503, 40, dict(initial_infected_perc=0.02, initial_immune_perc=0.01, length=100, height=100,
                                     population_size=80, contagion_distance=5., critical_limit=0.05))]


def main():
    out = []
    for name, np_seed, seed, iters, kw in CASES:
        np.random.seed(np_seed)
        sim = Simulation(**kw)
        sim.initialize()
        pop = sim.get_population()
        f = lambda v: float(np.asarray(v, dtype=np.float64).ravel()[0])  # noqa: E731
        init = dict(x=[int(a.x) for a in pop], y=[int(a.y) for a in pop], status=[STC[a.status] for a in pop],
                    age=[int(a.age) for a in pop], stratum=[int(a.social_stratum) for a in pop],
                    wealth=[f(a.wealth) for a in pop])
        sim.get_statistics()
        rows, new_per_step = [], []
        with KeyedSimulationRandom(sim, seed) as kr:
            for _ in range(iters):
                before = sum(1 for a in pop if a.status == Status.Infected)
                sim.execute()
                st = sim.get_statistics('all')
                kr.next_step()
                rows.append([float(st[k]) for k in KEYS])
        out.append(dict(name=name, seed=seed, iterations=iters, kwargs=kw, total_wealth=float(sim.total_wealth),
                        keys=list(KEYS), initial=init, rows=rows,
                        final=dict(x=[int(a.x) for a in pop], y=[int(a.y) for a in pop],
                                   status=[STC[a.status] for a in pop],
                                   severity=[SEVC[a.infected_status] for a in pop],
                                   infected_time=[int(a.infected_time) for a in pop],
                                   wealth=[f(a.wealth) for a in pop])))
        print(name, rows[-1][:4])
    with open(os.path.join(ROOT, "tests", "golden", "simulation_sequential.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
