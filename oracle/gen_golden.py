"""Generates tests/golden/*.json from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the authoring container, where /root/reference exists:

    python oracle/gen_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures:

  simulation_kat.json        known-answer vectors: `np.random.seed(s); Simulation(**kw).initialize();
                             execute() x k` -> full population state, contact pairs and statistics
                             (covid_abs/abs.py:116-139, 232-314)
  simulation_frozen.json     deterministic-schedule fixtures of SURVEY.md appendix C.3 (rate=1, amplitudes 0,
                             np.random.random patched to a constant)
  simulation_ensemble.json   per-iteration mean/std over many seeds of the statistics batch_experiment would
                             collect (experiments.py:193-196), for the statistical-parity tests
"""
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

from covid_abs.abs import Simulation  # noqa: E402
from covid_abs.agents import Status, InfectionSeverity  # noqa: E402

STC = {Status.Susceptible: 0, Status.Infected: 1, Status.Recovered_Immune: 2, Status.Death: 3}
SEVC = {InfectionSeverity.Exposed: 0, InfectionSeverity.Asymptomatic: 1, InfectionSeverity.Hospitalization: 2,
        InfectionSeverity.Severe: 3}

LEGACY = dict(initial_infected_perc=0.02, initial_immune_perc=0.01, length=100, height=100, population_size=80,
              contagion_distance=5., critical_limit=0.05)   # run_batch.py:547-568


def snapshot(sim):
    pop = sim.get_population()
    stats = sim.get_statistics('all')
    return dict(
        x=[int(a.x) for a in pop], y=[int(a.y) for a in pop],
        status=[STC[a.status] for a in pop], severity=[SEVC[a.infected_status] for a in pop],
        infected_time=[int(a.infected_time) for a in pop], age=[int(a.age) for a in pop],
        stratum=[int(a.social_stratum) for a in pop],
        wealth_hex=[float(np.asarray(a.wealth).ravel()[0]).hex() for a in pop],
        stats={k: float(v).hex() for k, v in stats.items()},
        stat_keys=list(stats.keys()))


def pair_list(sim):
    """The list `contacts` of abs.py:251-259, recomputed with the reference's own distance()."""
    from covid_abs.abs import distance
    pop = sim.get_population()
    out = []
    for i in range(sim.population_size):
        for j in range(i + 1, sim.population_size):
            if distance(pop[i], pop[j]) <= sim.contagion_distance:
                out.append([i, j])
    return out


def kat():
    cases = []
    for name, kw, seeds, snaps, steps in (("defaults", {}, (0, 1, 2), (1, 2, 10, 30, 60), 60),
                                          ("legacy_do_nothing", LEGACY, (0, 5), (1, 20, 40, 80), 80)):
        for seed in seeds:
            np.random.seed(seed)
            sim = Simulation(**kw)
            sim.initialize()
            case = dict(name=name, kwargs=kw, seed=seed, initial=snapshot(sim), after={})
            sim.statistics = None
            for it in range(1, steps + 1):
                sim.execute()
                if it in snaps:
                    snap = snapshot(sim)
                    snap["pairs"] = pair_list(sim)
                    case["after"][str(it)] = snap
                    sim.statistics = None  # keep the run identical to one that never asked for statistics
            cases.append(case)
    return cases


def frozen():
    """6 agents at (10+i, 10), age 30, stratum 2, wealth 100, rate 1, amplitudes 0."""
    from covid_abs.agents import Agent
    out = []
    orig = np.random.random
    np.random.random = lambda *a: 0.999999
    try:
        for seed_agent in (0, 2, 5):
            sim = Simulation(population_size=6, length=100, height=100, contagion_rate=1.0,
                             amplitudes={Status.Susceptible: 0, Status.Recovered_Immune: 0, Status.Infected: 0})
            pop = [Agent(x=np.int64(10 + i), y=np.int64(10), age=30, social_stratum=2, wealth=100.0,
                         status=Status.Infected if i == seed_agent else Status.Susceptible) for i in range(6)]
            sim.set_population(pop)
            steps = []
            for _ in range(2):
                sim.execute()
                steps.append(dict(status=''.join(a.status.value for a in sim.get_population()),
                                  infected_time=[int(a.infected_time) for a in sim.get_population()],
                                  wealth=[float(np.asarray(a.wealth).ravel()[0]) for a in sim.get_population()]))
            out.append(dict(seed_agent=seed_agent, steps=steps))
    finally:
        np.random.random = orig
    return out


def ensemble(kw, seeds, iterations):
    keys = None
    table = []
    for seed in seeds:
        np.random.seed(seed)
        sim = Simulation(**kw)
        sim.initialize()
        sim.get_statistics(kind='all')          # what batch_experiment does for experiment 0 (experiments.py:187-189)
        rows = []
        for _ in range(iterations):
            sim.execute()
            st = sim.get_statistics(kind='all')
            keys = list(st.keys())
            rows.append([float(st[k]) for k in keys])
        table.append(rows)
    arr = np.array(table)   # [seeds, iterations, metrics]
    return dict(kwargs=kw, seeds=[int(s) for s in seeds], iterations=iterations, keys=keys,
                mean=arr.mean(axis=0).tolist(), std=arr.std(axis=0).tolist(),
                # final-iteration samples for KS tests
                final=arr[:, -1, :].tolist(), mid=arr[:, iterations // 2, :].tolist())


def main():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "simulation_kat.json"), "w") as f:
        json.dump(kat(), f)
    with open(os.path.join(OUT, "simulation_frozen.json"), "w") as f:
        json.dump(frozen(), f)
    ens = dict(defaults=ensemble({}, range(256), 60), legacy_do_nothing=ensemble(LEGACY, range(256), 80))
    with open(os.path.join(OUT, "simulation_ensemble.json"), "w") as f:
        json.dump(ens, f)
    print("golden fixtures written to", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
