"""Ensemble curves of the UNMODIFIED reference GraphSimulation under its own numpy stream (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_graph_ensemble.py      # writes tests/golden/ensemble_graph.json (a few minutes, all cores)

For scenario families of run_batch.py, `seeds` runs of `iterations` hours each (np.random.seed(s), s = 0..seeds-1, the
reference's native draw order); per iteration the mean and standard deviation across seeds of the 14 statistics
batch_experiment collects (experiments.py:193-196).  Used by the statistical-parity tests of the GPU path, whose
draws are counter-based and therefore only statistically comparable with these.
"""
import json
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

CASES = [("none", 48, 264, dict(population_size=100, initial_infected_perc=0.05)),
         ("lockdown", 48, 264, dict(population_size=100, initial_infected_perc=0.05)),
         ("isolation", 48, 264, dict(population_size=100, initial_infected_perc=0.05))]


def one(args):
    scenario, seed, iters, over = args
    warnings.filterwarnings("ignore")
    from covid_abs.network.graph_abs import GraphSimulation
    from covid_abs.agents import Status
    from covid_abs.network.agents import EconomicalStatus
    import covid_abs.network.util as util
    from oracle.graph_shim import reference_simulation
    from oracle.graph_oracle import GRAPH_STAT_KEYS
    sim, _ = reference_simulation((GraphSimulation, Status, EconomicalStatus, util), scenario, seed, **over)
    rows = []
    for _ in range(iters):
        try:
            sim.execute()
        except ValueError:      # the reference's accounting can raise on empty lists; batch_experiment drops the run
            return None
        st = sim.get_statistics('all')
        rows.append([float(st[k]) for k in GRAPH_STAT_KEYS])
    return rows


def main():
    from oracle.graph_oracle import GRAPH_STAT_KEYS
    out = dict(stat_keys=list(GRAPH_STAT_KEYS), cases=[])
    with mp.Pool(os.cpu_count()) as pool:
        for scenario, seeds, iters, over in CASES:
            res = pool.map(one, [(scenario, s, iters, over) for s in range(seeds)])
            arr = np.array([r for r in res if r is not None])
            out["cases"].append(dict(scenario=scenario, seeds=int(arr.shape[0]), iterations=iters, overrides=over,
                                     mean=arr.mean(axis=0).tolist(), std=arr.std(axis=0).tolist()))
            print(scenario, arr.shape, "final Infected mean", arr[:, -1, 8].mean())
    with open(os.path.join(ROOT, "tests", "golden", "ensemble_graph.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
