"""256-seed ensembles of the UNMODIFIED reference GraphSimulation at N = 300 (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_graph_ensemble256.py [procs]     # writes tests/golden/ensemble_graph256.json (~1.5 h on 6 cores)

run_batch.py scenario1 / scenario2 / scenario3 (do nothing, lockdown, conditional lockdown; run_batch.py:87-123) with
the reference's own `global_parameters` (population_size = 300, 1 % infected, 1 % immune), `np.random.seed(s)` for
s = 0..255, 480 hourly iterations each (20 days: past the epidemic peak of scenario1).  Per iteration the mean and
standard deviation across seeds of the 14 statistics batch_experiment collects (experiments.py:193-196), plus the
per-seed rows at four sample iterations for the two-sample Kolmogorov-Smirnov test.  The GPU path draws from Philox,
so it is compared with these distributions, not with trajectories (tests/test_gpu_statistical.py).
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

SEEDS, ITERS = 256, 480
SAMPLES = (119, 239, 359, 479)
CASES = [("none", dict()), ("lockdown", dict()), ("conditional_lockdown", dict())]


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
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    path = os.path.join(ROOT, "tests", "golden", "ensemble_graph256.json")
    out = dict(stat_keys=list(GRAPH_STAT_KEYS), samples=list(SAMPLES), cases=[])
    with mp.Pool(procs) as pool:
        for scenario, over in CASES:
            res = pool.map(one, [(scenario, s, ITERS, over) for s in range(SEEDS)], chunksize=1)
            arr = np.array([r for r in res if r is not None])
            case = dict(scenario=scenario, seeds=int(arr.shape[0]), iterations=ITERS, overrides=over,
                        mean=arr.mean(axis=0).tolist(), std=arr.std(axis=0).tolist(),
                        rows_at={str(t): arr[:, t, :].tolist() for t in SAMPLES})
            out["cases"].append(case)
            print(scenario, arr.shape, "peak mean Infected", arr[:, :, 8].mean(axis=0).max(), flush=True)
            with open(path, "w") as f:     # partial results survive an interrupted run
                json.dump(out, f)


if __name__ == "__main__":
    main()
