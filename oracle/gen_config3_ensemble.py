"""256-seed ensembles of the UNMODIFIED reference `Simulation` at config 3's own parameters (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_config3_ensemble.py [procs]    # writes tests/golden/simulation_config3_ensemble.json (~10 min)

SURVEY.md 8d config 3 at a size the reference's O(N^2) loop can run: N = 300 at the same density (0.2 agents per unit
area -> 39 x 39), masks (contagion_distance .05, contagion_rate .1, run_batch.py:178-187), amplitudes 5, the conditional
lockdown triggers of run_batch.py:628-643 as the reference takes them (Python lambdas), 1 % immune and
  * `config3_masks`:        1 % infected -- the configuration itself;
  * `config3_masks_i20`:   20 % infected -- the same rules where infections actually happen every step, so that the
    comparison has statistical power (and the first trigger fires).
Per iteration the mean / standard deviation over np.random.seed(0..255) of the 12 statistics, plus the per-seed rows at
the middle and last iteration for the two-sample KS test (tests/test_gpu_statistical.py).
"""
import json
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

BASE = dict(population_size=300, length=39, height=39, initial_immune_perc=0.01, contagion_distance=0.05,
            contagion_rate=0.1, total_wealth=1e4 * 300 / 20, critical_limit=0.6)
CASES = {"config3_masks": dict(BASE, initial_infected_perc=0.01),
         "config3_masks_i20": dict(BASE, initial_infected_perc=0.20)}
SEEDS, ITERATIONS = 256, 60


def one(args):
    name, seed = args
    warnings.filterwarnings("ignore")
    from covid_abs.abs import Simulation
    from covid_abs.agents import Status
    amp = lambda v: {Status.Susceptible: v, Status.Recovered_Immune: v, Status.Infected: v}
    np.random.seed(seed)
    sim = Simulation(amplitudes=amp(5), triggers_simulation=[
        {'condition': lambda a: a.get_statistics()['Infected'] >= .2, 'attribute': 'amplitudes',
         'action': lambda a: amp(1.5)},
        {'condition': lambda a: a.get_statistics()['Infected'] <= .05, 'attribute': 'amplitudes',
         'action': lambda a: amp(5)}], **CASES[name])
    sim.initialize()
    sim.get_statistics(kind='all')          # what batch_experiment does for experiment 0 (experiments.py:187-189)
    rows, keys = [], None
    for _ in range(ITERATIONS):
        sim.execute()
        st = sim.get_statistics(kind='all')
        keys = list(st.keys())
        rows.append([float(st[k]) for k in keys])
    return keys, rows


def main():
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    out = {}
    with mp.Pool(procs) as pool:
        for name in CASES:
            res = pool.map(one, [(name, s) for s in range(SEEDS)], chunksize=4)
            arr = np.array([r[1] for r in res])
            out[name] = dict(kwargs=CASES[name], seeds=list(range(SEEDS)), iterations=ITERATIONS, keys=res[0][0],
                             mean=arr.mean(axis=0).tolist(), std=arr.std(axis=0).tolist(),
                             final=arr[:, -1, :].tolist(), mid=arr[:, ITERATIONS // 2, :].tolist())
            print(name, arr.shape, "final Infected mean", arr[:, -1, 1].mean(), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "simulation_config3_ensemble.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
