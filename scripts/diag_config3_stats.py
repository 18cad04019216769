"""Prints the statistics behind tests/test_gpu_statistical.py::test_config3_parameters_... for both schedules."""
import json, os, sys
import numpy as np
from scipy import stats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from covid19_agentbasedsimulation_b200.abs import Simulation, Trigger
from covid19_agentbasedsimulation_b200.agents import Status
amp = lambda v: {Status.Susceptible: v, Status.Recovered_Immune: v, Status.Infected: v}
fix = json.load(open(os.path.join(ROOT, "tests", "golden", "simulation_config3_ensemble.json")))
OFFSETS = [None, 1000003, 2000003, 3000017]
for case, ens in fix.items():
    keys, iters, seeds = ens["keys"], ens["iterations"], len(ens["seeds"])
    for schedule, off in [(sc, o) for sc in ("synchronous", "sequential") for o in OFFSETS]:
        rows = np.empty((seeds, iters, len(keys)))
        for s in range(seeds):
            np.random.seed(s)
            sim = Simulation(schedule=schedule, seed=None if off is None else off + s, amplitudes=amp(5), triggers_simulation=[
                Trigger('Infected', '>=', .2, 'amplitudes', amp(1.5)), Trigger('Infected', '<=', .05, 'amplitudes', amp(5))],
                **ens["kwargs"])
            sim.initialize()
            rows[s] = sim.run(iters)
        mean_ref, std_ref = np.array(ens["mean"]), np.array(ens["std"])
        se = np.sqrt(std_ref ** 2 / seeds + rows.std(axis=0) ** 2 / seeds) + 1e-12
        z = np.abs(rows.mean(axis=0) - mean_ref) / se
        live = std_ref > 0
        ps = []
        for when, sample in (("mid", iters // 2), ("final", iters - 1)):
            ref = np.array(ens[when])
            for k, key in enumerate(keys[:7]):
                if np.ptp(ref[:, k]) == 0 and np.ptp(rows[:, sample, k]) == 0:
                    continue
                ps.append((round(float(stats.ks_2samp(rows[:, sample, k], ref[:, k]).pvalue), 4), when, key))
        print(case, schedule, off, "inside1.96 %.3f zmax %.2f" % (np.mean(z[live] < 1.96), z[live].max()), sorted(ps)[:6],
              "means final I/R ours %.4f/%.4f ref %.4f/%.4f" % (rows[:, -1, 1].mean(), rows[:, -1, 2].mean(), mean_ref[-1, 1], mean_ref[-1, 2]))
