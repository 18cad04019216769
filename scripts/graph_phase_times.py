"""Where an executed hour of k_graph_run goes, by phase and hour class (timing-only build, -DCABS_GRAPH_TIMERS).

    scripts/with_variant.sh build/variants/graph_timers.so python scripts/graph_phase_times.py [--seeds 42] [--persons 100000]
"""
import argparse
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=42)
    ap.add_argument("--persons", type=int, default=100000)
    ap.add_argument("--days", type=str, default="1,10,25")
    ap.add_argument("--scenario", type=int, default=-1, help="only this paper scenario (0..6); default: all seven mixed")
    a = ap.parse_args()
    from covid19_agentbasedsimulation_b200 import _native
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation
    from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    lib = _native.load()
    have = hasattr(lib, "cabs_graph_phase_clocks")
    n = a.persons
    side = int((n / 0.0012) ** .5)
    jobs = [(s, k) for s in range(len(sc.PAPER_SCENARIOS)) for k in range(a.seeds) if a.scenario < 0 or s == a.scenario]
    if a.scenario >= 0:
        print("scenario %d: %s" % (a.scenario, sc.PAPER_SCENARIOS[a.scenario]["name"]))

    def make(job):
        s, k = job
        kw = dict(sc.global_parameters)
        kw.update(sc.PAPER_SCENARIOS[s])
        kw.pop("name")
        kw.update(population_size=n, length=side, height=side, total_wealth=1e7 * n / 300, device=0)
        sim = GraphSimulation(seed=10_000 * s + k, **kw)
        return sim, sim.build_world_fast(seed=(s << 32) | k)

    with ThreadPoolExecutor(min(32, os.cpu_count() or 8)) as pool:
        made = list(pool.map(make, jobs))
    ens = GraphEnsemble([m[0] for m in made], device=0, history_capacity=2048)
    ens.initialize([m[1] for m in made])
    buf = (C.c_ulonglong * 32)()

    def clocks():
        if not have:
            return None
        try:
            lib.cabs_graph_phase_clocks.argtypes = [C.POINTER(C.c_ulonglong)]
            lib.cabs_graph_phase_clocks(buf)
        except Exception:
            return None
        return np.array(list(buf), dtype=np.float64).reshape(4, 8)

    done = 0
    for day in [int(d) for d in a.days.split(",")]:
        skip = day * 24 - done
        if skip > 0:
            ens.handle.run(skip)
            done += skip
        clocks()
        ens.handle.run(48)
        done += 48
        ms = ens.last_run_ms()
        c = clocks()
        rows = ens.handle.stats_all(done - 1, 1)
        inf = float(np.mean(rows[:, 0, 8]))
        print("day %d: 48 iterations (34 executed) of %d sims x %d persons: %.2f ms = %.3g agent-updates/s; mean Infected %.4f"
              % (day, len(jobs), n, ms, len(jobs) * n * 34 / (ms * 1e-3), inf))
        if c is not None and c.sum() > 0:
            tot = c[:3].sum()
            print("   share of block-clocks   setup+p1 persons | p2 stocks | p3 accounts | p4 contacts | p5 statistics")
            for cls, name in enumerate(("hour 0 (new day)", "hours 8-16", "hours 17-23")):
                print("   %-18s %5.1f %%  | " % (name, 100 * c[cls].sum() / tot) +
                      " | ".join("%5.1f %%" % (100 * c[cls][k] / tot) for k in range(1, 6)))
            print("   all                         | " + " | ".join("%5.1f %%" % (100 * c[:, k].sum() / tot) for k in range(1, 6)))
            if c[:3, 6].sum() > 0:
                print("   phase 5 split: person loop (contacts + counts) %.1f %%, house loop %.1f %%, reductions + row %.1f %%"
                      % (100 * c[:3, 6].sum() / tot, 100 * c[:3, 7].sum() / tot, 100 * c[:3, 5].sum() / tot))
            ph = len(jobs) * n * 34.0
            k = c[3]
            if k[0] > 0:
                print("   contacts per person-hour: pulls %.3f, bitmap hits %.3f, items visited %.3f, draws %.4f, infections %.5f; "
                      "sources per sim-hour %.0f; longest walk of one thread %d items"
                      % (k[0] / ph, k[1] / ph, k[2] / ph, k[3] / ph, k[4] / ph, k[5] / (len(jobs) * 34.0), int(k[6])))
    ens.close()


if __name__ == "__main__":
    main()
