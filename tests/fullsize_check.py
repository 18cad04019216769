"""Debug aid: the 10^7-agent population of tests/test_gpu_fullsize.py, device run against the CPU restatement, agent by agent.

    python tests/fullsize_check.py [--lib PATH] [--dirty] [--steps 2] [--max-cells 0]

`--dirty` fills and frees a few GB of device memory first, so that fresh allocations do not come back zeroed.
The CPU rows and populations are cached in /tmp so that several invocations (different libraries / modes) share them.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cpu_side(n, side, steps, pop):
    path = "/tmp/fullsize_cpu_{}_{}.npz".format(n, steps)
    if os.path.exists(path):
        return dict(np.load(path))
    from oracle.sync_oracle import SyncSimulation
    x, y, status, age, stratum, wealth = pop
    s = SyncSimulation(seed=5, population_size=n, length=side, height=side, total_wealth=5e9, contagion_distance=1.0,
                       contagion_rate=0.9)
    s.set_population(x, y, status, age, stratum, wealth)
    out = {}
    rows = []
    for it in range(steps):
        s.execute(need_pairs=False)
        st = s.get_statistics()
        rows.append([st[k] for k in st])
        for k in ("x", "y", "status", "severity", "infected_time", "wealth"):
            out["{}_{}".format(k, it)] = np.array(getattr(s, k))
    out["rows"] = np.array(rows)
    np.savez(path, **out)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--dirty", action="store_true")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--max-cells", type=int, default=0)
    ap.add_argument("--one-call", action="store_true", help="enqueue all steps with one run() call")
    args = ap.parse_args()
    from covid19_agentbasedsimulation_b200 import _native
    if args.lib:
        _native.LIB_PATH = os.path.abspath(args.lib)
    import torch
    if args.dirty:
        junk = [torch.full((1 << 28,), -1, dtype=torch.int32, device="cuda") for _ in range(6)]
        torch.cuda.synchronize()
        del junk
        torch.cuda.empty_cache()
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from test_gpu_fullsize import _population
    n = args.n
    side = int(round((n / 0.2) ** 0.5))
    pop = _population(n, side, 1)
    t0 = time.time()
    cpu = cpu_side(n, side, args.steps, pop)
    print("cpu side: {:.1f} s".format(time.time() - t0), flush=True)
    sim = Simulation(seed=5, population_size=n, length=side, height=side, total_wealth=5e9, max_cells=args.max_cells,
                     contagion_distance=1.0, contagion_rate=0.9)
    sim.set_population_arrays(*pop)
    tag = "lib={} dirty={} max_cells={} one_call={}".format(os.path.basename(args.lib or "default"), args.dirty,
                                                          args.max_cells, args.one_call)
    bad = 0
    if args.one_call:
        rows = sim.run(args.steps)
        its = [args.steps - 1]
    else:
        rows = np.concatenate([sim.run(1) for _ in range(1)])
        its = [0]
    it = its[0]
    for step_i in range(len(rows)):
        same = np.array_equal(rows[step_i], cpu["rows"][step_i])
        print(tag, "row", step_i, "equal" if same else "DIFFERENT")
        if not same:
            bad += 1
            print("  device:", rows[step_i][:7])
            print("  cpu   :", cpu["rows"][step_i][:7])
    h = sim._download()
    order = np.argsort(h["id"], kind="stable")
    ids_ok = np.array_equal(h["id"][order], np.arange(n, dtype=h["id"].dtype))
    print(tag, "ids are a permutation:", ids_ok)
    if not ids_ok:
        u, c = np.unique(h["id"], return_counts=True)
        print("  distinct ids", len(u), "duplicated", int((c > 1).sum()), "max multiplicity", int(c.max()))
        bad += 1
    for k in ("x", "y", "status", "severity", "infected_time", "wealth"):
        dev = h[k][order]
        ref = cpu["{}_{}".format(k, it)]
        if k == "wealth":
            ne = dev.view(np.uint64) != ref.view(np.uint64)
        else:
            ne = dev.astype(np.int64) != ref.astype(np.int64)
        cnt = int(ne.sum())
        print(tag, "after step", it + 1, k, "mismatches", cnt)
        if cnt:
            bad += 1
            w = np.nonzero(ne)[0][:8]
            print("   ids", w.tolist(), "device", dev[w].tolist(), "cpu", ref[w].tolist())
    if not args.one_call and args.steps > 1:
        more = sim.run(args.steps - 1)
        for j, r in enumerate(more):
            same = np.array_equal(r, cpu["rows"][1 + j])
            print(tag, "row", 1 + j, "equal" if same else "DIFFERENT")
            bad += 0 if same else 1
    print(tag, "RESULT", "OK" if bad == 0 else "BAD")


if __name__ == "__main__":
    main()
