"""Scenario x seed ensemble of GraphSimulation sharded over the GPUs of one box (BASELINE.json configs[4], SURVEY.md 8d
config 5): 7 paper scenarios x `seeds` seeds x `persons` persons.  One process per GPU (torchrun) or a single process;
simulations are dealt round-robin to the ranks (network/ensemble.py::shard, interleaved: the same scenario mix everywhere) and never communicate; at the end every
rank reduces its own Min/Sum/SumSq/Max per (scenario, iteration, metric) on the device and the ranks merge those few
numbers (what batch_experiment writes per scenario, experiments.py:200-208).

    python bench.py --workload ensemble [--gpus N] [--ensemble-seeds 1024] [--ensemble-persons 100000]
    python scripts/ensemble_bench.py --seeds 16 --persons 100000 --iterations 48

Every simulation has its OWN world (batch_experiment builds one per experiment, experiments.py:185-186), built by the
native host builder (cabs_graph_build_world) in a thread pool.
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# bytes a step has to move per executed person-hour: phase 1 reads x, y, attr, house, employer
# (20) and writes x, y, attr (12); the order-dependent stock pass re-reads one event word the first pass wrote (16);
# the contact pass reads x, y, attr of the moved persons (12) and writes + reads one 16-byte cell-list record per
# Susceptible person (~16 on average over an epidemic); the statistics read attr (4)
B_GRAPH_KERNEL = 80.0
B_GRAPH_SURVEY = 164.0   # the Simulation constant 147 + house, employer, incomes/expenses, flags (17)


def executed_hours(first, count, sleep=True):
    """Iterations of [first, first + count) that run the full hour (run_batch.py:8-13: hours 1..7 are slept)."""
    return sum(1 for it in range(first, first + count) if not (sleep and it % 24 != 0 and it % 24 < 8))


def run(args, rank, world, local):
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from covid19_agentbasedsimulation_b200 import build
    build.build_library()
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation
    from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble, shard
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    sys.path.insert(0, ROOT)
    from bench import ClockSampler, peaks

    sampler = ClockSampler(local)
    sampler.start()
    n, iters = int(args.ensemble_persons), int(args.ensemble_iterations)
    seeds = int(args.ensemble_seeds) or 128 * world      # default: 896 simulations per GPU (config 5: 1 024 seeds on 8)
    side = int((n / 0.0012) ** .5)                      # the density of run_batch.py: 300 persons on 500 x 500
    scenarios = sc.PAPER_SCENARIOS
    jobs = [(s, k) for s in range(len(scenarios)) for k in range(seeds)]
    # interleaved: every rank gets the same mix of scenarios (their seeds stay consecutive in its list)
    mine = [jobs[j] for j in shard(len(jobs), rank, world, interleaved=True)]

    def make(job):
        s, k = job
        kw = dict(sc.global_parameters)
        kw.update(scenarios[s])
        kw.pop("name")
        kw.update(population_size=n, length=side, height=side, total_wealth=1e7 * n / 300, device=local)
        sim = GraphSimulation(seed=10_000 * s + k, **kw)
        return sim, sim.build_world_fast(seed=(s << 32) | k)     # one world per simulation (experiments.py:185-186)

    t0 = time.perf_counter()
    threads = max(1, min(32, (os.cpu_count() or 8) // max(1, world)))
    with ThreadPoolExecutor(threads) as pool:
        made = list(pool.map(make, mine))
    t_build = time.perf_counter() - t0
    sims = [m[0] for m in made]
    worlds = [m[1] for m in made]

    warm = 24
    ens = GraphEnsemble(sims, device=local, history_capacity=max(256, iters + warm + 8))
    t0 = time.perf_counter()
    ens.initialize(worlds)
    ens.handle.sync()
    t_upload = time.perf_counter() - t0
    ens.run(warm)                                        # one full day of warm-up (also covers the first slept night)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    s0 = sampler.mark()
    first = ens.iteration + 1
    ens.handle.run(iters)
    ens.iteration += iters
    ms = ens.last_run_ms()                               # CUDA events on the handle's stream around the launch
    s1 = sampler.mark()
    barrier()
    hours = executed_hours(first, iters)
    # ---- what batch_experiment computes per scenario (experiments.py:200-208), reduced on the device: this rank's part
    #      of every scenario as {min, sum, sum of squares, max} per (iteration, metric); the parts are merged below
    t0 = time.perf_counter()
    mom = np.zeros((len(scenarios), iters, 14, 4))
    mom[..., 0] = np.inf
    mom[..., 3] = -np.inf
    for s in range(len(scenarios)):
        idx = [j for j, (sj, _) in enumerate(mine) if sj == s]
        if idx:                                          # the seeds of a scenario are consecutive simulations
            mom[s] = ens.handle.moments(idx[0], len(idx), first, iters)
    t_agg = time.perf_counter() - t0
    agg_bytes = int(mom.nbytes)
    # ---- end to end through the public classes: upload the worlds (host columns), run, read the aggregate back
    e2e_iters = min(iters, 24)
    t0 = time.perf_counter()
    ens2 = GraphEnsemble(sims, device=local, history_capacity=max(256, e2e_iters + 8))
    ens2.initialize(worlds)
    agg2 = ens2.run_aggregated(e2e_iters)
    e2e_s = time.perf_counter() - t0
    e2e_hours = executed_hours(0, e2e_iters)
    h2d = int(sum(v.nbytes for w, _, _ in worlds for v in w.values() if isinstance(v, np.ndarray)))
    ens2.close()
    clocks = sampler.stop(s0, s1)

    if world > 1:                                        # the only communication of the sweep: three small all-reduces
        t = torch.tensor(mom, device="cuda")
        mn, sm, mxm = t[..., 0].contiguous(), t[..., 1:3].contiguous(), t[..., 3].contiguous()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dist.all_reduce(mxm, op=dist.ReduceOp.MAX)
        mom = torch.cat([mn.unsqueeze(-1), sm, mxm.unsqueeze(-1)], dim=-1).cpu().numpy()
        tm = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tm[0].item()), float(tm[1].item())
    avg = mom[..., 1] / seeds                             # [scenario, iteration, metric]
    std = np.sqrt(np.maximum(mom[..., 2] / seeds - avg ** 2, 0.0))
    if rank == 0:
        peak, peak_kind = peaks()
        value = len(jobs) * n * hours / (ms * 1e-3)
        per_gpu = value / world
        line = {
            "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
            "steps": iters, "warmup": warm, "ms_per_step": ms / iters, "higher_is_better": True,
            "scaling": "weak" if not int(args.ensemble_seeds) else "strong",
            "vs_baseline": None, "dtype": "i32/u8/i64 fixed-point money", "data": "synthetic",
            "config": {"workload": "config5: GraphSimulation ensemble, 7 paper scenarios x {} seeds x {} persons "
                                   "(BASELINE.json configs[4]); a step = one hourly iteration of every simulation; "
                                   "agent-updates count executed hours only ({} of {}: hours 1..7 are slept)".format(
                                       seeds, n, hours, iters),
                       "n_sims": len(jobs), "sims_per_gpu": len(mine), "persons": n, "length": side,
                       "engine": "GraphSimulation (covid_abs/network/graph_abs.py), one world per simulation",
                       "multi_gpu": "whole simulations sharded over the ranks, no data-path communication",
                       "cache": "{:.0f} MB of person state per GPU >> 126 MB L2".format(60.0 * n * len(mine) / 1e6)},
            "clocks": clocks,
            "e2e": {"value": len(jobs) * n * e2e_hours / e2e_s, "unit": "agent-updates/s",
                    "what": "GraphEnsemble.initialize(worlds) [upload of every world from host columns] + "
                            "run_aggregated({}) [launch + device Min/Avg/Std/Max + readback], host wall clock".format(e2e_iters),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(agg2.nbytes), "seconds": e2e_s},
            "gpu_launches": 1,
            "roofline": {"bound": "hbm", "kernel": "k_graph_run", "unit": "GB/s", "peak": peak, "peak_source": peak_kind,
                         "alg_bytes_per_person_hour": B_GRAPH_KERNEL,
                         "achieved": per_gpu * B_GRAPH_KERNEL / 1e9, "frac": per_gpu * B_GRAPH_KERNEL / 1e9 / peak,
                         "traffic": None,
                         "whole_step": {"alg_bytes_per_agent_update": B_GRAPH_SURVEY,
                                        "achieved": per_gpu * B_GRAPH_SURVEY / 1e9,
                                        "frac": per_gpu * B_GRAPH_SURVEY / 1e9 / peak}},
            "cpu_baseline": None,
            "host": {"world_build_s": t_build, "world_build_threads": threads, "upload_s": t_upload,
                     "aggregate_readback_ms": t_agg * 1e3, "aggregate_bytes": agg_bytes},
            "final_infected_avg_std_by_scenario": [[float(avg[s, -1, 8]), float(std[s, -1, 8])]
                                                   for s in range(len(scenarios))],
            "final_death_avg_by_scenario": [float(avg[s, -1, 10]) for s in range(len(scenarios))],
        }
        print(json.dumps(line))
    ens.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--persons", type=int, default=100000)
    ap.add_argument("--iterations", type=int, default=48)
    a = ap.parse_args()
    args = argparse.Namespace(ensemble_persons=a.persons, ensemble_seeds=a.seeds, ensemble_iterations=a.iterations)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    run(args, rank, world, local)


if __name__ == "__main__":
    main()
