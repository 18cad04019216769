"""Scenario x seed ensemble of GraphSimulation sharded over the GPUs of one box (BASELINE.json configs[4]):
7 paper scenarios x `--seeds` seeds x `--persons` persons.  One process per GPU (torchrun) or a single process;
simulations are block-distributed over the ranks (network/ensemble.py::shard), no communication until the final
gather of the [simulation, iteration, 14] statistics.

    python scripts/ensemble_bench.py --seeds 16 --persons 100000 --iterations 240
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/ensemble_bench.py --seeds 1024
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--persons", type=int, default=100000)
    ap.add_argument("--iterations", type=int, default=240)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation
    from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble, shard
    from covid19_agentbasedsimulation_b200.network import scenarios as sc

    n = args.persons
    side = int((n / 0.0012) ** .5)                      # the density of run_batch.py: 300 persons on 500 x 500
    jobs = [(s, k) for s in range(len(sc.PAPER_SCENARIOS)) for k in range(args.seeds)]
    mine = [jobs[j] for j in shard(len(jobs), rank, world)]
    t0 = time.perf_counter()
    worlds, sims, built = {}, [], []
    for s, k in mine:
        kw = dict(sc.global_parameters)
        kw.update(sc.PAPER_SCENARIOS[s])
        kw.pop("name")
        kw.update(population_size=n, length=side, height=side, total_wealth=1e7 * n / 300, device=local)
        if s not in worlds:                              # one initial world per scenario, shared by its seeds
            np.random.seed(1000 + s)
            worlds[s] = GraphSimulation(seed=0, **kw).build_world()
        sims.append(GraphSimulation(seed=10_000 * s + k, **kw))
        built.append(worlds[s])
    t_build = time.perf_counter() - t0
    ens = GraphEnsemble(sims, device=local, history_capacity=max(2048, args.iterations))
    ens.initialize(built)
    table = ens.run(args.iterations)                    # [my simulations, iterations, 14]
    ms = ens.last_run_ms()
    # what crosses the bus afterwards: every row of every simulation, or the Min/Avg/Std/Max reduced on the device
    t0 = time.perf_counter()
    rows_again = ens.handle.stats_all(0, args.iterations)
    t_rows = time.perf_counter() - t0
    t0 = time.perf_counter()
    agg = ens.handle.aggregate(0, args.iterations)
    t_agg = time.perf_counter() - t0
    assert np.array_equal(rows_again, table) and np.array_equal(agg[:, :, 3], table.max(axis=0))
    executed = sum(1 for it in range(args.iterations) if not (it % 24 != 0 and it % 24 < 8))
    mine_rate = len(sims) * n * executed / (ms * 1e-3)
    final = np.zeros((len(sc.PAPER_SCENARIOS), 3))      # per scenario: sum of final Infected, Death, count
    for (s, k), rows in zip(mine, table):
        final[s] += (rows[-1, 8], rows[-1, 10], 1.0)
    if world > 1:
        t = torch.tensor(np.concatenate([final.ravel(), [ms]]), device="cuda")
        mx = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        final = t[:-1].cpu().numpy().reshape(final.shape)
        ms = float(mx[-1].item())
    if rank == 0:
        total = len(jobs) * n * executed / (ms * 1e-3)
        print(json.dumps(dict(
            config="ensemble: 7 paper scenarios x {} seeds x {} persons, {} hourly iterations, {} GPU(s), one block per "
                   "simulation, no communication".format(args.seeds, n, args.iterations, world),
            n_sims=len(jobs), sims_per_gpu=len(sims), ms=ms, agent_updates_per_s=total, per_gpu_first_rank=mine_rate,
            host_world_build_s=t_build,
            readback_first_rank=dict(rows_ms=t_rows * 1e3, rows_bytes=int(rows_again.nbytes),
                                     aggregate_ms=t_agg * 1e3, aggregate_bytes=int(agg.nbytes)),
            mean_final_infected=[float(final[s, 0] / max(1.0, final[s, 2])) for s in range(final.shape[0])],
            mean_final_death=[float(final[s, 1] / max(1.0, final[s, 2])) for s in range(final.shape[0])])))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
