"""Run under torchrun on N GPUs: one slab per process over NCCL must reproduce the one-device run bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_slab_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)   # a stuck rank reports where instead of hanging the test
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from covid19_agentbasedsimulation_b200.slab import SlabSimulation, owner_of
    from oracle.sync_oracle import synthetic_population
    n, side, steps = 200000, 1000, 15
    pop, tw = synthetic_population(n, side, side, seed=4, infected=0.05, immune=0.01)
    kw = dict(seed=21, population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=1.0,
              contagion_rate=0.9, device=local)
    if os.environ.get("SLAB_CRIT"):   # a critical limit that is crossed during the run: abs.py:214-217 kills the severe
        kw["critical_limit"] = float(os.environ["SLAB_CRIT"])   # cases, on every rank from the SAME global statistics
        steps = 30
    transport = os.environ.get("SLAB_TRANSPORT", "collective")
    sim = SlabSimulation(distributed=True, slab_devices=[local], transport=transport, **kw)
    mine = owner_of(pop["x"], side, world) == rank
    ids = np.arange(n, dtype=np.uint32)
    sim.set_population_arrays(pop["x"][mine], pop["y"][mine], pop["status"][mine], pop["age"][mine],
                              pop["stratum"][mine], pop["wealth"][mine], ids=ids[mine], population_size=n)
    table = sim.run(steps)
    part = sim._download()
    ok = True
    if rank == 0:
        one = Simulation(**kw)
        one.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
        ref = one.run(steps)
        ok = bool(np.array_equal(ref, table))
        full = one._download()
        sel = part["id"]
        for k in ("x", "y", "status", "severity", "infected_time"):
            ok &= bool(np.array_equal(full[k][sel], part[k]))
        ok &= full["wealth"][sel].tobytes() == part["wealth"].tobytes()
    counts = [None] * world
    dist.all_gather_object(counts, int(len(part["id"])))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_SLAB_CHECK", transport, "OK" if int(flag.item()) == 1 and sum(counts) == n else "FAIL", counts)
    sim.close() if transport == "p2p" else None
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
