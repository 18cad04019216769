"""Debug helper: one slab per process; dumps every rank's Python stack if it is stuck."""
import faulthandler
import os
import sys
import time

rank = int(os.environ["RANK"])
log = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "dbg_rank%d.log" % rank), "w")
faulthandler.dump_traceback_later(int(os.environ.get("DBG_AFTER", "70")), exit=True, file=log)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def say(*a):
    print(time.strftime("%H:%M:%S"), *a, file=log, flush=True)


import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

say("imports done")
world, local = int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
say("pg ready")
from covid19_agentbasedsimulation_b200.slab import SlabSimulation, owner_of  # noqa: E402
from oracle.sync_oracle import synthetic_population  # noqa: E402
transport = os.environ.get("SLAB_TRANSPORT", "collective")
n, side, steps = 200000, 1000, 10
pop, tw = synthetic_population(n, side, side, seed=4, infected=0.05, immune=0.01)
kw = dict(seed=21, population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=1.0,
          contagion_rate=0.9, device=local)
sim = SlabSimulation(distributed=True, slab_devices=[local], transport=transport, **kw)
mine = owner_of(pop["x"], side, world) == rank
ids = np.arange(n, dtype=np.uint32)
say("uploading", int(mine.sum()))
sim.set_population_arrays(pop["x"][mine], pop["y"][mine], pop["status"][mine], pop["age"][mine],
                          pop["stratum"][mine], pop["wealth"][mine], ids=ids[mine], population_size=n)
say("uploaded; syncing")
sim.sync()
say("static build done", sim.resident_counts())
for k in range(steps):
    sim._build(True)
    sim.sync()
    say("step", k, "done")
say("stats", sim.get_statistics('all')["Infected"])
if transport == "p2p":
    sim.close()
dist.destroy_process_group()
say("bye")
