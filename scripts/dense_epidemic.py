"""Per-step kernel times of the Simulation step through an unmasked epidemic (r = 1, rate .9) at 10^7 agents."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covid19_agentbasedsimulation_b200.abs import Simulation  # noqa: E402

n, side = 10_000_000, 7071
rng = np.random.default_rng(0)
x = rng.integers(0, side + 1, size=n, dtype=np.int32)
y = rng.integers(0, side + 1, size=n, dtype=np.int32)
age = (rng.beta(2, 5, size=n) * 100).astype(np.uint8)
stratum = rng.integers(0, 5, size=n, dtype=np.uint8)
status = np.zeros(n, dtype=np.uint8)
status[:n // 100] = 1
sim = Simulation(seed=1, population_size=n, length=side, height=side, total_wealth=5e9, contagion_distance=1.0,
                 contagion_rate=0.9, history_capacity=256, max_cells=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
sim.set_population_arrays(x, y, status, age, stratum, np.full(n, 100.0))
h = sim._handle
h.profile_enable(True)
total = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 45):
    h.step(1)
    prof = h.profile_read()
    st = h.stats()[0]
    ms = sum(t for _, t in prof.values())
    total += ms
    if it % 3 == 2 or ms > 0.5:
        print("step %2d S %.2f I %.2f edge %d: %.3f ms  " % (it, st[0], st[1], h.info()["cell_edge"], ms) +
              "  ".join("%s %.3f" % (k.replace("k_", "")[:8], t / c) for k, (c, t) in prof.items()))
print("total %.1f ms for the run" % total)
