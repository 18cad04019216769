"""Small GraphSimulation ensemble for ncu: 148 simulations x 20 000 persons, 48 hourly iterations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation  # noqa: E402
from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble  # noqa: E402
from covid19_agentbasedsimulation_b200.network import scenarios as sc  # noqa: E402

n, sims_n, iters = 20000, int(sys.argv[1]) if len(sys.argv) > 1 else 148, 48
kw = dict(sc.global_parameters)
kw.update(sc.scenario1)
kw.pop("name")
side = int((n / 0.0012) ** .5)
kw.update(population_size=n, length=side, height=side, total_wealth=1e7 * n / 300, initial_infected_perc=0.05)
np.random.seed(0)
proto = GraphSimulation(seed=0, **kw)
w = proto.build_world()
sims = [GraphSimulation(seed=100 + k, **kw) for k in range(sims_n)]
ens = GraphEnsemble(sims, history_capacity=2048)
ens.initialize(w)
ens.run(iters)       # warm-up, also lets the epidemic grow
ens.run(iters)
ms = ens.last_run_ms()
executed = sum(1 for it in range(iters, 2 * iters) if not (it % 24 != 0 and it % 24 < 8))
print("graph ensemble: %d sims x %d persons, %d iterations (%d executed): %.2f ms, %.3g agent-updates/s, %.1f us per executed hour"
      % (sims_n, n, iters, executed, ms, sims_n * n * executed / (ms * 1e-3), 1e3 * ms / executed))
