"""Timing of the GraphSimulation configs (BASELINE.json configs[1] and [4]-style) on one GPU."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation  # noqa: E402
from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble  # noqa: E402
from covid19_agentbasedsimulation_b200.network import scenarios as sc  # noqa: E402


def scenario_kwargs(s, **over):
    kw = dict(sc.global_parameters)
    kw.update(s)
    kw.pop("name")
    kw.update(over)
    return kw


def config1(iterations=2000):
    """covid_abs Simulation, 'do nothing', reference defaults (N=20 on 10x10): the reference's own CPU-runnable case
    (0.56 ms per iteration there).  Launch-bound on a GPU: 4 kernels per step."""
    from covid19_agentbasedsimulation_b200.abs import Simulation
    np.random.seed(0)
    sim = Simulation(seed=1, history_capacity=4096)
    sim.initialize()
    sim.run(50)
    h = sim._handle
    h.event_record(0)
    h.step(iterations)
    h.event_record(1)
    ms = h.event_elapsed_ms(0, 1)
    return dict(config="Simulation() reference defaults, 20 agents, {} iterations enqueued at once".format(iterations),
                ms=ms, us_per_iteration=1e3 * ms / iterations, agent_updates_per_s=20 * iterations / (ms * 1e-3))


def config2(iterations=1440):
    """GraphSimulation 1k people, Lockdown scenario (run_batch.py:103-112), 1 440 hourly iterations."""
    np.random.seed(0)
    sim = GraphSimulation(seed=1, history_capacity=2048, **scenario_kwargs(sc.scenario2, population_size=1000))
    t0 = time.perf_counter()
    sim.initialize()
    t_init = time.perf_counter() - t0
    sim.run(24)   # warm-up
    rows = sim.run(iterations)
    ms = sim.last_run_ms()
    executed = sum(1 for it in range(24, 24 + iterations) if not (it % 24 != 0 and it % 24 < 8))
    return dict(config="GraphSimulation 1000 persons, scenario2 (lockdown), {} hourly iterations in one launch".format(iterations),
                ms=ms, us_per_iteration=1e3 * ms / iterations, executed_iterations=executed,
                agent_updates_per_s=1000 * executed / (ms * 1e-3), host_initialize_s=t_init,
                final_infected=float(rows[-1][8]), final_death=float(rows[-1][10]))


def ensemble(n_persons=2000, seeds=64, iterations=240):
    """7 paper scenarios x `seeds` seeds x n_persons, one world per scenario shared by its seeds."""
    sims, worlds = [], []
    np.random.seed(0)
    for s in sc.PAPER_SCENARIOS:
        kw = scenario_kwargs(s, population_size=n_persons, length=int((n_persons / 0.0012) ** .5),
                             height=int((n_persons / 0.0012) ** .5), total_wealth=1e7 * n_persons / 300)
        proto = GraphSimulation(seed=0, **kw)
        w = proto.build_world()
        for k in range(seeds):
            sims.append(GraphSimulation(seed=1000 + k, **kw))
            worlds.append(w)
    ens = GraphEnsemble(sims, history_capacity=max(2048, iterations))
    ens.initialize(worlds)
    ens.run(24)
    table = ens.run(iterations)
    ms = ens.last_run_ms()
    executed = sum(1 for it in range(24, 24 + iterations) if not (it % 24 != 0 and it % 24 < 8))
    return dict(config="ensemble: 7 paper scenarios x {} seeds x {} persons, {} hourly iterations in one launch".format(
        seeds, n_persons, iterations), n_sims=len(sims), ms=ms,
        agent_updates_per_s=len(sims) * n_persons * executed / (ms * 1e-3),
        mean_final_infected_by_scenario=[float(table[k * seeds:(k + 1) * seeds, -1, 8].mean()) for k in range(7)])


if __name__ == "__main__":
    out = dict(config1=config1(), config2=config2(), ensemble_small=ensemble(2000, 32, 120))
    if "--big" in sys.argv:
        out["ensemble_big"] = ensemble(20000, 64, 120)
    print(json.dumps(out))
