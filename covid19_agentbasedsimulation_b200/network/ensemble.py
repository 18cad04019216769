"""Scenario x seed ensembles of GraphSimulation in one launch.

`batch_experiment` (covid_abs/experiments.py:167-216) runs `experiments` independent simulations one after the
other; on the device they are one handle with `n_sims = experiments`: one thread block per simulation, all of them
stepping concurrently with no communication, and the `[simulation, iteration, 14]` statistics come back in one copy.
Across GPUs an ensemble is sharded by simulation index (`shard`), again with no communication until the final gather.
"""
import numpy as np

from covid19_agentbasedsimulation_b200 import _native
from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation, GRAPH_STAT_KEYS  # noqa: F401


def shard(n_items, rank, world, interleaved=False):
    """Item indices owned by `rank` of `world` (sizes differ by at most one): a contiguous block, or -- interleaved --
    every world-th item.  A scenario x seed sweep listed scenario-major should be interleaved: the cost of an hour
    depends on the scenario (a lockdown is twice the contact work of "do nothing" early on), and a block distribution
    hands whole scenarios to single GPUs (8 GPUs: 1.47e11 agent-updates/s against 8 x 2.45e10 on the mixed load of one)."""
    if interleaved:
        return range(int(rank), int(n_items), int(world))
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


class GraphEnsemble(object):
    """Many GraphSimulation objects sharing one device handle.

    sims: GraphSimulation instances (constructed, not initialised); every one keeps its own kwargs, scenario and seed.
    """

    def __init__(self, sims, device=0, history_capacity=2048):
        self.sims = list(sims)
        self.device, self.history_capacity = int(device), int(history_capacity)
        self.handle = None
        self.iteration = -1

    def initialize(self, worlds=None):
        """Build (graph_abs.py:89-188, in order, from np.random) or take the worlds, then upload all of them.

        worlds: optional list of (world, mode, sleep) per simulation, or one such tuple shared by all of them
        (a seed ensemble over one initial world)."""
        built = []
        for k, sim in enumerate(self.sims):
            if worlds is None:
                built.append(sim.build_world())
            elif isinstance(worlds, tuple):
                built.append(worlds)
            else:
                built.append(worlds[k])
        n = max(len(w["px"]) for w, _, _ in built)
        nh = max(len(w["hx"]) for w, _, _ in built)
        nb = max(len(w["bx"]) for w, _, _ in built)
        if self.handle is not None:
            self.handle.close()
        self.handle = _native.GraphHandle(len(self.sims), max(1, n), max(1, nh), nb, device=self.device,
                                          history_capacity=self.history_capacity)
        for k, (sim, (world, mode, sleep)) in enumerate(zip(self.sims, built)):
            sim.set_world(world, mode, sleep, handle=self.handle, index=k)
        self.iteration = -1

    def run(self, iterations):
        """All simulations advance `iterations` hours in one launch -> float64 [n_sims, iterations, 14]."""
        iterations = int(iterations)
        out = np.empty((len(self.sims), iterations, _native.GRAPH_NUM_STATS))
        done = 0
        while done < iterations:
            chunk = min(self.history_capacity, iterations - done)
            first = self.iteration + 1
            self.handle.run(chunk)
            self.iteration += chunk
            for sim in self.sims:
                sim.iteration = self.iteration
                sim._world = None
                sim._entities = None
                sim.statistics = None
            out[:, done:done + chunk] = self.handle.stats_all(first, chunk)
            done += chunk
        return out

    def run_aggregated(self, iterations):
        """`iterations` hours of every simulation, then what batch_experiment computes from their rows
        (experiments.py:200-208), reduced on the device: float64 [iterations, 14, 4] = Min, Avg, Std (ddof=0), Max.
        Only the aggregate crosses the bus (config 5: 0.6 MB instead of the 1.1 GB of per-simulation rows)."""
        iterations = int(iterations)
        out = np.empty((iterations, _native.GRAPH_NUM_STATS, 4))
        done = 0
        while done < iterations:
            chunk = min(self.history_capacity, iterations - done)
            first = self.iteration + 1
            self.handle.run(chunk)
            self.iteration += chunk
            out[done:done + chunk] = self.handle.aggregate(first, chunk)
            done += chunk
        for sim in self.sims:
            sim.iteration = self.iteration
            sim._world = None
            sim._entities = None
            sim.statistics = None
        return out

    def last_run_ms(self):
        return self.handle.last_run_ms()

    def close(self):
        if self.handle is not None:
            self.handle.close()
            self.handle = None
