"""Feeds counter-based draws to the UNMODIFIED reference `Simulation` (TEST INFRASTRUCTURE ONLY).

Same idea as oracle/graph_shim.py: while the reference's `execute()` runs, `np.random.randn/rand/random` are
replaced by functions that look at the calling frame (move / update of which agent, contact of which pair) and
return the draw the counter-based map of oracle/philox.py assigns to that site:

    move    abs.py:174-175  np.random.randn(1) x2 -> the Box-Muller pair of the move stream's two words (float32, so
                            that the product with the amplitude is the binary32 product the device computes)
            abs.py:188      np.random.rand(1)     -> u01(economy stream word 0)
    update  abs.py:206,219  np.random.random() x2 -> u01(economy stream word 1), u01(death stream word 0)
    contact abs.py:150      np.random.random()    -> u01(pair stream word 0 of the unordered pair)

With it the reference -- its own sequential loop, including in-step infection chains -- must produce the trajectory
of `SyncSimulation(schedule="sequential")`, and the GPU's sequential schedule must equal both.
"""
import sys

import numpy as np

from oracle import philox


class KeyedSimulationRandom(object):
    def __init__(self, sim, seed):
        self.sim, self.seed = sim, int(seed)
        self.index = {id(a): k for k, a in enumerate(sim.population)}
        self.step = 0
        self.calls = {}
        self._saved = None

    def __enter__(self):
        self._saved = (np.random.randn, np.random.rand, np.random.random)
        np.random.randn, np.random.rand, np.random.random = self.randn, self.rand, self.random
        return self

    def __exit__(self, *exc):
        np.random.randn, np.random.rand, np.random.random = self._saved
        return False

    def next_step(self):
        self.step += 1
        self.calls = {}

    def _agent(self, frame):
        return self.index[id(frame.f_locals["agent"])]

    def randn(self, n):
        f = sys._getframe(1)
        assert f.f_code.co_name == "move"
        a = self._agent(f)
        w = philox.agent_draws(self.seed, a, self.step, philox.STREAM_MOVE)
        z = philox.normal_pair(w[0], w[1])
        k = self.calls.get(("n", a), 0)
        self.calls[("n", a)] = k + 1
        return np.array([z[k]], dtype=np.float32)

    def rand(self, n):
        f = sys._getframe(1)
        assert f.f_code.co_name == "move"
        w = philox.agent_draws(self.seed, self._agent(f), self.step, philox.STREAM_ECONOMY)
        return np.array([philox.u01(w[0])], dtype=np.float64)

    def random(self):
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name == "update":
            a = self._agent(f)
            k = self.calls.get(("u", a), 0)
            self.calls[("u", a)] = k + 1
            if k == 0:
                return float(philox.u01(philox.agent_draws(self.seed, a, self.step, philox.STREAM_ECONOMY)[1]))
            return float(philox.u01(philox.agent_draws(self.seed, a, self.step, philox.STREAM_DEATH)[0]))
        if name == "contact":
            a, b = self.index[id(f.f_locals["agent1"])], self.index[id(f.f_locals["agent2"])]
            lo, hi = min(a, b), max(a, b)
            return float(philox.u01(philox.pair_draws(self.seed, lo, hi, self.step)[0]))
        raise RuntimeError("unexpected np.random.random() in " + name)
