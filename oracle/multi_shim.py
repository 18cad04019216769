"""Feeds counter-based draws to the UNMODIFIED reference `MultiPopulationSimulation` (TEST INFRASTRUCTURE ONLY).

Extension of oracle/sim_shim.py to several populations: the agent found in the calling frame tells which population
(and so which seed) a move / update / contact draw belongs to; a `contact` called from
`MultiPopulationSimulation.execute` (abs.py:365-366) is a cross contact and gets the draw of
oracle/multi_oracle.py: counter (id in m, step, 3 | code << 8, id in n), word 0 when the target is in m, word 1 when it
is in n.
"""
import sys

import numpy as np

from oracle import philox
from oracle.multi_oracle import STREAM_CROSS


class KeyedMultiRandom(object):
    def __init__(self, multi, seeds, seed):
        self.multi, self.seeds, self.seed = multi, [int(s) for s in seeds], int(seed)
        self.where = {}
        for m, sim in enumerate(multi.simulations):
            for k, a in enumerate(sim.population):
                self.where[id(a)] = (m, k)
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

    def randn(self, n):
        f = sys._getframe(1)
        assert f.f_code.co_name == "move"
        m, a = self.where[id(f.f_locals["agent"])]
        w = philox.agent_draws(self.seeds[m], a, self.step, philox.STREAM_MOVE)
        z = philox.normal_pair(w[0], w[1])
        k = self.calls.get(("n", m, a), 0)
        self.calls[("n", m, a)] = k + 1
        return np.array([z[k]], dtype=np.float32)

    def rand(self, n):
        f = sys._getframe(1)
        assert f.f_code.co_name == "move"
        m, a = self.where[id(f.f_locals["agent"])]
        w = philox.agent_draws(self.seeds[m], a, self.step, philox.STREAM_ECONOMY)
        return np.array([philox.u01(w[0])], dtype=np.float64)

    def random(self):
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name == "update":
            m, a = self.where[id(f.f_locals["agent"])]
            k = self.calls.get(("u", m, a), 0)
            self.calls[("u", m, a)] = k + 1
            if k == 0:
                return float(philox.u01(philox.agent_draws(self.seeds[m], a, self.step, philox.STREAM_ECONOMY)[1]))
            return float(philox.u01(philox.agent_draws(self.seeds[m], a, self.step, philox.STREAM_DEATH)[0]))
        if name == "contact":
            (m1, a), (m2, b) = self.where[id(f.f_locals["agent1"])], self.where[id(f.f_locals["agent2"])]
            if m1 == m2:        # inside one population (abs.py:261-265)
                lo, hi = min(a, b), max(a, b)
                return float(philox.u01(philox.pair_draws(self.seeds[m1], lo, hi, self.step)[0]))
            count = len(self.multi.simulations)
            lo_pop, hi_pop = min(m1, m2), max(m1, m2)
            ia, ib = (a, b) if m1 < m2 else (b, a)          # id in the lower / in the higher population
            k0, k1 = philox.split_seed(self.seed)
            r = philox.philox4x32_10(np.uint64(ia), self.step + 1, STREAM_CROSS | ((lo_pop * count + hi_pop) << 8),
                                     np.uint64(ib), k0, k1)
            return float(philox.u01(r[0 if m1 < m2 else 1]))   # agent1 is the target
        raise RuntimeError("unexpected np.random.random() in " + name)
