"""Feeds counter-based draws to the UNMODIFIED reference GraphSimulation (TEST INFRASTRUCTURE ONLY).

The reference draws from numpy's global stream in loop order (network/agents.py:82,128,131,165,303,322,338,348,
380,403; graph_abs.py:290-294), which a parallel step cannot reproduce.  `KeyedRandom` replaces
`np.random.normal/random/randint` while the reference's `execute()` runs: each call looks at the calling frame
(which method of which Person / Business is drawing, in which iteration) and returns the draw `GraphDraws`
assigns to that site.  With it the reference -- its own code, its own sequential loop -- and
`oracle.graph_oracle.GraphOracle` must produce the same trajectory, which is how the restatement is pinned and
how `tests/golden/graph_*.json` are generated (oracle/gen_graph_golden.py).
"""
import sys

import numpy as np

from oracle import philox
from oracle.graph_oracle import (GraphDraws, ST_MOVE, ST_UPDATE, ST_HOSP_MOVE, ST_DEATH_MOVE, ST_GOV, ST_ACCOUNT,
                                 GOV_ID, BUSINESS_ID0)


class KeyedRandom(object):
    def __init__(self, sim, seed):
        self.sim = sim
        self.draws = GraphDraws(seed)
        self.pindex = {id(a): k for k, a in enumerate(sim.population)}
        self.bindex = {id(b): k for k, b in enumerate(sim.business)}
        self.calls = {}
        self._saved = None

    # ---- context manager: active only while the reference steps
    def __enter__(self):
        self._saved = (np.random.normal, np.random.random, np.random.randint)
        np.random.normal, np.random.random, np.random.randint = self.normal, self.random, self.randint
        return self

    def __exit__(self, *exc):
        np.random.normal, np.random.random, np.random.randint = self._saved
        return False

    def _nth(self, key):
        n = self.calls.get(key, 0)
        self.calls[key] = n + 1
        return n

    @staticmethod
    def _inside_person_update(frame):
        f = frame
        while f is not None:
            if f.f_code.co_name == "update" and type(f.f_locals.get("self")).__name__ == "Person":
                return True
            f = f.f_back
        return False

    def normal(self, loc, scale, size):
        f = sys._getframe(1)
        person = f.f_locals["self"]
        pid = self.pindex[id(person)]
        name = f.f_code.co_name
        if self._inside_person_update(f):
            stream = ST_HOSP_MOVE if name == "move_to" else ST_DEATH_MOVE
        else:
            stream = ST_MOVE
        z0, z1 = self.draws.normal2(pid, self.sim.iteration, stream)
        return np.array([loc + np.float64(z0) * scale, loc + np.float64(z1) * scale])

    def random(self):
        f = sys._getframe(1)
        name = f.f_code.co_name
        it = self.sim.iteration
        if name == "update":
            pid = self.pindex[id(f.f_locals["self"])]
            w = self.draws.update_uniforms(pid, it)
            return float(philox.u01(w[self._nth(("u", pid, it))]))
        if name == "contact":
            a, b = self.pindex[id(f.f_locals["agent1"])], self.pindex[id(f.f_locals["agent2"])]
            return float(philox.u01(self.draws.contact(a, b, it)[2]))
        raise RuntimeError("unexpected np.random.random() in " + name)

    def randint(self, low, high):
        f = sys._getframe(1)
        name = f.f_code.co_name
        it = self.sim.iteration
        me = f.f_locals.get("self")
        if name == "contact":
            a, b = self.pindex[id(f.f_locals["agent1"])], self.pindex[id(f.f_locals["agent2"])]
            return self.draws.contact(a, b, it)[self._nth(("c", a, b, it))]
        if high <= low:
            raise ValueError("low >= high")          # what numpy does (network/agents.py:127-131 on an empty list)
        if name == "supply":
            buyer = f.f_locals["agent"]
            b = self.bindex[id(me)]
            if id(buyer) in self.pindex:
                return self.draws.shop_qty(self.pindex[id(buyer)], it, b)
            return low + int(GraphDraws.randint(self.draws.words(GOV_ID, it, ST_GOV)[1], high - low))
        if name == "update":                          # the government picks a business (network/agents.py:165)
            return low + int(GraphDraws.randint(self.draws.words(GOV_ID, it, ST_GOV)[0], high - low))
        if name == "accounting":
            b = self.bindex[id(me)]
            return low + int(GraphDraws.randint(self.draws.words(BUSINESS_ID0 + b, it, ST_ACCOUNT)[0], high - low))
        raise RuntimeError("unexpected np.random.randint() in " + name)


SCENARIO_MODES = ("none", "lockdown", "conditional_lockdown", "vertical", "isolation")


def reference_simulation(ref_modules, scenario, np_seed, isolation_rate=0.5, **overrides):
    """Initialise an unmodified reference GraphSimulation for one of the run_batch.py scenario families
    (global_parameters of run_batch.py:53-85; callbacks restated from run_batch.py:8-50, which cannot be imported
    because the script runs at import).  Returns (sim, isolated_ids)."""
    GraphSimulation, Status, EconomicalStatus, util = ref_modules
    params = dict(length=500, height=500, population_size=300, homemates_avg=3, homeless_rate=0.0005,
                  amplitudes={Status.Susceptible: 10, Status.Recovered_Immune: 10, Status.Infected: 10},
                  critical_limit=0.01, contagion_rate=.9, incubation_time=5, contagion_time=10, recovering_time=20,
                  total_wealth=10000000, total_business=9, minimum_income=900.0, minimum_expense=600.0,
                  public_gdp_share=0.1, business_gdp_share=0.5, unemployment_rate=0.12, business_distance=20,
                  initial_infected_perc=.01, initial_immune_perc=.01, contagion_distance=1.)
    params.update(overrides)
    isolated = []

    def lockdown(a):
        if a.house is not None:
            a.house.checkin(a)
        return True

    callbacks = {'on_execute': lambda x: util.sleep(x)}
    if scenario == "lockdown":
        callbacks['on_person_move'] = lockdown
    elif scenario == "conditional_lockdown":
        callbacks['on_person_move'] = lambda a: lockdown(a) if a.environment.get_statistics()['Infected'] > .05 else False
    elif scenario == "vertical":
        callbacks['on_person_move'] = lambda a: lockdown(a) if a.economical_status == EconomicalStatus.Inactive else False
    elif scenario == "isolation":
        callbacks['post_initialize'] = lambda x: util.sample_isolated(x, isolation_rate, isolated)
        callbacks['on_person_move'] = lambda a: util.check_isolation(isolated, a)
    elif scenario != "none":
        raise ValueError(scenario)
    np.random.seed(np_seed)
    sim = GraphSimulation(callbacks=callbacks, **params)
    sim.initialize()
    return sim, list(isolated)
