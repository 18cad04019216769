"""Host side of the GraphSimulation mirror (no GPU): the world builder consumes np.random exactly like the
reference's initialize (graph_abs.py:89-188), so the fixture worlds (built by the reference) are reproduced bit
for bit from the same seed; scenario lowering; API surface."""
import glob
import json
import os

import numpy as np
import pytest

from graph_util import PERSON, HOUSE, BUSINESS

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "graph_*.json")))
SCEN = {"none": "scenario1", "lockdown": "scenario2", "conditional_lockdown": "scenario3", "vertical": "scenario4"}


def _scenario_kwargs(fix):
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    p = fix["world"]["params"]
    kw = dict(sc.global_parameters)
    if fix["scenario"] == "isolation":
        kw.update(sc.partial_isolation(0.5))
    else:
        kw.update(getattr(sc, SCEN[fix["scenario"]]))
    kw.pop("name")
    kw.update(fix["overrides"])
    return kw


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_world_builder_reproduces_reference_initialize(path):
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation
    fix = json.load(open(path))
    sim = GraphSimulation(**_scenario_kwargs(fix))
    np.random.seed(fix["np_seed"])
    world, mode, sleep = sim.build_world()
    ref = fix["world"]
    assert mode == ref["params"]["mode"] and sleep == ref["params"]["sleep"]
    for k in PERSON + HOUSE + BUSINESS:
        a, b = np.asarray(world[k]), np.asarray(ref[k])
        if a.dtype.kind == "f" or b.dtype.kind == "f":
            assert a.astype(np.float64).tobytes() == b.astype(np.float64).tobytes(), k
        else:
            assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), k
    for ours, (grp, key) in (("gov_x", ("gov", "x")), ("gov_y", ("gov", "y")), ("gov_wealth", ("gov", "wealth")),
                             ("gov_fixed", ("gov", "fixed")), ("health_x", ("health", "x")),
                             ("health_y", ("health", "y")), ("health_fixed", ("health", "fixed"))):
        assert world[ours] == ref[grp][key], ours


def test_lowering_rejects_python_callables():
    from covid19_agentbasedsimulation_b200.network.graph_abs import lower_callbacks
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    from covid19_agentbasedsimulation_b200 import _native
    assert lower_callbacks(sc.scenario2["callbacks"])[:2] == (_native.GRAPH_MODE_LOCKDOWN, True)
    assert lower_callbacks(sc.scenario8["callbacks"])[3] == {"contagion_rate": 0.1}
    assert lower_callbacks(sc.scenario6["callbacks"])[2] == 0.5
    with pytest.raises(NotImplementedError):
        lower_callbacks({"on_execute": lambda x: x.iteration % 2 == 0})     # not a shipped behaviour
    with pytest.raises(NotImplementedError):
        lower_callbacks({"on_person_move": lambda a: a.id % 2 == 0})
    with pytest.raises(NotImplementedError):
        lower_callbacks({"on_contact": lambda a, b: True})
    with pytest.raises(NotImplementedError):
        lower_callbacks({"post_initialize": sc.sample_isolated(.3)})


def test_clock_predicates_match_reference_quirks():
    from covid19_agentbasedsimulation_b200.network import util
    assert not util.work_day(0) and util.work_day(24) and not util.work_day(6 * 24)        # days 1 and 7 are off
    assert util.new_month(0) and util.new_month(720) and not util.new_month(721)
    it = 24 + 16                                                                            # hour 16 of a work day
    assert not (util.bed_time(it) or util.work_time(it) or util.lunch_time(it) or util.free_time(it))
    assert util.lunch_time(24 + 12) and util.work_time(24 + 12)


def test_constructor_defaults_match_reference_kwargs():
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation, GRAPH_STAT_KEYS
    g = GraphSimulation()
    assert (g.total_business, g.business_distance, g.homeless_rate, g.unemployment_rate, g.homemates_avg,
            g.homemates_std, g.public_gdp_share, g.business_gdp_share, g.incubation_time, g.contagion_time,
            g.recovering_time, g.iteration) == (10, 10, 0.01, 0.09, 3, 1, 0.1, 0.5, 5, 10, 20, -1)
    assert GRAPH_STAT_KEYS[:7] == ("Q1", "Q2", "Q3", "Q4", "Q5", "Business", "Government")
