"""Off-lattice mode: binary64 positions and per-agent triggers (abs.py:86-95, 169-172, 244-247).

tests/golden/simulation_offlattice.json holds trajectories of the UNMODIFIED reference stepping with population
triggers written as Python lambdas, under keyed draws (oracle/gen_offlattice_golden.py).  CPU: the oracle's sequential
schedule with the declarative form of the same rules reproduces them exactly, and the lambdas themselves are recognised
and lowered to the same rules.  GPU (marked): the device's sequential schedule equals both, its synchronous schedule
equals the synchronous oracle, and the contact-pair hook returns the reference's floating-point pair set.
"""
import json
import os

import numpy as np
import pytest

from oracle.seq_oracle import S, I, R, D
from oracle.sync_oracle import SyncSimulation, contact_pairs_float

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simulation_offlattice.json")
CASES = json.load(open(GOLDEN))
ST = {"Susceptible": S, "Infected": I, "Recovered_Immune": R, "Death": D}
SEV = {"Exposed": 0, "Asymptomatic": 1, "Hospitalization": 2, "Severe": 3}


def _when_oracle(w):
    st = None if "status" not in w else {ST[k] for k in w["status"]}
    sev = None if "infected_status" not in w else {SEV[k] for k in w["infected_status"]}
    lo, hi = w.get("min_age"), w.get("max_age")
    return lambda a: ((st is None or a.status in st) and (sev is None or a.infected_status in sev) and
                      (lo is None or a.age >= lo) and (hi is None or a.age <= hi))


def _oracle_rules(spec):
    out = []
    for r in spec:
        if r["kind"] == "move":
            t = r["target"] if r["target"] == "stay" else tuple(r["target"])
            out.append(dict(kind="move", when=_when_oracle(r["when"]), target=t, sigma=r["sigma"]))
        elif r["attribute"] == "wealth":
            out.append(dict(kind="attr", when=_when_oracle(r["when"]), attribute="wealth", p=r["scale"], q=r["shift"]))
        else:
            out.append(dict(kind="attr", when=_when_oracle(r["when"]), attribute=r["attribute"], p=0.0, q=r["value"]))
    return out


def _oracle(case, schedule):
    kw = {k: v for k, v in case["kwargs"].items() if k not in ("initial_infected_perc", "initial_immune_perc")}
    orc = SyncSimulation(seed=case["seed"], schedule=schedule, total_wealth=case["total_wealth"],
                         rules=_oracle_rules(case["spec"]), **kw)
    i = case["initial"]
    orc.set_population(i["x"], i["y"], i["status"], i["age"], i["stratum"], i["wealth"])
    return orc


def _device_rules(spec):
    from covid19_agentbasedsimulation_b200.rules import MoveRule, AttributeRule, When
    from covid19_agentbasedsimulation_b200.agents import Status, InfectionSeverity
    out = []
    for r in spec:
        w = r["when"]
        when = When(status=[Status[k] for k in w["status"]] if "status" in w else None,
                    infected_status=[InfectionSeverity[k] for k in w["infected_status"]] if "infected_status" in w else None,
                    min_age=w.get("min_age"), max_age=w.get("max_age"))
        if r["kind"] == "move":
            out.append(MoveRule(when, r["target"] if r["target"] == "stay" else tuple(r["target"]), r["sigma"]))
        elif r["attribute"] == "wealth":
            out.append(AttributeRule(when, "wealth", scale=r["scale"], shift=r["shift"]))
        else:
            out.append(AttributeRule(when, r["attribute"], value=r["value"]))
    return out


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_with_declarative_rules_reproduces_the_reference_with_lambdas(case):
    orc = _oracle(case, "sequential")
    for it, row in enumerate(case["rows"]):
        st = orc.execute()
        got = [float(st[k]) for k in case["keys"]]
        assert got[:7] == row[:7], it
        assert np.allclose(got[7:], row[7:], rtol=1e-12, atol=1e-6), it      # Q sums: fixed point vs binary64 order
    f = case["final"]
    assert orc.x.tolist() == f["x"] and orc.y.tolist() == f["y"]             # binary64, bit for bit
    assert orc.status.tolist() == f["status"] and orc.severity.tolist() == f["severity"]
    assert orc.infected_time.tolist() == f["infected_time"]
    assert orc.wealth.tobytes() == np.array(f["wealth"], dtype=np.float64).tobytes()
    assert any(v != np.floor(v) for v in f["x"])                             # the case really left the lattice


def test_reference_lambdas_are_recognised_and_lowered():
    from covid19_agentbasedsimulation_b200.rules import lower_trigger, MoveRule, AttributeRule, When, tabulate
    from covid19_agentbasedsimulation_b200.agents import Status, InfectionSeverity
    state = np.random.get_state()[1].copy()
    stay = lower_trigger({'condition': lambda a: a.status == Status.Infected and
                          a.infected_status == InfectionSeverity.Asymptomatic, 'attribute': 'move',
                          'action': lambda a: (a.x + np.random.normal(0, 0.3, 1), a.y + np.random.normal(0, 0.3, 1))})
    assert isinstance(stay, MoveRule) and stay.target == 'stay' and stay.sigma == 0.3
    assert np.array_equal(tabulate(stay.when), tabulate(When(status=Status.Infected,
                                                             infected_status=InfectionSeverity.Asymptomatic)))
    point = lower_trigger({'condition': lambda a: a.age >= 40, 'attribute': 'move',
                           'action': lambda a: (12.5 + np.random.normal(0, 2.0, 1), 7.25 + np.random.normal(0, 2.0, 1))})
    assert point.target == (12.5, 7.25) and point.sigma == 2.0
    tax = lower_trigger({'condition': lambda a: a.age > 60, 'attribute': 'wealth', 'action': lambda w: w * 0.99})
    assert isinstance(tax, AttributeRule) and (tax.p, tax.q) == (0.99, 0.0)
    reset = lower_trigger({'condition': lambda a: a.status == Status.Recovered_Immune, 'attribute': 'infected_time',
                           'action': lambda t: 7})
    assert (reset.attribute, reset.q) == ("infected_time", 7.0)
    cure = lower_trigger({'condition': lambda a: a.social_stratum == 4, 'attribute': 'status',
                          'action': lambda s: Status.Recovered_Immune})
    assert cure.q == 2.0
    assert np.array_equal(np.random.get_state()[1], state)                   # probing consumed nothing from np.random
    with pytest.raises(NotImplementedError):                                 # depends on the position: not tabulable
        lower_trigger({'condition': lambda a: a.x > 5, 'attribute': 'move', 'action': lambda a: (a.x, a.y)})
    with pytest.raises(NotImplementedError):                                 # not point + normal
        lower_trigger({'condition': lambda a: True, 'attribute': 'move', 'action': lambda a: (a.x * 2.0, a.y)})
    with pytest.raises(NotImplementedError):
        lower_trigger({'condition': lambda a: True, 'attribute': 'wealth', 'action': lambda w: w * w})


def test_float_pair_predicate_is_the_reference_one():
    # 0.3 / 0.4 / 0.5: sqrt(fl(.09) + fl(.16)) rounds to exactly 0.5 in binary64 -- inside; a few ulps further the
    # rounded square root exceeds 0.5 -- outside.  The verdict is the rounded arithmetic's, not the real numbers'.
    far = 0.4
    while np.sqrt(np.float64(0.3) * 0.3 + far * far) <= 0.5:
        far = np.nextafter(far, 1.0)
    x = np.array([0.0, 0.3, 0.3])
    y = np.array([0.0, 0.4, far])
    assert contact_pairs_float(x, y, 0.5).tolist() == [[0, 1], [1, 2]]
    assert contact_pairs_float(x, np.array([0.0, 0.4, np.nextafter(far, 0.0)]), 0.5).tolist() == [[0, 1], [0, 2], [1, 2]]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_sequential_equals_the_reference_trajectory(native_lib, case):
    from covid19_agentbasedsimulation_b200.abs import Simulation
    kw = {k: v for k, v in case["kwargs"].items() if k not in ("initial_infected_perc", "initial_immune_perc")}
    sim = Simulation(seed=case["seed"], schedule="sequential", total_wealth=case["total_wealth"],
                     triggers_population=_device_rules(case["spec"]), **kw)
    i = case["initial"]
    sim.set_population_arrays(i["x"], i["y"], i["status"], i["age"], i["stratum"], i["wealth"])
    orc = _oracle(case, "sequential")
    for it in range(case["iterations"]):
        sim.execute()
        st = orc.execute()
        g = sim.get_statistics('all')
        assert [float(g[k]) for k in case["keys"]] == [float(st[k]) for k in case["keys"]], it
        assert np.array_equal(sim.get_contact_pairs(), orc.last_pairs), it
    h, f = sim._download(), case["final"]
    assert h["x"].tolist() == f["x"] and h["y"].tolist() == f["y"]
    assert h["status"].tolist() == f["status"] and h["severity"].tolist() == f["severity"]
    assert h["infected_time"].tolist() == f["infected_time"]
    assert h["wealth"].tobytes() == np.array(f["wealth"], dtype=np.float64).tobytes()
    assert np.array_equal(np.array(case["rows"])[:, :7], np.array(
        [[float(v) for v in r[:7]] for r in case["rows"]]))


@pytest.mark.gpu
def test_gpu_synchronous_off_lattice_equals_oracle_at_scale(native_lib):
    """20 000 agents, unmodified lambdas handed to append_trigger_population, tiled kernels, run() in one enqueue."""
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from covid19_agentbasedsimulation_b200.agents import Status
    from oracle.sync_oracle import synthetic_population
    n, side = 20000, 200
    pop, tw = synthetic_population(n, side, side, seed=3, infected=0.03)
    kw = dict(population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=0.8, contagion_rate=0.7)
    gpu = Simulation(seed=9, **kw)
    gpu.append_trigger_population(lambda a: a.status == Status.Infected, 'move',
                                  lambda a: (a.x + np.random.normal(0, 0.4, 1), a.y + np.random.normal(0, 0.4, 1)))
    gpu.append_trigger_population(lambda a: a.age < 18, 'wealth', lambda w: w + 1.5)
    gpu.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    orc = SyncSimulation(seed=9, rules=[
        dict(kind="move", when=lambda a: a.status == I, target="stay", sigma=0.4),
        dict(kind="attr", when=lambda a: a.age < 18, attribute="wealth", p=1.0, q=1.5)], **kw)
    orc.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    rows = gpu.run(8)
    for it in range(8):
        o = orc.execute()
        assert rows[it].tolist() == [float(v) for v in o.values()], it
    h = gpu._download()
    assert np.array_equal(h["x"], orc.x) and np.array_equal(h["y"], orc.y) and np.array_equal(h["status"], orc.status)
    assert h["wealth"].tobytes() == orc.wealth.tobytes()
    assert np.array_equal(gpu.get_contact_pairs(), orc.last_pairs)
    assert isinstance(gpu.get_positions()[0][0], float)


@pytest.mark.gpu
def test_lattice_handle_switches_to_float64_when_a_trigger_is_appended_mid_run(native_lib):
    from covid19_agentbasedsimulation_b200.abs import Simulation, MoveRule, When
    from covid19_agentbasedsimulation_b200.agents import Status
    from oracle.sync_oracle import synthetic_population
    n, side = 3000, 80
    pop, tw = synthetic_population(n, side, side, seed=5, infected=0.05)
    kw = dict(population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=1.0)
    gpu = Simulation(seed=4, **kw)
    gpu.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    orc = SyncSimulation(seed=4, **kw)
    orc.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    for _ in range(3):
        gpu.execute()
        orc.execute()
    gpu.triggers_population.append(MoveRule(When(status=Status.Infected), 'stay', 0.25))
    orc.rules.append(dict(kind="move", when=lambda a: a.status == I, target="stay", sigma=0.25))
    orc.float_positions = True
    orc.x, orc.y = orc.x.astype(np.float64), orc.y.astype(np.float64)
    for it in range(4):
        gpu.execute()
        o = orc.execute()
        g = gpu.get_statistics('all')
        assert [float(g[k]) for k in o] == [float(v) for v in o.values()], it     # same draw clock across the switch
    h = gpu._download()
    assert np.array_equal(h["x"], orc.x) and np.array_equal(h["status"], orc.status)
