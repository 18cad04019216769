"""MultiPopulationSimulation (abs.py:328-415): oracle pinned to the reference, host mirror, device parity.

tests/golden/multi_population.json was produced by the UNMODIFIED reference class stepping under keyed draws
(oracle/multi_shim.py, oracle/gen_multi_golden.py): the oracle with the reference's own loop order
(cross_schedule="sequential", populations on the sequential schedule) must reproduce it exactly.  The device computes
the synchronous cross rule: it must equal the oracle's synchronous mode bit for bit (marked gpu), and that mode equals
the sequential one whenever no in-loop chain occurs.
"""
import json
import os

import numpy as np
import pytest

from oracle.multi_oracle import MultiPopulationOracle, KEYS
from oracle.sync_oracle import SyncSimulation

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multi_population.json")
CASES = json.load(open(GOLDEN))
REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")
AMP = {"s": 0, "i": 1, "c": 2}      # Status value -> status code of the oracle


def _oracle(case, schedule, cross_schedule):
    sims = []
    for k, kw in enumerate(case["sims"]):
        kw = {a: b for a, b in kw.items() if a not in ("initial_infected_perc", "initial_immune_perc", "amplitudes")}
        amp = {AMP[a]: b for a, b in case["sims"][k]["amplitudes"].items()}
        s = SyncSimulation(seed=case["seeds"][k], schedule=schedule, total_wealth=case["total_wealth"][k],
                           amplitudes=amp, **kw)
        i = case["initial"][k]
        s.set_population(i["x"], i["y"], i["status"], i["age"], i["stratum"], i["wealth"])
        sims.append(s)
    return MultiPopulationOracle(sims, case["positions"], case["multi"]["contagion_distance"], case["seed"],
                                 cross_schedule=cross_schedule)


def _rows(orc, iterations):
    rows, counts = [], []
    for _ in range(iterations):
        orc.execute()
        st = orc.get_statistics()
        rows.append([float(st[k]) for k in KEYS])
        counts.append([[int(np.sum(s.status == c)) for c in range(4)] for s in orc.simulations])
    return rows, counts


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_reference(case):
    orc = _oracle(case, "sequential", "sequential")
    rows, counts = _rows(orc, case["iterations"])
    assert counts == case["status_counts"]
    got, want = np.array(rows), np.array(case["rows"])
    assert np.array_equal(got[:, :8], want[:, :8])
    assert np.allclose(got[:, 8:], want[:, 8:], rtol=1e-12, atol=1e-6)      # Q sums: fixed point vs binary64 order
    assert orc.cross_infections > 0
    for s, f in zip(orc.simulations, case["final"]):
        assert s.x.tolist() == f["x"] and s.y.tolist() == f["y"] and s.status.tolist() == f["status"]
        assert s.severity.tolist() == f["severity"] and s.infected_time.tolist() == f["infected_time"]
        assert s.wealth.tobytes() == np.array(f["wealth"], dtype=np.float64).tobytes()


def test_statistics_keep_the_reference_quirk():
    case = CASES[1]
    orc = _oracle(case, "synchronous", "synchronous")
    orc.execute()
    st = orc.get_statistics()
    assert list(st.keys()) == list(KEYS) and st["Exposed"] == 0.0
    last = orc.simulations[-1]
    total = sum(s.population_size for s in orc.simulations)
    assert st["Infected"] == np.float64(np.sum(last.status == 1)) / total      # the last population only
    assert st["Q3"] == last.get_statistics()["Q3"]


def test_synchronous_cross_rule_versus_sequential():
    """With the populations themselves on the same schedule, the two cross rules differ only through an agent that is
    infected and transmits inside one double loop: rare where two towns touch at a corner (identical for many steps),
    immediate where three dense populations overlap -- the documented divergence of the device's rule."""
    same = {}
    for case in CASES:
        a = _oracle(case, "sequential", "sequential")
        b = _oracle(case, "sequential", "synchronous")
        same[case["name"]] = 0
        for _ in range(case["iterations"]):
            a.execute()
            b.execute()
            if not all(np.array_equal(x.status, y.status) for x, y in zip(a.simulations, b.simulations)):
                break
            same[case["name"]] += 1
    assert same["legacy_two_towns"] >= 10
    assert same["three_overlapping"] < CASES[1]["iterations"]


def test_host_mirror_api():
    from covid19_agentbasedsimulation_b200.abs import MultiPopulationSimulation, Simulation
    m = MultiPopulationSimulation(length=200, height=200, contagion_distance=5.)
    assert m.simulations == [] and m.positions == [] and m.total_population == 0
    s1, s2 = Simulation(population_size=80), Simulation(population_size=60)
    m.append(s1, (0, 0))
    m.append(s2, (80, 80))
    assert m.total_population == 140 and m.positions == [(0, 0), (80, 80)]
    assert m._offset(1) == (80, 80)
    m.positions[1] = (80.5, 80)
    with pytest.raises(NotImplementedError):
        m._offset(1)
    with pytest.raises(RuntimeError):
        m.execute()                    # not initialised: no device work was attempted


@pytest.mark.reference
def test_live_reference_under_keyed_draws_equals_oracle():
    import sys
    import warnings
    if not os.path.isdir(os.path.join(REF, "covid_abs")):
        pytest.skip("reference not present on this machine")
    warnings.filterwarnings("ignore")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from covid_abs.abs import Simulation, MultiPopulationSimulation
    from covid_abs.agents import Status
    from oracle.multi_shim import KeyedMultiRandom
    np.random.seed(5)
    amp = {Status.Susceptible: 4, Status.Infected: 4, Status.Recovered_Immune: 4}
    kws = [dict(population_size=40, length=30, height=30, contagion_distance=2., initial_infected_perc=0.1,
                amplitudes=amp),
           dict(population_size=30, length=30, height=30, contagion_distance=2., initial_infected_perc=0.0,
                contagion_rate=0.6, amplitudes=amp)]
    sims = [Simulation(**kw) for kw in kws]
    multi = MultiPopulationSimulation(simulations=sims, positions=[(0, 0), (10, 5)], total_population=70,
                                      contagion_distance=2.5)
    multi.initialize()
    code = {Status.Susceptible: 0, Status.Infected: 1, Status.Recovered_Immune: 2, Status.Death: 3}
    osims = []
    for k, (sim, kw) in enumerate(zip(sims, kws)):
        pop = sim.get_population()
        o = SyncSimulation(seed=70 + k, schedule="sequential", total_wealth=sim.total_wealth,
                           amplitudes={0: 4, 1: 4, 2: 4},
                           **{a: b for a, b in kw.items() if a not in ("initial_infected_perc", "amplitudes")})
        o.set_population([int(a.x) for a in pop], [int(a.y) for a in pop], [code[a.status] for a in pop],
                         [int(a.age) for a in pop], [int(a.social_stratum) for a in pop],
                         [float(np.asarray(a.wealth).ravel()[0]) for a in pop])
        sim.get_statistics()
        osims.append(o)
    orc = MultiPopulationOracle(osims, [(0, 0), (10, 5)], 2.5, 77, cross_schedule="sequential")
    with KeyedMultiRandom(multi, [70, 71], 77) as kr:
        for it in range(25):
            multi.execute()
            orc.execute()
            kr.next_step()
            for sim, o in zip(sims, osims):
                assert [code[a.status] for a in sim.get_population()] == o.status.tolist(), it
            ref, got = multi.get_statistics('all'), orc.get_statistics()
            assert [float(ref[k]) for k in KEYS[:8]] == [float(got[k]) for k in KEYS[:8]]
    assert orc.cross_infections > 0


def _device_multi(case, schedule):
    from covid19_agentbasedsimulation_b200.abs import MultiPopulationSimulation, Simulation
    from covid19_agentbasedsimulation_b200.agents import Status
    sims = []
    for k, kw in enumerate(case["sims"]):
        kw = dict(kw)
        kw["amplitudes"] = {Status(a): b for a, b in kw["amplitudes"].items()}
        s = Simulation(seed=case["seeds"][k], schedule=schedule, total_wealth=case["total_wealth"][k], **kw)
        i = case["initial"][k]
        s.set_population_arrays(i["x"], i["y"], i["status"], i["age"], i["stratum"], i["wealth"])
        sims.append(s)
    multi = MultiPopulationSimulation(simulations=sims, positions=[tuple(p) for p in case["positions"]],
                                      total_population=sum(s.population_size for s in sims), seed=case["seed"],
                                      **case["multi"])
    return multi


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", ["synchronous", "sequential"])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_equals_oracle(native_lib, case, schedule):
    orc = _oracle(case, schedule, "synchronous")
    multi = _device_multi(case, schedule)
    for it in range(case["iterations"]):
        orc.execute()
        multi.execute()
        want, got = orc.get_statistics(), multi.get_statistics('all')
        assert list(got.keys()) == list(KEYS)
        assert [float(got[k]) for k in KEYS] == [float(want[k]) for k in KEYS], it
    assert orc.cross_infections > 0
    for sim, o in zip(multi.simulations, orc.simulations):
        h = sim._download()
        order = np.argsort(h["id"], kind="stable")
        for k in ("x", "y", "status", "severity", "infected_time"):
            assert np.array_equal(h[k][order].astype(np.int64), getattr(o, k)), k
        assert h["wealth"][order].tobytes() == o.wealth.tobytes()
    assert len(multi.get_positions()) == multi.total_population
    assert len(multi.get_population()) == multi.total_population


@pytest.mark.gpu
def test_device_cross_contacts_at_larger_sizes(native_lib):
    """Two populations of 3 000 and 5 000 agents (tiled step path, several blocks of the cross kernel)."""
    from covid19_agentbasedsimulation_b200.abs import MultiPopulationSimulation, Simulation
    from oracle.sync_oracle import synthetic_population
    specs = [(3000, 120, 31, 0.05), (5000, 150, 32, 0.0)]
    sims, osims = [], []
    for n, side, seed, inf in specs:
        pop, tw = synthetic_population(n, side, side, seed=seed, infected=inf, immune=0.01)
        kw = dict(seed=seed, population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=1.0,
                  contagion_rate=0.8)
        s = Simulation(**kw)
        s.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
        o = SyncSimulation(**kw)
        o.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
        sims.append(s)
        osims.append(o)
    multi = MultiPopulationSimulation(simulations=sims, positions=[(0, 0), (60, 40)], total_population=8000, seed=99,
                                      contagion_distance=1.5)
    orc = MultiPopulationOracle(osims, [(0, 0), (60, 40)], 1.5, 99)
    for it in range(12):
        multi.execute()
        orc.execute()
        for sim, o in zip(sims, osims):
            h = sim._download()
            order = np.argsort(h["id"], kind="stable")
            assert np.array_equal(h["status"][order].astype(np.int64), o.status), it
    assert orc.cross_infections > 0
    got, want = multi.get_statistics('all'), orc.get_statistics()
    assert [float(got[k]) for k in KEYS] == [float(want[k]) for k in KEYS]


@pytest.mark.gpu
def test_batch_experiment_drives_multi_population(native_lib, tmp_path):
    """run_batch.py:645-678: the same population objects are re-initialised by every experiment."""
    from covid19_agentbasedsimulation_b200.abs import MultiPopulationSimulation, Simulation
    from covid19_agentbasedsimulation_b200.agents import Status
    from covid19_agentbasedsimulation_b200.experiments import batch_experiment
    common = dict(initial_infected_perc=0.05, length=100, height=100, population_size=80, contagion_distance=5.,
                  critical_limit=.1)
    sim1 = Simulation(amplitudes={Status.Susceptible: 5, Status.Recovered_Immune: 5, Status.Infected: 5}, **common)
    sim2 = Simulation(amplitudes={Status.Susceptible: 1, Status.Recovered_Immune: 1, Status.Infected: 1}, **common)
    np.random.seed(3)
    df = batch_experiment(4, 30, str(tmp_path / "m.csv"), simulation_type=MultiPopulationSimulation, length=200,
                          height=200, contagion_distance=5., critical_limit=.1, simulations=[sim1, sim2],
                          positions=[(0, 0), (80, 80)], total_population=160)
    assert list(df.columns) == ['Iteration', 'Metric', 'Min', 'Avg', 'Std', 'Max']
    assert len(df) == 30 * len(KEYS) and list(df['Metric'][:len(KEYS)]) == list(KEYS)
    s = df[df['Metric'] == 'Susceptible']
    assert s['Max'].max() <= 0.5 + 1e-12          # the quirk: the last population's count over both populations
