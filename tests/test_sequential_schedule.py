"""The sequential contact schedule (abs.py:261-265 as written, in-step infection chains).

tests/golden/simulation_sequential.json was produced by the UNMODIFIED reference stepping under keyed draws
(oracle/sim_shim.py, oracle/gen_sequential_golden.py).  The oracle's sequential schedule must reproduce it exactly
(CPU); the GPU's sequential schedule must equal both (marked gpu); the synchronous schedule must differ on these cases
(they were chosen so that chains occur) -- which is the documented divergence of the default schedule.
"""
import json
import os

import numpy as np
import pytest

from oracle.sync_oracle import SyncSimulation

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simulation_sequential.json")
CASES = json.load(open(GOLDEN))
REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")


def _oracle(case, schedule):
    kw = {k: v for k, v in case["kwargs"].items() if k not in ("initial_infected_perc", "initial_immune_perc")}
    orc = SyncSimulation(seed=case["seed"], schedule=schedule, total_wealth=case["total_wealth"], **kw)
    i = case["initial"]
    orc.set_population(i["x"], i["y"], i["status"], i["age"], i["stratum"], i["wealth"])
    return orc


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_sequential_reproduces_reference(case):
    orc = _oracle(case, "sequential")
    for it, row in enumerate(case["rows"]):
        st = orc.execute()
        got = [float(st[k]) for k in case["keys"]]
        assert got[:7] == row[:7], it
        assert np.allclose(got[7:], row[7:], rtol=1e-12, atol=1e-6), it      # Q sums: fixed point vs binary64 order
    f = case["final"]
    assert orc.x.tolist() == f["x"] and orc.y.tolist() == f["y"] and orc.status.tolist() == f["status"]
    assert orc.severity.tolist() == f["severity"] and orc.infected_time.tolist() == f["infected_time"]
    assert orc.wealth.tobytes() == np.array(f["wealth"], dtype=np.float64).tobytes()


def test_synchronous_schedule_differs_where_chains_occur():
    differs = 0
    for case in CASES:
        orc = _oracle(case, "synchronous")
        rows = np.array([[float(orc.execute()[k]) for k in case["keys"][:4]] for _ in range(case["iterations"])])
        differs += int(not np.array_equal(rows, np.array(case["rows"])[:, :4]))
    assert differs >= 2


@pytest.mark.reference
def test_live_reference_under_keyed_draws_equals_oracle():
    import sys
    import warnings
    if not os.path.isdir(os.path.join(REF, "covid_abs")):
        pytest.skip("reference not present on this machine")
    warnings.filterwarnings("ignore")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from covid_abs.abs import Simulation
    from covid_abs.agents import Status
    from oracle.sim_shim import KeyedSimulationRandom
    stc = {Status.Susceptible: 0, Status.Infected: 1, Status.Recovered_Immune: 2, Status.Death: 3}
    kw = dict(population_size=60, length=25, height=25, contagion_distance=2., initial_infected_perc=0.05)
    np.random.seed(9)
    ref = Simulation(**kw)
    ref.initialize()
    pop = ref.get_population()
    orc = SyncSimulation(seed=31, schedule="sequential", total_wealth=ref.total_wealth, population_size=60, length=25,
                         height=25, contagion_distance=2.)
    orc.set_population([a.x for a in pop], [a.y for a in pop], [stc[a.status] for a in pop], [a.age for a in pop],
                       [a.social_stratum for a in pop], [float(np.asarray(a.wealth).ravel()[0]) for a in pop])
    ref.get_statistics()
    with KeyedSimulationRandom(ref, 31) as kr:
        for it in range(40):
            ref.execute()
            ref.get_statistics('all')
            kr.next_step()
            orc.execute()
            assert [int(a.x) for a in pop] == orc.x.tolist() and [stc[a.status] for a in pop] == orc.status.tolist(), it
            assert np.array([float(np.asarray(a.wealth).ravel()[0]) for a in pop]).tobytes() == orc.wealth.tobytes(), it


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpu_sequential_equals_reference_and_oracle(native_lib, case):
    from covid19_agentbasedsimulation_b200.abs import Simulation
    kw = {k: v for k, v in case["kwargs"].items() if k not in ("initial_infected_perc", "initial_immune_perc")}
    sim = Simulation(seed=case["seed"], schedule="sequential", total_wealth=case["total_wealth"], **kw)
    i = case["initial"]
    sim.set_population_arrays(i["x"], i["y"], i["status"], i["age"], i["stratum"], i["wealth"])
    orc = _oracle(case, "sequential")
    table = sim.run(case["iterations"])
    for it in range(case["iterations"]):
        st = orc.execute()
        assert table[it].tolist() == [float(st[k]) for k in case["keys"]], it
    assert np.array_equal(table[:, :7], np.array(case["rows"])[:, :7])
    h, f = sim._download(), case["final"]
    assert h["x"].tolist() == f["x"] and h["y"].tolist() == f["y"] and h["status"].tolist() == f["status"]
    assert h["wealth"].tobytes() == np.array(f["wealth"], dtype=np.float64).tobytes()


@pytest.mark.gpu
def test_gpu_sequential_large_population_matches_oracle(native_lib):
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from oracle.sync_oracle import synthetic_population
    n, side = 20000, 120
    pop, tw = synthetic_population(n, side, side, seed=8, infected=0.01)
    kw = dict(population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=1.5, contagion_rate=0.95)
    gpu = Simulation(seed=77, schedule="sequential", **kw)
    gpu.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    orc = SyncSimulation(seed=77, schedule="sequential", **kw)
    orc.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    for it in range(6):
        gpu.execute()
        o = orc.execute()
        g = gpu.get_statistics('all')
        assert [float(g[k]) for k in o] == [float(v) for v in o.values()], it
    assert np.array_equal(gpu._download()["status"], orc.status)
