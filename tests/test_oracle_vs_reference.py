"""Runs the UNMODIFIED reference next to the oracle (only where /root/reference exists: the authoring
container).  The committed fixtures of tests/golden cover the same ground where it does not."""
import os
import sys
import warnings

import numpy as np
import pytest

REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")
pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "covid_abs")), reason="reference not mounted")]


@pytest.fixture(scope="module")
def ref():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, REF)
    try:
        from covid_abs import abs as ref_abs, agents as ref_agents
        yield ref_abs, ref_agents
    finally:
        sys.path.remove(REF)


def run_both(ref, seed, iterations, **kw):
    from oracle.seq_oracle import SeqSimulation
    ref_abs, ref_agents = ref
    St, Sev = ref_agents.Status, ref_agents.InfectionSeverity
    stc = {St.Susceptible: 0, St.Infected: 1, St.Recovered_Immune: 2, St.Death: 3}
    sevc = {Sev.Exposed: 0, Sev.Asymptomatic: 1, Sev.Hospitalization: 2, Sev.Severe: 3}
    rkw = dict(kw)
    if "amplitudes" in kw:
        rkw["amplitudes"] = {s: kw["amplitudes"][c] for s, c in stc.items() if c in kw["amplitudes"]}
    np.random.seed(seed)
    r = ref_abs.Simulation(**rkw)
    r.initialize()
    np.random.seed(seed)
    o = SeqSimulation(**kw)
    o.initialize()
    st_r = st_o = np.random.get_state()
    for it in range(iterations):
        np.random.set_state(st_r); r.execute(); st_r = np.random.get_state()
        np.random.set_state(st_o); o.execute(); st_o = np.random.get_state()
        rs, os_ = r.get_statistics('all'), o.get_statistics('all')
        assert list(rs.keys()) == list(os_.keys())
        for k in rs:
            assert float(rs[k]) == float(os_[k]), (seed, it, k)
        pop = r.get_population()
        assert [int(a.x) for a in pop] == o.x.tolist()
        assert [int(a.y) for a in pop] == o.y.tolist()
        assert [stc[a.status] for a in pop] == o.status.tolist()
        assert [sevc[a.infected_status] for a in pop] == o.severity.tolist()
        assert [a.infected_time for a in pop] == o.infected_time.tolist()
        assert [float(np.asarray(a.wealth).ravel()[0]) for a in pop] == o.wealth.tolist()


@pytest.mark.parametrize("seed", range(6))
def test_defaults_bit_identical(ref, seed):
    run_both(ref, seed, 50)


@pytest.mark.parametrize("seed", range(2))
def test_legacy_scenario_bit_identical(ref, seed):
    from oracle.seq_oracle import S, I, R
    run_both(ref, seed, 60, initial_infected_perc=0.02, initial_immune_perc=0.01, length=100, height=100,
             population_size=80, contagion_distance=5., critical_limit=0.05, amplitudes={S: 5, R: 5, I: 5})


def test_product_initialize_matches_reference(ref):
    """The product's host-side initialize() consumes np.random like abs.py:97-139: same initial population."""
    from covid19_agentbasedsimulation_b200.abs import Simulation as Product
    ref_abs, _ = ref
    for seed in range(3):
        np.random.seed(seed)
        r = ref_abs.Simulation(population_size=50)
        r.initialize()
        np.random.seed(seed)
        p = Product(population_size=50)
        captured = {}
        p.set_population_arrays = lambda x, y, status, age, stratum, wealth, **k: captured.update(
            x=x, y=y, status=status, age=age, stratum=stratum, wealth=wealth)
        p.initialize()
        pop = r.get_population()
        assert [int(a.x) for a in pop] == list(captured["x"])
        assert [int(a.y) for a in pop] == list(captured["y"])
        assert [a.age for a in pop] == list(captured["age"])
        assert [a.social_stratum for a in pop] == list(captured["stratum"])
        assert [float(a.wealth) for a in pop] == list(captured["wealth"])
        assert [a.status.value for a in pop] == ['sicm'[s] for s in captured["status"]]
