"""Behaviour of the covid_abs-style host API on the GPU: triggers, attribute changes between steps, accessors,
history ring, error paths (abs.py:57-95, 232-322)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pop(n, side, seed, infected=0.05):
    from oracle.sync_oracle import synthetic_population
    return synthetic_population(n, side, side, seed=seed, infected=infected, immune=0.01)


def _gpu(pop, tw, side, **kw):
    from covid19_agentbasedsimulation_b200.abs import Simulation
    sim = Simulation(population_size=len(pop["x"]), length=side, height=side, total_wealth=tw, **kw)
    sim.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    return sim


def test_python_lambda_triggers_equal_declarative_ones(native_lib):
    """The conditional lockdown of run_batch.py:628-643 written with lambdas (host, abs.py:267-271) and as device
    triggers gives the same run."""
    from covid19_agentbasedsimulation_b200.abs import Trigger
    from covid19_agentbasedsimulation_b200.agents import Status
    pop, tw = _pop(5000, 100, 2, infected=0.15)
    low = {Status.Susceptible: 1.5, Status.Recovered_Immune: 1.5, Status.Infected: 1.5}
    high = {Status.Susceptible: 5, Status.Recovered_Immune: 5, Status.Infected: 5}
    a = _gpu(pop, tw, 100, seed=4, triggers_simulation=[Trigger('Infected', '>=', .2, 'amplitudes', low),
                                                         Trigger('Infected', '<=', .05, 'amplitudes', high)])
    b = _gpu(pop, tw, 100, seed=4)
    # the reference's memo semantics: the lambda sees the statistics memoised after the previous step
    memo = {"prev": b.get_statistics('all')}
    b.append_trigger_simulation(lambda s: memo["prev"]['Infected'] >= .2, 'amplitudes', lambda x: low)
    b.append_trigger_simulation(lambda s: memo["prev"]['Infected'] <= .05, 'amplitudes', lambda x: high)
    rows_a = a.run(30)
    for it in range(30):
        b.execute()
        cur = dict(b.get_statistics('all'))
        assert [float(cur[k]) for k in cur] == rows_a[it].tolist(), it
        memo["prev"] = cur
    assert b.amplitudes in (low, high)


def test_attribute_change_between_steps_reaches_the_device(native_lib):
    from oracle.sync_oracle import SyncSimulation
    pop, tw = _pop(4000, 90, 3)
    sim = _gpu(pop, tw, 90, seed=6)
    orc = SyncSimulation(seed=6, population_size=4000, length=90, height=90, total_wealth=tw)
    orc.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    for it in range(12):
        if it == 4:
            sim.contagion_rate, orc.contagion_rate = 0.2, 0.2
        if it == 8:
            sim.contagion_distance, orc.contagion_distance = 2.0, 2.0
        sim.execute()
        o = orc.execute()
        g = sim.get_statistics('all')
        assert [float(g[k]) for k in o] == [float(v) for v in o.values()], it
    assert np.array_equal(sim.get_contact_pairs(), orc.last_pairs)


def test_accessors_and_population_round_trip(native_lib):
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from covid19_agentbasedsimulation_b200.agents import Status, Agent
    np.random.seed(5)
    sim = Simulation(seed=8)
    sim.initialize()
    sim.execute()
    pop = sim.get_population()
    assert len(pop) == 20 and all(isinstance(a, Agent) for a in pop)
    assert sim.get_positions() == [[a.x, a.y] for a in pop]
    assert sim.get_description() == [a.status.name for a in pop]
    assert len(sim.get_description(complete=True)) == 20
    assert set(sim.get_statistics('info')) == {"Susceptible", "Infected", "Recovered_Immune", "Death", "Asymptomatic",
                                               "Hospitalization", "Severe"}
    assert set(sim.get_statistics('ecom')) == {"Q1", "Q2", "Q3", "Q4", "Q5"}
    before = sim.get_statistics('all')
    clone = Simulation(seed=8)
    clone.set_population(pop)                      # abs.py:65-69
    assert clone.population_size == 20
    after = clone.get_statistics('all')
    assert [float(before[k]) for k in before] == [float(after[k]) for k in after]
    assert sum(1 for a in clone.get_population() if a.status == Status.Infected) == \
        sum(1 for a in pop if a.status == Status.Infected)


def test_history_ring_and_error_paths(native_lib):
    from covid19_agentbasedsimulation_b200 import _native
    from covid19_agentbasedsimulation_b200.abs import Simulation
    pop, tw = _pop(1000, 50, 1)
    sim = _gpu(pop, tw, 50, seed=1, history_capacity=8)
    rows = sim.run(8)
    assert rows.shape == (8, 12)
    long = sim.run(20)                              # longer than the ring: falls back to step-by-step reads
    assert long.shape == (20, 12)
    with pytest.raises(_native.NativeError) as e:
        sim._handle.stats_history(0, 4)             # overwritten long ago
    assert e.value.code == _native.ERR_CAPACITY
    with pytest.raises(_native.NativeError):
        sim._handle.stats_history(27, 5)            # not executed yet
    # per-agent triggers the device cannot express are refused (no CPU fallback); the expressible ones run
    # (tests/test_offlattice.py)
    sim.append_trigger_population(lambda a: a.wealth > 10, 'wealth', lambda w: 0)     # condition on wealth: not tabulable
    with pytest.raises(NotImplementedError):
        sim.execute()
    sim.triggers_population = []
    sim.append_trigger_population(lambda a: True, 'wealth', lambda w: 0)
    sim.execute()
    assert sum(float(sim.get_statistics('all')[k]) for k in ("Q1", "Q2", "Q3", "Q4", "Q5")) == 0.0
    # coordinates beyond the supported box are reported, not silently wrapped
    far = Simulation(seed=1, population_size=2, length=10, height=10)
    far.set_population_arrays([0, 2 ** 29], [0, 0], [0, 1], [30, 30], [1, 1], [1.0, 1.0])
    with pytest.raises(_native.NativeError) as e:
        far.get_statistics('all')
    assert e.value.code == _native.ERR_DOMAIN
    # a bigger population than the handle was created for re-creates the handle
    small = _gpu(*_pop(100, 20, 2), 20, seed=2)
    big, tw2 = _pop(5000, 80, 3)
    small.length = small.height = 80
    small.set_population_arrays(big["x"], big["y"], big["status"], big["age"], big["stratum"], big["wealth"])
    small.execute()
    assert abs(sum(float(small.get_statistics('info')[k]) for k in ("Susceptible", "Infected", "Recovered_Immune",
                                                                    "Death")) - 1.0) < 1e-12


def test_small_population_kernel_and_tiled_path_agree(native_lib):
    """Populations of at most 1024 agents run in one resident block; the tiled path (forced here through the
    sequential schedule's machinery being off, i.e. a 1025-agent population with one far-away bystander) and a mid-run
    switch of schedule give the same trajectory as the oracle."""
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from oracle.sync_oracle import SyncSimulation, synthetic_population
    n, side = 1024, 60
    pop, tw = synthetic_population(n, side, side, seed=12, infected=0.05)
    kw = dict(population_size=n, length=side, height=side, total_wealth=tw, contagion_distance=1.5)
    small = Simulation(seed=3, **kw)
    small.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    orc = SyncSimulation(seed=3, **kw)
    orc.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    rows = small.run(10)                                     # resident kernel, one launch
    assert small.device_info()["kernel_launches"] < 20
    for it in range(10):
        o = orc.execute()
        assert rows[it].tolist() == [float(v) for v in o.values()], it
    assert np.array_equal(small.get_contact_pairs(), orc.last_pairs)   # re-sorts on demand
    small.schedule = "sequential"                            # tiled path from here on; no chains at this density?
    orc.schedule = "sequential"
    for it in range(5):
        small.execute()
        o = orc.execute()
        g = small.get_statistics('all')
        assert [float(g[k]) for k in o] == [float(v) for v in o.values()], it
    h = small._download()
    assert np.array_equal(h["x"], orc.x) and np.array_equal(h["status"], orc.status)
    assert h["wealth"].tobytes() == orc.wealth.tobytes()


def test_bind_device_equals_host_upload(native_lib):
    """cabs_bind_device: columns that already live on the GPU (torch tensors) give the same run as the host upload."""
    import torch
    from covid19_agentbasedsimulation_b200 import _native
    from covid19_agentbasedsimulation_b200.abs import Simulation
    pop, tw = _pop(30000, 400, 11)
    host = _gpu(pop, tw, 400, seed=5)
    dev = Simulation(seed=5, population_size=30000, length=400, height=400, total_wealth=tw)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    dev.set_population_tensors(t(pop["x"], np.int32), t(pop["y"], np.int32), t(pop["status"], np.uint8),
                               t(pop["age"], np.uint8), t(pop["stratum"], np.uint8), t(pop["wealth"], np.float64))
    assert np.array_equal(host.run(6), dev.run(6))
    a, b = host._download(), dev._download()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    with pytest.raises(_native.NativeError):        # a host pointer is refused, not dereferenced on the device
        dev._handle.bind_device(4, {k: np.zeros(4, dtype=dt).ctypes.data for k, dt in _native.SOA_DTYPES})


def test_strided_snapshot_is_the_live_view_feed(native_lib):
    """graphics.py reads get_positions() + get_description() per frame; the strided snapshot is the same data for
    every stride-th agent of the population list."""
    pop, tw = _pop(50000, 500, 13)
    sim = _gpu(pop, tw, 500, seed=6)
    sim.run(5)
    full_pos, full_desc, full_c = sim.get_positions(), sim.get_description(), sim.get_description(complete=True)
    for stride in (1, 7, 1000):
        s = sim.snapshot(stride)
        assert len(s["x"]) == (50000 + stride - 1) // stride
        assert [[int(a), int(b)] for a, b in zip(s["x"], s["y"])] == full_pos[::stride]
    assert sim.get_positions(stride=7) == full_pos[::7]
    assert sim.get_description(stride=7) == full_desc[::7]
    assert sim.get_description(complete=True, stride=7) == full_c[::7]
