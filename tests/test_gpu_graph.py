"""GraphSimulation on the GPU (through the C ABI) against the oracle and the reference's golden trajectories."""
import glob
import json
import os

import numpy as np
import pytest

from graph_util import product_world, oracle_world, sim_kwargs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "graph_*.json")))


def _compare_state(sim, orc):
    w = sim.world()
    for k, o in (("px", orc.px), ("py", orc.py), ("status", orc.status), ("severity", orc.severity),
                 ("infected_time", orc.time), ("employer", orc.employer), ("hire_seq", orc.hire_seq),
                 ("hsize", orc.hsize), ("bstocks", orc.bstocks), ("bsales", orc.bsales), ("bnum_employees", orc.bnum)):
        assert np.array_equal(np.asarray(w[k]).astype(np.int64), o), k
    for k, o in (("pwealth", orc.pwealth), ("hwealth", orc.hwealth), ("hfixed", orc.hfixed), ("hincomes", orc.hincomes),
                 ("hexpenses", orc.hexpenses), ("bwealth", orc.bwealth), ("bfixed", orc.bfixed),
                 ("bincomes", orc.bincomes)):
        assert np.array_equal(np.rint(np.asarray(w[k]) * orc.scale).astype(np.int64), o), k
    assert int(round(w["gov_wealth"] * orc.scale)) == orc.gov["wealth"]
    assert int(round(w["health_wealth"] * orc.scale)) == orc.health["wealth"]
    assert int(round(w["health_expenses"] * orc.scale)) == orc.health["expenses"]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-5] for p in GOLDEN])
def test_reference_trajectories(native_lib, path):
    """The fixtures were produced by the unmodified reference under keyed draws: same epidemiological rows exactly,
    money rows to fixed-point rounding; and bit equality with the oracle in every column."""
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation, GRAPH_STAT_KEYS
    from oracle.graph_oracle import GraphOracle
    fix = json.load(open(path))
    p = fix["world"]["params"]
    sim = GraphSimulation(seed=fix["seed"], history_capacity=1024, **sim_kwargs(p))
    sim.set_world(product_world(fix["world"]), mode=p["mode"], sleep=p["sleep"])
    init = sim.get_statistics('all')
    orc = GraphOracle(fix["world"], seed=fix["seed"])
    assert [float(init[k]) for k in GRAPH_STAT_KEYS] == [orc.stats[k] for k in GRAPH_STAT_KEYS]
    rows = sim.run(fix["iterations"])
    gold = np.array(fix["rows"])
    assert np.array_equal(rows[:, 7:], gold[:, 7:])
    assert np.allclose(rows[:, :7], gold[:, :7], rtol=0, atol=1e-9)
    for it in range(fix["iterations"]):
        st = orc.execute()
        assert rows[it].tolist() == [st[k] for k in GRAPH_STAT_KEYS], it
    _compare_state(sim, orc)
    assert sim.world()["px"].tolist() == fix["final"]["px"] and sim.world()["status"].tolist() == fix["final"]["status"]


@pytest.mark.parametrize("scenario,n", [("scenario1", 1500), ("scenario2", 1000), ("scenario3", 1000), ("scenario6", 800)])
def test_built_world_stepwise_equals_batched_equals_oracle(native_lib, scenario, n):
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation, GRAPH_STAT_KEYS, default_money_bits
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    from oracle.graph_oracle import GraphOracle
    kw = dict(sc.global_parameters)
    kw.update(getattr(sc, scenario))
    kw.pop("name")
    kw.update(population_size=n, initial_infected_perc=0.03, length=300, height=300)
    np.random.seed(5)
    a = GraphSimulation(seed=99, **kw)
    world, mode, sleep = a.build_world()
    a.set_world(world, mode, sleep)
    b = GraphSimulation(seed=99, **kw)
    b.set_world(world, mode, sleep)
    amp = [10.0, 10.0, 10.0, 0.0]
    params = dict(population_size=n, total_wealth=kw["total_wealth"], contagion_distance=kw["contagion_distance"],
                  contagion_rate=a.contagion_rate, critical_limit=kw["critical_limit"], amplitudes=amp,
                  minimum_income=kw["minimum_income"], minimum_expense=kw["minimum_expense"],
                  business_distance=kw["business_distance"], incubation_time=5, contagion_time=10, recovering_time=20,
                  mode=mode, sleep=sleep, money_bits=default_money_bits(kw["total_wealth"]))
    orc = GraphOracle(oracle_world(world, params), seed=99)
    iters = 60
    batched = a.run(iters)
    for it in range(iters):
        b.execute()
        s = b.get_statistics('all')
        o = orc.execute()
        assert [float(s[k]) for k in GRAPH_STAT_KEYS] == batched[it].tolist() == [o[k] for k in GRAPH_STAT_KEYS], it
    _compare_state(a, orc)
    assert b.iteration == a.iteration == iters - 1


@pytest.mark.parametrize("variant", ["distance3", "distance2", "distance_half", "scarce_stocks", "plentiful_stocks", "crowded_6000", "peak_filter"])
def test_contact_and_stock_paths_equal_oracle(native_lib, variant):
    """The branches the shipped scenarios do not reach: cells larger than one lattice point (reach 2 and 3: the nearest
    point of a cell decides whether it is looked up), same-point-only contacts (distance < 1), businesses whose opening
    stock does not cover the hour's requests (the ordered max-plus scan instead of the one-step settlement) or always
    does, the
    peak of an epidemic (sources filtered by the cells the last Susceptible persons can reach), and a
    crowded world (several groups per thread, the contact table beyond its smallest size, piles counted in shared
    memory) -- every column and every statistics row against the oracle."""
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation, GRAPH_STAT_KEYS, default_money_bits
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    from oracle.graph_oracle import GraphOracle
    kw = dict(sc.global_parameters)
    kw.update(sc.scenario1)
    kw.pop("name")
    n, side, iters = 1200, 120, 72
    kw.update(initial_infected_perc=0.05)
    if variant == "distance3":
        kw.update(contagion_distance=3.0, contagion_rate=0.3)
    elif variant == "distance2":
        kw.update(contagion_distance=2.0, contagion_rate=0.4)
    elif variant == "distance_half":
        kw.update(contagion_distance=0.5)
    elif variant == "scarce_stocks":
        kw.update(business_distance=45)
    elif variant == "crowded_6000":
        n, side, iters = 6000, 160, 30
    elif variant == "peak_filter":          # nine days: from day 5 on the sources outnumber the Susceptible 5 : 1 and
        n, side, iters = 800, 100, 216      # only those a Susceptible person can reach enter the contact table
    kw.update(population_size=n, length=side, height=side)
    np.random.seed(11)
    a = GraphSimulation(seed=17, **kw)
    world, mode, sleep = a.build_world()
    if variant == "scarce_stocks":
        world["bstocks"][:] = 0
    if variant == "plentiful_stocks":       # every request of every hour is covered: the one-step settlement only
        world["bstocks"][:] = 1000000
    a.set_world(world, mode, sleep)
    params = dict(population_size=n, total_wealth=kw["total_wealth"], contagion_distance=kw["contagion_distance"],
                  contagion_rate=a.contagion_rate, critical_limit=kw["critical_limit"], amplitudes=[10.0, 10.0, 10.0, 0.0],
                  minimum_income=kw["minimum_income"], minimum_expense=kw["minimum_expense"],
                  business_distance=kw["business_distance"], incubation_time=5, contagion_time=10, recovering_time=20,
                  mode=mode, sleep=sleep, money_bits=default_money_bits(kw["total_wealth"]))
    orc = GraphOracle(oracle_world(world, params), seed=17)
    rows = a.run(iters)
    for it in range(iters):
        o = orc.execute()
        assert rows[it].tolist() == [o[k] for k in GRAPH_STAT_KEYS], (variant, it)
    _compare_state(a, orc)
    if variant == "scarce_stocks":
        assert int(np.asarray(a.world()["bsales"]).sum()) > 0      # requests were served out of checked-in stock


def test_entities_are_materialised_like_the_reference_lists(native_lib):
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    kw = dict(sc.global_parameters)
    kw.update(sc.scenario1)
    kw.pop("name")
    np.random.seed(1)
    sim = GraphSimulation(seed=3, **kw)
    sim.initialize()
    sim.run(30)
    assert len(sim.population) == 300 and len(sim.houses) >= 100 and len(sim.business) == 9
    assert sum(len(b.employees) for b in sim.business) == sum(1 for p in sim.population if p.employer is not None)
    assert sim.government.price == 1.0 and sim.healthcare is not None
    assert set(sim.get_statistics('ecom')) == {"Q1", "Q2", "Q3", "Q4", "Q5", "Business", "Government"}
    assert len(sim.get_positions()) == 300


def test_ensemble_handle_equals_individual_runs_and_batch_experiment(native_lib, tmp_path):
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation, GRAPH_STAT_KEYS
    from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble, shard
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    from covid19_agentbasedsimulation_b200.experiments import batch_experiment
    assert [list(shard(10, r, 4)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert [list(shard(10, r, 4, interleaved=True)) for r in range(4)] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    names = ("scenario1", "scenario2", "scenario3", "scenario4", "scenario6", "scenario8", "scenario10")
    kws = []
    for name in names:
        kw = dict(sc.global_parameters)
        kw.update(getattr(sc, name))
        kw.pop("name")
        kw.update(population_size=120, initial_infected_perc=0.05)
        kws.append(kw)
    np.random.seed(2)
    sims = [GraphSimulation(seed=100 + k, **kw) for k, kw in enumerate(kws)]
    worlds = [s.build_world() for s in sims]
    ens = GraphEnsemble(sims)
    ens.initialize(worlds)
    table = ens.run(100)
    assert table.shape == (7, 100, 14)
    for k, kw in enumerate(kws):
        one = GraphSimulation(seed=100 + k, **kw)
        one.set_world(*worlds[k])                  # on_initialize overrides (masks: contagion_rate) apply here too
        assert one.contagion_rate == (0.1 if names[k] in ("scenario8", "scenario10") else 0.9)
        assert np.array_equal(one.run(100), table[k]), names[k]
    assert not np.array_equal(table[0], table[1])
    # batch_experiment drives the same ensemble path and keeps the reference's CSV schema (experiments.py:212)
    kw = dict(kws[0])
    np.random.seed(4)
    df = batch_experiment(6, 50, str(tmp_path / "g.csv"), simulation_type=GraphSimulation, seed=7, **kw)
    assert list(df.columns) == ['Iteration', 'Metric', 'Min', 'Avg', 'Std', 'Max']
    assert len(df) == 50 * 14 and list(df['Metric'][:14]) == list(GRAPH_STAT_KEYS)
    inf = df[df['Metric'] == 'Infected']
    assert (inf['Min'] <= inf['Avg']).all() and (inf['Avg'] <= inf['Max']).all() and inf['Std'].max() > 0


def test_device_aggregate_equals_numpy_over_the_rows(native_lib):
    """cabs_graph_aggregate = the Min/Avg/Std/Max loop of batch_experiment (experiments.py:200-208) over the handle's
    simulations.  Min and Max are exact; Avg and Std are float64 reductions in a different (fixed) order than numpy's
    pairwise sum: tolerance 1e-13 relative for Avg, 1e-10 for Std (a sum of squared differences)."""
    from covid19_agentbasedsimulation_b200.network.graph_abs import GraphSimulation
    from covid19_agentbasedsimulation_b200.network.ensemble import GraphEnsemble
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    kw = dict(sc.global_parameters)
    kw.update(sc.scenario2)
    kw.pop("name")
    kw.update(population_size=90, initial_infected_perc=0.05)
    n_sims, iterations = 37, 60          # 37: lanes with one and with two simulations each
    np.random.seed(11)
    world = GraphSimulation(seed=0, **kw).build_world()
    tables = []
    for cap in (2048, 16):               # 16: the history ring wraps, the run proceeds in chunks
        ens = GraphEnsemble([GraphSimulation(seed=500 + k, **kw) for k in range(n_sims)], history_capacity=cap)
        ens.initialize(world)
        if cap == 2048:
            table = ens.run(iterations)
            agg = ens.handle.aggregate(0, iterations)
            first = ens.handle.aggregate(-1, 3)       # starts at the initial state
            init = np.stack([ens.handle.stats(k, -1, 1)[0] for k in range(n_sims)])
        else:
            agg = ens.run_aggregated(iterations)
            ens.close()
            ens = GraphEnsemble([GraphSimulation(seed=500 + k, **kw) for k in range(n_sims)], history_capacity=cap)
            ens.initialize(world)
            assert np.array_equal(ens.run(iterations), table)     # strided copies of a wrapped ring
        tables.append(agg)
        ens.close()
    assert np.array_equal(tables[0], tables[1])
    agg = tables[0]
    assert agg.shape == (iterations, 14, 4)
    assert np.array_equal(agg[:, :, 0], table.min(axis=0)) and np.array_equal(agg[:, :, 3], table.max(axis=0))
    assert np.allclose(agg[:, :, 1], table.mean(axis=0), rtol=1e-13, atol=1e-300)
    assert np.allclose(agg[:, :, 2], table.std(axis=0), rtol=1e-10, atol=1e-15)
    assert agg[:, 8, 2].max() > 0                      # the seeds do differ (Infected)
    assert np.array_equal(first[0, :, 0], init.min(axis=0)) and np.array_equal(first[1:], agg[:2])
