"""Pins the sequential oracle to the golden vectors generated from the unmodified reference
(oracle/gen_golden.py -> tests/golden/*.json).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle.seq_oracle import SeqSimulation, FrozenDraws, S, I, R, D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def check_snapshot(sim, snap, tag):
    assert sim.x.tolist() == snap["x"], tag
    assert sim.y.tolist() == snap["y"], tag
    assert sim.status.tolist() == snap["status"], tag
    assert sim.severity.tolist() == snap["severity"], tag
    assert sim.infected_time.tolist() == snap["infected_time"], tag
    assert sim.age.tolist() == snap["age"], tag
    assert sim.stratum.tolist() == snap["stratum"], tag
    assert [float(w).hex() for w in sim.wealth] == snap["wealth_hex"], tag
    stats = sim.get_statistics('all')
    assert list(stats.keys()) == snap["stat_keys"], tag
    for k in snap["stat_keys"]:
        assert float(stats[k]).hex() == snap["stats"][k], (tag, k)


@pytest.mark.parametrize("case", load("simulation_kat.json"), ids=lambda c: "%s-seed%d" % (c["name"], c["seed"]))
def test_known_answer_vectors(case):
    """np.random.seed(s); initialize(); execute() x k -- every field, pair list and statistic bit-identical."""
    np.random.seed(case["seed"])
    sim = SeqSimulation(**case["kwargs"])
    sim.initialize()
    check_snapshot(sim, case["initial"], "initial")
    sim.statistics = None
    last = max(int(k) for k in case["after"])
    for it in range(1, last + 1):
        sim.execute()
        if str(it) in case["after"]:
            snap = case["after"][str(it)]
            check_snapshot(sim, snap, "after %d" % it)
            assert [list(p) for p in sim.contact_pairs()] == snap["pairs"]
            sim.statistics = None


def test_survey_appendix_c1_vectors():
    """SURVEY.md appendix C.1 (seed 0, defaults)."""
    np.random.seed(0)
    sim = SeqSimulation()
    sim.initialize()
    assert list(zip(sim.x.tolist(), sim.y.tolist()))[:5] == [(10, 6), (9, 3), (6, 4), (2, 0), (10, 9)]
    assert sim.age.tolist() == [22, 37, 44, 38, 47, 41, 5, 20, 19, 23, 63, 15, 11, 17, 29, 5, 12, 65, 27, 44]
    assert sim.status_string() == "icssssssssssssssssss"
    sim.execute()
    assert sim.last_contacts == [(1, 4), (1, 15), (3, 16), (4, 15), (9, 10), (12, 17)]
    st = sim.get_statistics('all')
    assert float(st["Q1"]) == 398.8894520966977 and float(st["Q5"]) == 5558.631757041177
    for _ in range(29):
        sim.execute()
    assert sim.status_string() == "ccsiicciciisiicmicis"
    assert sim.infected_time.tolist() == [0, 0, 0, 20, 19, 0, 0, 16, 0, 19, 19, 0, 15, 18, 0, 9, 18, 0, 20, 0]


@pytest.mark.parametrize("case", load("simulation_frozen.json"), ids=lambda c: "seed_agent%d" % c["seed_agent"])
def test_frozen_draw_fixtures(case):
    """sequential in-step chains under rate=1, frozen positions and draws."""
    n = 6
    status = [I if i == case["seed_agent"] else S for i in range(n)]
    sim = SeqSimulation(draws=FrozenDraws(random=0.999999), population_size=n, length=100, height=100,
                        contagion_rate=1.0, amplitudes={S: 0, R: 0, I: 0})
    sim.set_population(10 + np.arange(n), np.full(n, 10), status, np.full(n, 30), np.full(n, 2), np.full(n, 100.0))
    for step in case["steps"]:
        sim.execute()
        assert sim.status_string() == step["status"]
        assert sim.infected_time.tolist() == step["infected_time"]
        assert sim.wealth.tolist() == step["wealth"]


def test_ensemble_golden_reproduced_by_oracle():
    """The oracle under the same seeds gives the same ensemble curves (first 32 seeds recomputed exactly)."""
    ens = load("simulation_ensemble.json")["defaults"]
    keys = ens["keys"]
    seeds = ens["seeds"][:32]
    rows = []
    for s in seeds:
        np.random.seed(s)
        sim = SeqSimulation(**ens["kwargs"])
        sim.initialize()
        sim.get_statistics('all')
        run = []
        for _ in range(ens["iterations"]):
            sim.execute()
            st = sim.get_statistics('all')
            run.append([float(st[k]) for k in keys])
        rows.append(run)
    arr = np.array(rows)
    final = np.array(ens["final"])[:32]
    np.testing.assert_array_equal(arr[:, -1, :], final)
