"""Host-side logic that needs no GPU: batch_experiment contract, parameter lowering, trigger validation."""
import numpy as np
import pandas as pd
import pytest

from covid19_agentbasedsimulation_b200 import abs as pabs
from covid19_agentbasedsimulation_b200.agents import Status, InfectionSeverity, Agent
from covid19_agentbasedsimulation_b200.experiments import batch_experiment


class FakeSim(object):
    """Duck type batch_experiment drives (experiments.py:185-194)."""
    calls = 0

    def __init__(self, **kw):
        self.kw = kw
        self.t = 0
        FakeSim.calls += 1
        self.offset = FakeSim.calls
        self.fail_at = kw.get("fail_at")

    def initialize(self):
        pass

    def execute(self):
        self.t += 1
        if self.fail_at is not None and self.t == self.fail_at and self.offset % 2 == 0:
            raise RuntimeError("boom")

    def get_statistics(self, kind='info'):
        return {"A": float(self.t + self.offset), "B": float(self.t * 2)}


def test_batch_experiment_schema_and_aggregation(tmp_path):
    FakeSim.calls = 0
    f = tmp_path / "out.csv"
    df = batch_experiment(4, 3, str(f), simulation_type=FakeSim)
    assert list(df.columns) == ['Iteration', 'Metric', 'Min', 'Avg', 'Std', 'Max']
    assert len(df) == 3 * 2
    assert df["Metric"].tolist() == ["A", "B"] * 3          # key order of get_statistics is the row order
    row = df[(df.Iteration == 0) & (df.Metric == "A")].iloc[0]
    vals = np.array([1 + k for k in (1, 2, 3, 4)], dtype=float)   # t=1 after the first execute (experiments.py:193)
    assert row.Min == vals.min() and row.Max == vals.max() and row.Avg == vals.mean()
    assert row.Std == vals.std()                             # ddof = 0 (experiments.py:206)
    back = pd.read_csv(f)
    assert back.shape == df.shape


class FakeAggregating(FakeSim):
    """A simulation type that offers the device-side reduction hook (network/graph_abs.py::run_ensemble_aggregated)."""
    broken = False

    @classmethod
    def run_ensemble_aggregated(cls, experiments, iterations, **kw):
        if cls.broken:
            raise RuntimeError("no device")
        t = np.arange(1, iterations + 1, dtype=np.float64)
        agg = np.empty((iterations, 2, 4))
        agg[:, 0] = np.stack([t + 1, t + 2.5, np.full_like(t, 0.5), t + 4], axis=1)
        agg[:, 1] = np.stack([2 * t, 2 * t, np.zeros_like(t), 2 * t], axis=1)
        return ["A", "B"], agg


def test_batch_experiment_takes_the_device_aggregate_when_offered(tmp_path, capsys):
    FakeAggregating.broken = False
    df = batch_experiment(4, 3, str(tmp_path / "a.csv"), simulation_type=FakeAggregating)
    assert list(df.columns) == ['Iteration', 'Metric', 'Min', 'Avg', 'Std', 'Max']
    assert df["Metric"].tolist() == ["A", "B"] * 3 and df["Iteration"].tolist() == [0, 0, 1, 1, 2, 2]
    assert df.iloc[2].tolist() == [1, "A", 3.0, 4.5, 0.5, 6.0]
    assert pd.read_csv(tmp_path / "a.csv").shape == df.shape
    # a failing device path falls back to one experiment after the other (the reference's own loop)
    FakeAggregating.broken = True
    FakeSim.calls = 0
    df = batch_experiment(4, 3, str(tmp_path / "b.csv"), simulation_type=FakeAggregating)
    assert "aggregated ensemble run" in capsys.readouterr().out
    row = df[(df.Iteration == 0) & (df.Metric == "A")].iloc[0]
    assert row.Min == 2.0 and row.Max == 5.0


def test_batch_experiment_swallows_failed_experiments(tmp_path, capsys):
    FakeSim.calls = 0
    df = batch_experiment(4, 4, str(tmp_path / "o.csv"), simulation_type=FakeSim, fail_at=3)
    assert "Exception occurred in experiment" in capsys.readouterr().out
    assert len(df) == 4 * 2   # iterations 2,3 aggregate only the surviving experiments


def test_amplitude_array_and_scale_bits():
    amp = pabs.amplitude_array({Status.Susceptible: 5, Status.Recovered_Immune: 1.5, Status.Infected: 0})
    assert amp.tolist() == [5.0, 0.0, 1.5, 0.0]
    from oracle.sync_oracle import default_wealth_scale_bits
    for tw in (1, 1e4, 5e9, 1e7, 3.3e12, 1e18):
        assert pabs.wealth_scale_bits(tw) == default_wealth_scale_bits(tw)
        assert 1024 * tw * 2.0 ** pabs.wealth_scale_bits(tw) < 2.0 ** 62 or pabs.wealth_scale_bits(tw) == 0


def test_trigger_validation():
    pabs.Trigger('Infected', '>=', .2, 'amplitudes', {Status.Susceptible: 1})
    with pytest.raises(ValueError):
        pabs.Trigger('Nope', '>=', .2, 'amplitudes', {})
    with pytest.raises(ValueError):
        pabs.Trigger('Infected', '==', .2, 'amplitudes', {})


def test_enums_match_reference_encodings():
    assert [s.value for s in Status] == ['s', 'i', 'c', 'm']
    assert [s.name for s in Status] == ['Susceptible', 'Infected', 'Recovered_Immune', 'Death']
    assert [s.value for s in InfectionSeverity] == ['e', 'a', 'h', 'g']
    a = Agent(status=Status.Infected)
    assert a.get_description() == "Infected(Asymptomatic)" and str(a) == "Infected"


def test_constructor_defaults_match_reference_kwargs():
    s = pabs.Simulation()
    assert (s.population_size, s.length, s.height) == (20, 10, 10)
    assert (s.initial_infected_perc, s.initial_immune_perc) == (0.05, 0.05)
    assert (s.contagion_distance, s.contagion_rate, s.critical_limit) == (1., 0.9, 0.6)
    assert s.amplitudes == {Status.Susceptible: 5, Status.Recovered_Immune: 5, Status.Infected: 5}
    assert (s.minimum_income, s.minimum_expense, s.total_wealth) == (1.0, 1.0, 10 ** 4)
    assert s.statistics is None and s.triggers_simulation == [] and s.triggers_population == []


def test_constants_match_oracle_tables():
    from covid19_agentbasedsimulation_b200 import common
    from oracle import seq_oracle
    assert common.age_hospitalization_probs == seq_oracle.AGE_HOSP
    assert common.age_severe_probs == seq_oracle.AGE_SEVERE
    assert common.age_death_probs == seq_oracle.AGE_DEATH
    assert common.basic_income.tolist() == seq_oracle.BASIC_INCOME.tolist() == [1.0, 2.0, 3.25, 5.0, 13.75]


def test_shard_block_and_interleaved():
    """network/ensemble.py::shard: every job exactly once; interleaved keeps a scenario-major list scenario-major."""
    from covid19_agentbasedsimulation_b200.network.ensemble import shard
    assert [list(shard(10, r, 4)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert [list(shard(10, r, 4, interleaved=True)) for r in range(4)] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    jobs = [(s, k) for s in range(7) for k in range(16)]
    for world in (1, 2, 3, 8):
        parts = [[jobs[j] for j in shard(len(jobs), r, world, interleaved=True)] for r in range(world)]
        assert sorted(sum(parts, [])) == jobs
        for mine in parts:
            assert mine == sorted(mine)                                  # the seeds of a scenario stay consecutive
            assert max(sum(1 for s, _ in mine if s == sc) for sc in range(7)) <= -(-16 // world) + 1
