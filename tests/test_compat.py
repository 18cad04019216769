"""The drop-in seam: `covid_abs` imports resolve to this package (compat.install) and the reference's own scenario
callables (run_batch.py:8-50, 173-175 -- Python lambdas) are recognised and lowered to the kernel flags of their
declarative counterparts.  CPU only: nothing here launches a kernel."""
import sys

import numpy as np
import pytest


def test_compat_aliases_the_reference_module_tree():
    import covid19_agentbasedsimulation_b200.compat as compat
    had = {k: v for k, v in sys.modules.items() if k == "covid_abs" or k.startswith("covid_abs.")}
    compat.install()
    try:
        # what an unmodified user script does (run_batch.py:1-4)
        from covid_abs.abs import Simulation, MultiPopulationSimulation          # noqa: F401
        from covid_abs.agents import Status, InfectionSeverity, Agent            # noqa: F401
        from covid_abs.common import basic_income, lorenz_curve                  # noqa: F401
        from covid_abs.experiments import batch_experiment
        from covid_abs.network.graph_abs import GraphSimulation
        from covid_abs.network.util import new_day, bed_time, work_day           # noqa: F401
        from covid_abs.network.agents import EconomicalStatus, Person            # noqa: F401
        import covid19_agentbasedsimulation_b200.abs as ours
        import covid19_agentbasedsimulation_b200.experiments as ours_exp
        import covid19_agentbasedsimulation_b200.network.graph_abs as ours_graph
        assert Simulation is ours.Simulation and batch_experiment is ours_exp.batch_experiment
        assert GraphSimulation is ours_graph.GraphSimulation
        with pytest.raises(ImportError):
            import covid_abs.graphics  # noqa: F401   presentation code: refused loudly, never a silent CPU path
        sim = Simulation(population_size=30, length=12)
        assert sim.population_size == 30 and sim.contagion_rate == 0.9
    finally:
        compat.uninstall()
    now = {k: v for k, v in sys.modules.items() if k == "covid_abs" or k.startswith("covid_abs.")}
    assert now == had and not compat.installed()


def _reference_style_callables():
    """run_batch.py:8-50 and 173-175, restated (the script cannot be imported: it runs 14 400 iterations at import)."""
    from covid19_agentbasedsimulation_b200.network.util import new_day, bed_time
    from covid19_agentbasedsimulation_b200.network.agents import EconomicalStatus
    isolated = []

    def sleep(a):
        if not new_day(a.iteration) and bed_time(a.iteration):
            return True
        return False

    def lockdown(a):
        if a.house is not None:
            a.house.checkin(a)
        return True

    def conditional_lockdown(a):
        if a.environment.get_statistics()['Infected'] > .05:
            return lockdown(a)
        else:
            return False

    def sample_isolated(environment, isolation_rate=.5, list_isolated=isolated):
        for a in environment.population:
            test = np.random.rand()
            if test <= isolation_rate:
                list_isolated.append(a.id)

    def check_isolation(list_isolated, agent):
        if agent.id in list_isolated:
            agent.move_to_home()
            return True
        return False

    def vertical_isolation(a):
        if a.economical_status == EconomicalStatus.Inactive:
            if a.house is not None:
                a.house.checkin(a)
            return True
        return False

    def pset(x, property, value):
        x.__dict__[property] = value
        return False

    return dict(sleep=sleep, lockdown=lockdown, conditional_lockdown=conditional_lockdown,
                sample_isolated=sample_isolated, check_isolation=check_isolation, vertical_isolation=vertical_isolation,
                pset=pset, isolated=isolated)


def test_reference_lambdas_lower_like_the_named_behaviours():
    from covid19_agentbasedsimulation_b200.network.graph_abs import lower_callbacks
    from covid19_agentbasedsimulation_b200.network import scenarios as sc
    f = _reference_style_callables()
    sleep, lockdown, isolated = f["sleep"], f["lockdown"], f["isolated"]
    lam = {   # the callbacks entries of run_batch.py:95-213, verbatim
        "scenario1": {'on_execute': lambda x: sleep(x)},
        "scenario2": {'on_execute': lambda x: sleep(x), 'on_person_move': lambda x: lockdown(x)},
        "scenario3": {'on_execute': lambda x: sleep(x), 'on_person_move': lambda x: f["conditional_lockdown"](x)},
        "scenario4": {'on_execute': lambda x: sleep(x), 'on_person_move': lambda x: f["vertical_isolation"](x)},
        "scenario5": {'on_execute': lambda x: sleep(x),
                      'post_initialize': lambda x: f["sample_isolated"](x, isolation_rate=.4, list_isolated=isolated),
                      'on_person_move': lambda x: f["check_isolation"](isolated, x)},
        "scenario8": {'on_execute': lambda x: sleep(x), 'on_initialize': lambda x: f["pset"](x, 'contagion_rate', 0.1)},
        "scenario9": {'on_execute': lambda x: sleep(x),
                      'post_initialize': lambda x: f["sample_isolated"](x, isolation_rate=.6, list_isolated=isolated),
                      'on_person_move': lambda x: f["check_isolation"](isolated, x),
                      'on_initialize': lambda x: f["pset"](x, 'contagion_rate', 0.1)},
    }
    state = np.random.get_state()[1].copy()
    for name, callbacks in lam.items():
        want = lower_callbacks(getattr(sc, name)["callbacks"])
        got = lower_callbacks(callbacks)
        assert got == want, (name, got, want)
    assert np.array_equal(np.random.get_state()[1], state)      # probing consumed nothing from np.random


def test_unknown_callables_are_refused():
    from covid19_agentbasedsimulation_b200.network.graph_abs import lower_callbacks
    f = _reference_style_callables()

    def half_lockdown(a):                      # threshold other than the shipped .05
        if a.environment.get_statistics()['Infected'] > .10:
            return f["lockdown"](a)
        return False
    with pytest.raises(NotImplementedError):
        lower_callbacks({'on_person_move': half_lockdown})
    with pytest.raises(NotImplementedError):
        lower_callbacks({'post_execute': lambda x: None})
