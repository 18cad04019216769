"""The scenario catalogue of run_batch.py restated declaratively.

run_batch.py configures GraphSimulation with Python lambdas (`callbacks={'on_execute': lambda x: sleep(x), ...}`,
run_batch.py:8-50,87-213) and cannot be imported (it runs 2 x 5 x 1 440 iterations at import).  Lambdas cannot run
on the device, so the same behaviours are named here and lowered to kernel flags:

    sleep                  run_batch.py:8-13    skip hours 1..7 of every day
    lockdown               run_batch.py:16-19   nobody moves; each executed hour charges expenses/720 to the house
    conditional_lockdown   run_batch.py:22-26   lockdown while the Infected fraction of the previous hour > .05
    vertical_isolation     run_batch.py:45-50   lockdown for the economically Inactive
    sample_isolated(rate)  run_batch.py:31-42   per-person isolation bit drawn at post_initialize; isolated persons
                                                 go home every executed hour (check_isolation)
    pset(attribute, value) run_batch.py:173-175 attribute override at on_initialize (masks: contagion_rate = .1)

`GraphSimulation(**{**global_parameters, **scenario2})` works exactly like run_batch.py:384-388.
"""
from covid19_agentbasedsimulation_b200.agents import Status


class Declarative(object):
    """A named behaviour that stands in for a run_batch.py callback."""

    def __init__(self, kind, **args):
        self.kind, self.args = kind, args

    def __repr__(self):
        return "Declarative({}, {})".format(self.kind, self.args)


sleep = Declarative("sleep")
lockdown = Declarative("lockdown")
conditional_lockdown = Declarative("conditional_lockdown")
vertical_isolation = Declarative("vertical_isolation")
check_isolation = Declarative("check_isolation")


def sample_isolated(isolation_rate=.5):
    return Declarative("sample_isolated", isolation_rate=float(isolation_rate))


def pset(attribute, value):
    return Declarative("pset", attribute=attribute, value=value)


global_parameters = dict(   # run_batch.py:53-85
    length=500, height=500,
    population_size=300, homemates_avg=3, homeless_rate=0.0005,
    amplitudes={Status.Susceptible: 10, Status.Recovered_Immune: 10, Status.Infected: 10},
    critical_limit=0.01, contagion_rate=.9, incubation_time=5, contagion_time=10, recovering_time=20,
    total_wealth=10000000, total_business=9, minimum_income=900.0, minimum_expense=600.0,
    public_gdp_share=0.1, business_gdp_share=0.5, unemployment_rate=0.12, business_distance=20)

_epidemic = dict(initial_infected_perc=.01, initial_immune_perc=.01, contagion_distance=1.)


def _isolation(rate):
    return {'on_execute': sleep, 'post_initialize': sample_isolated(rate), 'on_person_move': check_isolation}


scenario0 = dict(name='scenario0', initial_infected_perc=.0, initial_immune_perc=1., contagion_distance=1.)   # 87-93
scenario1 = dict(name='scenario1', callbacks={'on_execute': sleep}, **_epidemic)                               # 95-101
scenario2 = dict(name='scenario2', callbacks={'on_execute': sleep, 'on_person_move': lockdown}, **_epidemic)   # 103-112
scenario3 = dict(name='scenario3', callbacks={'on_execute': sleep, 'on_person_move': conditional_lockdown},
                 **_epidemic)                                                                                  # 114-123
scenario4 = dict(name='scenario4', callbacks={'on_execute': sleep, 'on_person_move': vertical_isolation},
                 **_epidemic)                                                                                  # 125-134
scenario5 = dict(name='scenario5', callbacks=_isolation(.4), **_epidemic)                                      # 136-146
scenario6 = dict(name='scenario6', callbacks=_isolation(.5), **_epidemic)                                      # 148-158
scenario7 = dict(name='scenario7', callbacks=_isolation(.6), **_epidemic)                                      # 160-170
_masks = dict(initial_infected_perc=.01, initial_immune_perc=.01, contagion_distance=.05)
scenario8 = dict(name='scenario8', callbacks={'on_execute': sleep, 'on_initialize': pset('contagion_rate', 0.1)},
                 **_masks)                                                                                     # 178-187
scenario9 = dict(name='scenario9', callbacks=dict(_isolation(.6), on_initialize=pset('contagion_rate', 0.1)),
                 **_masks)                                                                                     # 189-200
scenario10 = dict(name='scenario10', callbacks=dict(_isolation(.5), on_initialize=pset('contagion_rate', 0.1)),
                  **_masks)                                                                                    # 202-213


def partial_isolation(rate):
    """part_isol0 .. part_isol11 of run_batch.py:220-368 (isolation rates 0.1 .. 1.0)."""
    return dict(name='partialisolation{}'.format(rate), callbacks=_isolation(rate), **_epidemic)


# the seven scenarios of the paper, in run_batch.py naming (best-effort mapping)
PAPER_SCENARIOS = (scenario1, scenario2, scenario3, scenario4, scenario6, scenario8, scenario10)
