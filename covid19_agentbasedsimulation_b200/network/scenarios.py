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


# ------------------------------------------------------------------------------------------------------------------
# Recognition of the reference's own callables.  A user script written against covid_abs passes Python lambdas
# (run_batch.py:87-213: `'on_execute': lambda x: sleep(x)`, `'on_person_move': lambda x: lockdown(x)`, ...).  A
# lambda cannot run on the device, but the behaviours the reference ships are few and each is a pure function of a
# handful of observable inputs, so a callable is identified by PROBING it with instrumented stand-ins and comparing
# what it does with the truth table of each named behaviour.  A callable that matches none raises
# NotImplementedError (there is no CPU fallback); one that matches is lowered to the same kernel flag as its
# Declarative counterpart above, so unmodified scenario dictionaries of run_batch.py work as they are.
import itertools

_PROBE_IDS = itertools.count(10 ** 12)


class _ProbeHouse(object):
    def __init__(self):
        self.checkins = 0

    def checkin(self, agent):
        self.checkins += 1


class _ProbeEnvironment(object):
    def __init__(self, infected):
        self.infected = infected

    def get_statistics(self, kind='info'):
        return {'Infected': self.infected}


class _ProbePerson(object):
    def __init__(self, infected=0.0, inactive=False, ident=0, homeless=False):
        from covid19_agentbasedsimulation_b200.network.agents import EconomicalStatus
        self.house = None if homeless else _ProbeHouse()
        self.environment = _ProbeEnvironment(infected)
        self.economical_status = EconomicalStatus.Inactive if inactive else EconomicalStatus.Active
        self.id = ident
        self.went_home = 0

    def move_to_home(self):
        self.went_home += 1


def _observe_move(fn, **kw):
    p = _ProbePerson(**kw)
    result = bool(fn(p))
    return result, (p.house.checkins if p.house is not None else 0), p.went_home


def _recognise_on_execute(fn):
    from covid19_agentbasedsimulation_b200.network.util import new_day, bed_time

    class _Clock(object):
        pass
    got, want = [], []
    for it in range(24 * 9):
        c = _Clock()
        c.iteration = it
        got.append(bool(fn(c)))
        want.append((not new_day(it)) and bed_time(it))     # run_batch.py:8-13
    if got == want:
        return sleep
    if not any(got):
        return None                                          # a callback that never skips an hour
    raise NotImplementedError("on_execute callable is not the `sleep` behaviour of run_batch.py:8-13")


def _recognise_on_person_move(fn, probed_isolated_ids):
    table = {}
    for infected in (0.0, 0.04, 0.06, 1.0):
        for inactive in (False, True):
            for ident in (-12345, -98766):       # the shipped behaviours do not depend on who is asked
                obs = _observe_move(fn, infected=infected, inactive=inactive, ident=ident)
                if table.setdefault((infected, inactive), obs) != obs:
                    raise NotImplementedError("on_person_move callable depends on the person's id: not a shipped "
                                              "behaviour (run_batch.py:16-50)")
    held = lambda obs: obs[0] and obs[1] == 1 and obs[2] == 0          # stays put, house charged once (run_batch.py:16-19)
    free = lambda obs: (not obs[0]) and obs[1] == 0 and obs[2] == 0
    if all(held(o) for o in table.values()):
        if not _observe_move(fn, homeless=True)[0]:
            raise NotImplementedError("lockdown-like callable that releases the homeless is not a shipped behaviour")
        return lockdown
    if all(free(o) for o in table.values()):
        # maybe check_isolation (run_batch.py:38-42): acts only on ids its post_initialize sampler listed
        for ident in probed_isolated_ids:
            r, checkins, home = _observe_move(fn, ident=ident)
            if r and home == 1 and checkins == 0:
                return check_isolation
        return None                                                      # a callback that never intervenes
    if all(held(table[(i, a)]) == (i > 0.05) for (i, a) in table):
        # run_batch.py:22-26: lockdown while Infected > .05 -- find the exact threshold by bisection
        lo, hi = 0.04, 0.06
        for _ in range(64):
            mid = 0.5 * (lo + hi)
            if held(_observe_move(fn, infected=mid)):
                hi = mid
            else:
                lo = mid
        if not (lo <= 0.05 <= hi) or held(_observe_move(fn, infected=0.05)):
            raise NotImplementedError("conditional lockdown with a threshold other than `Infected > .05` "
                                      "(run_batch.py:22-26) is not lowered to the device")
        return conditional_lockdown
    if all(held(table[(i, a)]) == a for (i, a) in table):
        return vertical_isolation                                        # run_batch.py:45-50
    raise NotImplementedError("on_person_move callable matches none of lockdown / conditional_lockdown / "
                              "vertical_isolation / check_isolation (run_batch.py:16-50)")


def recognise_callbacks(callbacks, sim=None):
    """callbacks dict with Python callables and/or Declarative entries -> dict of Declarative entries only.
    `sim`: the simulation the callbacks belong to (an `on_initialize` callable may read its attributes)."""
    import numpy as np
    out = {}
    callables = {e: a for e, a in (callbacks or {}).items() if not isinstance(a, Declarative)}
    for event, action in (callbacks or {}).items():
        if isinstance(action, Declarative):
            out[event] = action
    if not callables:
        return out
    for event, fn in callables.items():
        if not callable(fn):
            raise NotImplementedError("callback {!r} is neither callable nor Declarative".format(event))
    isolated_probe_ids = []
    if 'post_initialize' in callables:
        fn = callables['post_initialize']
        move = callables.get('on_person_move')
        if move is None:
            raise NotImplementedError("post_initialize=sample_isolated needs on_person_move=check_isolation "
                                      "(run_batch.py:136-146)")

        class _Agent(object):
            def __init__(self, ident):
                self.id = ident

        class _Env(object):
            def __init__(self, ident):
                self.population = [_Agent(ident)]

        real_rand = np.random.rand
        state = {'value': 0.0, 'calls': 0}

        def fake_rand(*shape):
            state['calls'] += 1
            return state['value'] if not shape else np.full(shape, state['value'])

        def is_listed(value):
            """Does a person whose draw is `value` end up isolated?  (seen through the move callback)"""
            ident = -next(_PROBE_IDS)      # never reused: run_batch.py shares one `isolated` list between scenarios
            state['value'] = value
            before = state['calls']
            np.random.rand = fake_rand
            try:
                fn(_Env(ident))
            finally:
                np.random.rand = real_rand
            if state['calls'] - before != 1:
                raise NotImplementedError("post_initialize callable is not `sample_isolated` (run_batch.py:31-35): it "
                                          "must draw np.random.rand() once per person")
            r, checkins, home = _observe_move(move, ident=ident)
            return bool(r and home == 1)

        if not is_listed(0.0):
            if is_listed(1.0):
                raise NotImplementedError("post_initialize callable is not `sample_isolated` (run_batch.py:31-35)")
            rate = None                     # never isolates anybody: equivalent to no callback at all
        else:
            lo, hi = 0.0, 1.0               # listed at lo; find the largest listed draw: `test <= isolation_rate`
            if is_listed(1.0):
                lo = 1.0
            else:
                for _ in range(80):
                    mid = 0.5 * (lo + hi)
                    if mid == lo or mid == hi:
                        break
                    if is_listed(mid):
                        lo = mid
                    else:
                        hi = mid
            rate = lo
        if rate is not None:
            out['post_initialize'] = sample_isolated(rate)
            out['on_person_move'] = check_isolation
        callables = {e: a for e, a in callables.items() if e not in ('post_initialize', 'on_person_move')}
    for event, fn in callables.items():
        if event == 'on_execute':
            d = _recognise_on_execute(fn)
        elif event == 'on_person_move':
            d = _recognise_on_person_move(fn, isolated_probe_ids)
        elif event == 'on_initialize':
            # graph_abs.py:94 hands the simulation itself to the callable (run_batch.py:173-175 `pset`): run it on a
            # stand-in that carries the simulation's attributes and keep whatever it assigned
            class _Bag(object):
                pass
            bag = _Bag()
            if sim is not None:
                bag.__dict__.update({k: v for k, v in sim.__dict__.items() if not k.startswith('_')})
            before = dict(bag.__dict__)
            fn(bag)
            changed = {k: v for k, v in bag.__dict__.items() if k not in before or before[k] is not v}
            d = Declarative("pset", overrides=changed) if changed else None
        else:
            raise NotImplementedError("callback event {!r} with a Python callable cannot run on the device "
                                      "(only on_execute / on_person_move / on_initialize / post_initialize behaviours "
                                      "of run_batch.py are lowered)".format(event))
        if d is not None:
            out[event] = d
    return out
