"""Declarative per-agent rules: the device form of `triggers_population` (covid_abs/abs.py:86-95).

The reference takes Python lambdas:

    sim.append_trigger_population(lambda a: a.status == Status.Infected, 'move',
                                  lambda a: (a.x + np.random.normal(0, .01, 1), a.y + np.random.normal(0, .01, 1)))
    sim.append_trigger_population(lambda a: a.age > 60, 'wealth', lambda w: w * 0.99)

A 'move' entry replaces the random walk of the agents its condition selects (abs.py:169-172), any other entry rewrites
an attribute after `update()` (abs.py:244-247).  A lambda cannot run on the device; what the device takes is

    MoveRule(when, target='stay' | (x, y), sigma)       new position = target + N(0, sigma) per axis
    AttributeRule(when, attribute, value=... | scale=..., shift=...)

where `when` is a `When(...)` (conjunction of attribute sets) or ANY predicate of an agent's status, infected_status,
age and social_stratum given as a callable -- it is tabulated over those 8 000 combinations, so nothing about its form
matters.  `lower_trigger` also accepts the reference's own dict entries and recognises their shape by probing the
lambdas with instrumented stand-ins (the condition by tabulation; a move action by reading the arguments it passes to
np.random.normal; an attribute action by fitting new = p * old + q); entries that do not fit raise NotImplementedError --
there is no CPU fallback.
"""
import numpy as np

from covid19_agentbasedsimulation_b200 import _native
from covid19_agentbasedsimulation_b200.agents import Status, InfectionSeverity, STATUS_LIST, SEVERITY_LIST, STATUS_CODE, SEVERITY_CODE

N_STATUS, N_SEVERITY, N_AGE, N_STRATUM = 4, 4, 100, 5


class _ProbeAgent(object):
    """Stand-in handed to a condition / action while it is tabulated."""

    def __init__(self, status, severity, age, stratum, x=3.0, y=4.0, infected_time=0, wealth=100.0):
        self.status, self.infected_status = STATUS_LIST[status], SEVERITY_LIST[severity]
        self.age, self.social_stratum = age, stratum
        self.x, self.y, self.infected_time, self.wealth = x, y, infected_time, wealth
        self.id = 0


class When(object):
    """Conjunction of attribute sets; None = any.  `status` / `infected_status`: a member or an iterable of members."""

    def __init__(self, status=None, infected_status=None, min_age=None, max_age=None, social_stratum=None):
        def as_set(v, codes):
            if v is None:
                return None
            v = [v] if not isinstance(v, (list, tuple, set, frozenset)) else list(v)
            return frozenset(codes[m] if not isinstance(m, int) else m for m in v)
        self.status = as_set(status, STATUS_CODE)
        self.severity = as_set(infected_status, SEVERITY_CODE)
        self.min_age, self.max_age = min_age, max_age
        self.stratum = None if social_stratum is None else frozenset(
            [social_stratum] if isinstance(social_stratum, int) else social_stratum)

    def __call__(self, a):
        return ((self.status is None or STATUS_CODE[a.status] in self.status) and
                (self.severity is None or SEVERITY_CODE[a.infected_status] in self.severity) and
                (self.min_age is None or a.age >= self.min_age) and (self.max_age is None or a.age <= self.max_age) and
                (self.stratum is None or a.social_stratum in self.stratum))


def tabulate(condition):
    """8 000-bit table of a predicate of (status, infected_status, age, social_stratum), as uint32 words.  Raises
    NotImplementedError if the predicate also depends on something else the probe can vary (position, wealth,
    infected_time)."""
    bits = np.zeros(N_STATUS * N_SEVERITY * N_AGE * N_STRATUM, dtype=bool)
    k = 0
    for st in range(N_STATUS):
        for sev in range(N_SEVERITY):
            for age in range(N_AGE):
                for stratum in range(N_STRATUM):
                    bits[k] = bool(condition(_ProbeAgent(st, sev, age, stratum)))
                    k += 1
    # the table is all the device knows: make sure nothing else moves the answer
    rng = np.random.RandomState(12345)
    for _ in range(400):
        st, sev, age, stratum = rng.randint(N_STATUS), rng.randint(N_SEVERITY), rng.randint(N_AGE), rng.randint(N_STRATUM)
        other = _ProbeAgent(st, sev, age, stratum, x=float(rng.randn() * 50), y=float(rng.randn() * 50),
                            infected_time=int(rng.randint(0, 40)), wealth=float(rng.rand() * 1e4))
        if bool(condition(other)) != bits[((st * N_SEVERITY + sev) * N_AGE + age) * N_STRATUM + stratum]:
            raise NotImplementedError("population trigger condition depends on position, wealth or infected_time: only "
                                      "predicates of status, infected_status, age and social_stratum are lowered")
    words = np.zeros(_native.RULE_CONDITION_WORDS, dtype=np.uint32)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return words


class MoveRule(object):
    """`{'condition': when, 'attribute': 'move', 'action': lambda a: target + N(0, sigma)}` (abs.py:169-172)."""

    def __init__(self, when, target='stay', sigma=0.0):
        self.when, self.target, self.sigma = when, target, float(sigma)
        if target != 'stay':
            self.target = (float(target[0]), float(target[1]))

    def native(self):
        r = _native.Rule()
        r.kind = _native.RULE_MOVE_STAY if self.target == 'stay' else _native.RULE_MOVE_POINT
        if self.target != 'stay':
            r.target_x, r.target_y = self.target
        r.sigma = self.sigma
        r.condition[:] = tabulate(self.when).tolist()
        return r

    def key(self):
        return ("move", self.target, self.sigma, tabulate(self.when).tobytes())


class AttributeRule(object):
    """`{'condition': when, 'attribute': name, 'action': lambda old: new}` (abs.py:244-247).  Discrete attributes
    (status, infected_status, infected_time, age, social_stratum) are set to `value`; wealth becomes
    scale * wealth + shift."""

    def __init__(self, when, attribute, value=None, scale=1.0, shift=0.0):
        if attribute not in _native.RATTR:
            raise NotImplementedError("population trigger on attribute {!r} is not lowered to the device".format(attribute))
        self.when, self.attribute = when, attribute
        if attribute == "wealth":
            if value is not None:
                scale, shift = 0.0, float(value)
            self.p, self.q = float(scale), float(shift)
        else:
            if value is None:
                raise ValueError("a rule on {} needs value=".format(attribute))
            if isinstance(value, Status):
                value = STATUS_CODE[value]
            elif isinstance(value, InfectionSeverity):
                value = SEVERITY_CODE[value]
            self.p, self.q = 0.0, float(int(value))

    def native(self):
        r = _native.Rule()
        r.kind, r.attribute = _native.RULE_ATTRIBUTE, _native.RATTR[self.attribute]
        r.p, r.q = self.p, self.q
        r.condition[:] = tabulate(self.when).tolist()
        return r

    def key(self):
        return ("attr", self.attribute, self.p, self.q, tabulate(self.when).tobytes())


def _recognise_move_action(action):
    """A move action in the reference's idiom: (px + np.random.normal(0, s, 1), py + np.random.normal(0, s, 1)) with
    (px, py) the agent's own position or a fixed point.  Found by calling it with np.random.normal replaced by a
    recorder that returns zeros."""
    real = np.random.normal
    calls = []

    def recorder(loc=0.0, scale=1.0, size=None):
        calls.append((float(loc), float(scale)))
        return np.zeros(size) if size is not None else 0.0

    results = []
    np.random.normal = recorder
    try:
        for x, y in ((3.0, 4.0), (-7.5, 11.25), (100.0, 0.0)):
            calls.clear()
            out = action(_ProbeAgent(0, 1, 30, 2, x=x, y=y))
            ox, oy = (float(np.asarray(v).ravel()[0]) for v in out)
            results.append(((x, y), (ox, oy), list(calls)))
    finally:
        np.random.normal = real
    if any(len(c) != 2 or c[0][0] != 0.0 or c[1][0] != 0.0 or c[0][1] != c[1][1] for _, _, c in results):
        raise NotImplementedError("move trigger action is not `point + np.random.normal(0, sigma)` per axis (one draw "
                                  "per axis, same sigma)")
    sigma = results[0][2][0][1]
    if all(o == p for p, o, _ in results):
        return 'stay', sigma
    if all(o == results[0][1] for _, o, _ in results):
        return results[0][1], sigma
    raise NotImplementedError("move trigger action targets neither the agent's own position nor a fixed point")


def lower_trigger(trigger):
    """A `triggers_population` entry -> MoveRule / AttributeRule (reference dict entries are recognised by probing)."""
    if isinstance(trigger, (MoveRule, AttributeRule)):
        return trigger
    if not isinstance(trigger, dict) or not {'condition', 'attribute', 'action'} <= set(trigger):
        raise NotImplementedError("unsupported population trigger {!r}".format(trigger))
    cond, attr, action = trigger['condition'], trigger['attribute'], trigger['action']
    if attr == 'move':
        target, sigma = _recognise_move_action(action)
        return MoveRule(cond, target, sigma)
    if attr not in _native.RATTR:
        raise NotImplementedError("population trigger on attribute {!r} is not lowered to the device".format(attr))
    if attr == 'wealth':      # fit new = p * old + q on three probes, check on a fourth
        v = [float(np.asarray(action(x)).ravel()[0]) for x in (0.0, 1.0, 10.0, -3.5)]
        p, q = v[1] - v[0], v[0]
        if abs(p * 10.0 + q - v[2]) > 1e-9 * max(1.0, abs(v[2])) or abs(p * -3.5 + q - v[3]) > 1e-9 * max(1.0, abs(v[3])):
            raise NotImplementedError("wealth trigger action is not of the form scale * wealth + shift")
        return AttributeRule(cond, attr, scale=p, shift=q)
    probes = {'status': list(Status), 'infected_status': list(InfectionSeverity)}.get(attr, [0, 1, 7, 30])
    out = [action(x) for x in probes]
    if any(o != out[0] for o in out):
        raise NotImplementedError("trigger action on {} must set a constant value".format(attr))
    return AttributeRule(cond, attr, value=out[0])
