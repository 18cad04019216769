"""Off-lattice golden trajectories from the UNMODIFIED reference under keyed draws (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_offlattice_golden.py      # writes tests/golden/simulation_offlattice.json

The reference `Simulation` with `triggers_population` entries written the way the reference takes them -- Python
lambdas (abs.py:86-95; the move idiom of run_batch.py:446-450: a point plus np.random.normal(0, sigma, 1) per axis) --
stepped while np.random.* is replaced by the counter-based draws of oracle/philox.py (oracle/sim_shim.py, extended
here with np.random.normal for the move actions).  Positions become binary64 as soon as a move action fires, the
contact test is the reference's own floating-point one, and contacts are applied in its sequential order, so the
fixture pins: rule matching, the float walk and reflection, the binary64 distance predicate, in-step chains.
Each case also carries a declarative description of its triggers, from which the tests rebuild the device rules.
"""
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = os.environ.get("COVID_ABS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from covid_abs.abs import Simulation  # noqa: E402
from covid_abs.agents import Status, InfectionSeverity  # noqa: E402
from oracle import philox  # noqa: E402
from oracle.sim_shim import KeyedSimulationRandom  # noqa: E402

STC = {Status.Susceptible: 0, Status.Infected: 1, Status.Recovered_Immune: 2, Status.Death: 3}
SEVC = {InfectionSeverity.Exposed: 0, InfectionSeverity.Asymptomatic: 1, InfectionSeverity.Hospitalization: 2,
        InfectionSeverity.Severe: 3}


def normal_pair_for(shim, agent_index):
    w = philox.agent_draws(shim.seed, agent_index, shim.step, philox.STREAM_MOVE)
    return philox.normal_pair(w[0], w[1])


class KeyedWithNormal(KeyedSimulationRandom):
    """+ np.random.normal(0, sigma, 1) inside a move action: sigma x the move stream's normals (first call x, then y)."""

    def __enter__(self):
        super(KeyedWithNormal, self).__enter__()
        self._saved_normal = np.random.normal
        np.random.normal = self.normal
        return self

    def __exit__(self, *exc):
        np.random.normal = self._saved_normal
        return super(KeyedWithNormal, self).__exit__(*exc)

    def rand(self, n):
        # a scalar instead of a shape-(1,) array: agents moved by a rule never add the income, so the reference's
        # wealth list would mix floats and arrays, which np.sum (abs.py:310) refuses under numpy >= 2; same arithmetic
        f = sys._getframe(1)
        assert f.f_code.co_name == "move"
        w = philox.agent_draws(self.seed, self.index[id(f.f_locals["agent"])], self.step, philox.STREAM_ECONOMY)
        return np.float64(philox.u01(w[0]))

    def normal(self, loc=0.0, scale=1.0, size=None):
        f = sys._getframe(1)
        while f is not None and f.f_code.co_name != "move":
            f = f.f_back
        assert f is not None, "np.random.normal outside a move action"
        a = self.index[id(f.f_locals["agent"])]
        z = normal_pair_for(self, a)
        k = self.calls.get(("m", a), 0)
        self.calls[("m", a)] = k + 1
        return np.array([np.float64(loc) + np.float64(z[k]) * np.float64(scale)])


def cases():
    inf_asym = lambda a: a.status == Status.Infected and a.infected_status == InfectionSeverity.Asymptomatic
    return [
        dict(name="quarantine_in_place",
             kwargs=dict(population_size=70, length=30, height=30, contagion_distance=1.6, contagion_rate=0.8,
                         initial_infected_perc=0.1, initial_immune_perc=0.02),
             np_seed=11, seed=501, iterations=30,
             triggers=[{'condition': inf_asym, 'attribute': 'move',
                        'action': lambda a: (a.x + np.random.normal(0, 0.3, 1), a.y + np.random.normal(0, 0.3, 1))}],
             spec=[dict(kind="move", when=dict(status=["Infected"], infected_status=["Asymptomatic"]), target="stay",
                        sigma=0.3)]),
        dict(name="gather_and_tax",
             kwargs=dict(population_size=60, length=25, height=25, contagion_distance=0.75, contagion_rate=0.9,
                         initial_infected_perc=0.08, initial_immune_perc=0.0),
             np_seed=12, seed=502, iterations=25,
             triggers=[{'condition': lambda a: a.status == Status.Susceptible and a.age >= 40, 'attribute': 'move',
                        'action': lambda a: (12.5 + np.random.normal(0, 2.0, 1), 7.25 + np.random.normal(0, 2.0, 1))},
                       {'condition': lambda a: a.age > 60, 'attribute': 'wealth', 'action': lambda w: w * 0.99},
                       {'condition': lambda a: a.status == Status.Recovered_Immune, 'attribute': 'infected_time',
                        'action': lambda t: 7}],
             spec=[dict(kind="move", when=dict(status=["Susceptible"], min_age=40), target=[12.5, 7.25], sigma=2.0),
                   dict(kind="attr", when=dict(min_age=61), attribute="wealth", scale=0.99, shift=0.0),
                   dict(kind="attr", when=dict(status=["Recovered_Immune"]), attribute="infected_time", value=7)]),
    ]


def main():
    out = []
    for c in cases():
        np.random.seed(c["np_seed"])
        sim = Simulation(**c["kwargs"])
        sim.initialize()
        for t in c["triggers"]:
            sim.append_trigger_population(t['condition'], t['attribute'], t['action'])
        pop = sim.get_population()
        initial = dict(x=[int(a.x) for a in pop], y=[int(a.y) for a in pop], status=[STC[a.status] for a in pop],
                       age=[int(a.age) for a in pop], stratum=[int(a.social_stratum) for a in pop],
                       wealth=[float(np.asarray(a.wealth).ravel()[0]) for a in pop])
        sim.get_statistics()
        rows, keys = [], None
        with KeyedWithNormal(sim, c["seed"]) as kr:
            for _ in range(c["iterations"]):
                sim.execute()
                st = sim.get_statistics('all')
                keys = list(st.keys())
                rows.append([float(st[k]) for k in keys])
                kr.next_step()
        num = lambda v: float(np.asarray(v).ravel()[0])
        out.append(dict(name=c["name"], kwargs=c["kwargs"], seed=c["seed"], iterations=c["iterations"], spec=c["spec"],
                        total_wealth=float(sim.total_wealth), initial=initial, keys=keys, rows=rows,
                        final=dict(x=[num(a.x) for a in pop], y=[num(a.y) for a in pop],
                                   status=[STC[a.status] for a in pop], severity=[SEVC[a.infected_status] for a in pop],
                                   infected_time=[int(a.infected_time) for a in pop],
                                   wealth=[num(a.wealth) for a in pop])))
        print(c["name"], "final Infected", rows[-1][1], "non-lattice positions",
              int(sum(num(a.x) != np.floor(num(a.x)) for a in pop)))
    with open(os.path.join(ROOT, "tests", "golden", "simulation_offlattice.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
