"""Generates tests/golden/graph_*.json from the UNMODIFIED reference GraphSimulation (TEST INFRASTRUCTURE ONLY).

Run in the authoring container, where /root/reference exists:

    python oracle/gen_graph_golden.py

For every scenario family of run_batch.py the reference is initialised under `np.random.seed`, converted to a
`World` dict (oracle.graph_oracle.world_from_reference) and stepped with the keyed draws of
oracle/graph_shim.py.  The fixture holds the initial world, the statistics row after every iteration
(graph_abs.py:302-324) and the final person state -- all produced by the reference's own code.
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

from covid_abs.network.graph_abs import GraphSimulation  # noqa: E402
from covid_abs.agents import Status  # noqa: E402
from covid_abs.network.agents import EconomicalStatus  # noqa: E402
import covid_abs.network.util as util  # noqa: E402
from oracle.graph_oracle import (world_from_reference, GRAPH_STAT_KEYS, MODE_NONE, MODE_LOCKDOWN,  # noqa: E402
                                 MODE_CONDITIONAL_LOCKDOWN, MODE_VERTICAL)
from oracle.graph_shim import KeyedRandom, reference_simulation  # noqa: E402

MODS = (GraphSimulation, Status, EconomicalStatus, util)
MODES = {"none": MODE_NONE, "lockdown": MODE_LOCKDOWN, "conditional_lockdown": MODE_CONDITIONAL_LOCKDOWN,
         "vertical": MODE_VERTICAL, "isolation": MODE_NONE}
STC = {"s": 0, "i": 1, "c": 2, "m": 3}
SEVC = {"e": 0, "a": 1, "h": 2, "g": 3}

CASES = [  # name, scenario, np seed, draw seed, iterations, overrides
    ("graph_none", "none", 3, 77, 150, dict(population_size=90, initial_infected_perc=0.05)),
    ("graph_lockdown", "lockdown", 4, 78, 100, dict(population_size=90, initial_infected_perc=0.05)),
    ("graph_conditional", "conditional_lockdown", 5, 79, 150, dict(population_size=90, initial_infected_perc=0.04)),
    ("graph_vertical", "vertical", 6, 80, 100, dict(population_size=90, initial_infected_perc=0.05)),
    ("graph_isolation", "isolation", 7, 81, 100, dict(population_size=90, initial_infected_perc=0.05)),
    ("graph_masks", "none", 8, 82, 100, dict(population_size=90, initial_infected_perc=0.1, contagion_distance=.05,
                                             contagion_rate=.1)),
    ("graph_month", "none", 9, 83, 760, dict(population_size=60, initial_infected_perc=0.05)),
    # many homeless and unemployed persons (own-wealth flows, government payments at the monthly accounting)
    ("graph_homeless", "none", 10, 84, 760, dict(population_size=60, initial_infected_perc=0.05, homeless_rate=0.6,
                                                 unemployment_rate=0.5)),
    # few businesses (the other branch of graph_abs.py:138-143), per-status amplitudes, short infectious window
    ("graph_fewbusiness", "conditional_lockdown", 11, 85, 200,
     dict(population_size=80, initial_infected_perc=0.08, total_business=4, incubation_time=2, contagion_time=6,
          recovering_time=9, business_distance=60)),
    # BASELINE.json configs[1] at its own shape: 1 000 persons (333 houses, 9 businesses), scenario2 (lockdown), past the
    # first monthly accounting (iteration 720).  ~20 minutes of the reference's O(N^2) loop: generated on request only
    # (`python oracle/gen_graph_golden.py graph_config2`)
    ("graph_config2", "lockdown", 21, 90, 760, dict(population_size=1000)),
]
SLOW = {"graph_config2"}


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    return o


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    only = sys.argv[1:]
    for name, scenario, np_seed, seed, iters, over in CASES:
        if (only and name not in only) or (not only and name in SLOW):
            continue
        sim, iso = reference_simulation(MODS, scenario, np_seed, **over)
        world = world_from_reference(sim, mode=MODES[scenario], sleep=True, isolated_ids=iso)
        rows = []
        with KeyedRandom(sim, seed):
            for _ in range(iters):
                sim.execute()
                st = sim.get_statistics('all')
                rows.append([float(st[k]) for k in GRAPH_STAT_KEYS])
        pop = sim.population
        fix = dict(name=name, scenario=scenario, np_seed=np_seed, seed=seed, iterations=iters, overrides=over, world=world,
                   stat_keys=list(GRAPH_STAT_KEYS), rows=rows,
                   final=dict(px=[int(a.x) for a in pop], py=[int(a.y) for a in pop],
                              status=[STC[a.status.value] for a in pop],
                              severity=[SEVC[a.infected_status.value] for a in pop],
                              infected_time=[int(a.infected_time) for a in pop],
                              employer=[sim.business.index(a.employer) if a.employer is not None else -1 for a in pop],
                              hwealth=[float(np.asarray(h.wealth).ravel()[0]) for h in sim.houses],
                              bwealth=[float(np.asarray(b.wealth).ravel()[0]) for b in sim.business],
                              bstocks=[int(b.stocks) for b in sim.business]))
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(jsonable(fix), f)
        print(name, "ok", rows[-1][7:11])


if __name__ == "__main__":
    main()
