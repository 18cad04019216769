"""Times the UNMODIFIED reference (petroniocandido/COVID19_AgentBasedSimulation, `covid_abs`) on the host cores.

The reference is pure Python; `baseline/_ref` holds its own package, installed offline with
    python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>
(git-ignored: it is the reference's code, not ours; it travels to the GPU box with the snapshot).  This module only
imports and drives it through its public API -- `Simulation(**kwargs).initialize(); execute(); get_statistics('all')`
(covid_abs/experiments.py:185-194) -- nothing of this repository is on that path.

SURVEY.md section 8d, "CPU side-by-side (i)": the loop `for _ in range(T): sim.execute(); sim.get_statistics('all')` at
N in {20, 80, 300, 1000, 3000}, single core (the reference is single-threaded), with the fitted O(N^2) law and its
extrapolation to 10^7 agents (a direct run is impossible: abs.py:249 allocates an N x N matrix, 8e14 bytes at 10^7).
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), os.environ.get("COVID_ABS_REFERENCE", "/root/reference"))


def import_reference():
    """(module covid_abs.abs, module covid_abs.network.graph_abs, path) of the unmodified reference, or None."""
    for path in CANDIDATES:
        if path and os.path.isdir(os.path.join(path, "covid_abs")):
            saved = {k: v for k, v in sys.modules.items() if k == "covid_abs" or k.startswith("covid_abs.")}
            for k in saved:
                del sys.modules[k]
            sys.path.insert(0, path)
            try:
                warnings.filterwarnings("ignore")
                import covid_abs.abs as ref_abs
                import covid_abs.network.graph_abs as ref_graph
                import covid_abs.agents as ref_agents
                return ref_abs, ref_graph, ref_agents, path
            except Exception:
                continue
            finally:
                sys.path.remove(path)
                for k in [k for k in sys.modules if k == "covid_abs" or k.startswith("covid_abs.")]:
                    sys.modules.pop(k)     # the drop-in alias and the reference never share sys.modules
                sys.modules.update(saved)
    return None


def time_simulation(ref_abs, n, iterations, seed=0):
    """Seconds per iteration of the stock loop at density 0.2 agents / unit^2 (the reference default 20 / 100), config-3
    style parameters (masks: distance .05, rate .1; 1 % infected, 1 % immune, amplitudes 5)."""
    side = max(2, int(round((n / 0.2) ** 0.5)))
    np.random.seed(seed)
    sim = ref_abs.Simulation(population_size=n, length=side, height=side, initial_infected_perc=0.01,
                             initial_immune_perc=0.01, contagion_distance=0.05, contagion_rate=0.1,
                             total_wealth=1e4 * n / 20)
    sim.initialize()
    sim.execute()
    sim.get_statistics('all')
    t0 = time.perf_counter()
    for _ in range(iterations):
        sim.execute()
        sim.get_statistics('all')
    return (time.perf_counter() - t0) / iterations


def time_graph(ref_graph, ref_agents, n, iterations, seed=0):
    """Seconds per EXECUTED iteration of GraphSimulation, lockdown scenario (run_batch.py:103-112) without the sleep
    callback, so that every iteration runs the full hour (the quantity SURVEY.md section 6 quotes: 0.58 s at 1 000)."""
    Status = ref_agents.Status

    def lockdown(a):
        if a.house is not None:
            a.house.checkin(a)
        return True
    np.random.seed(seed)
    sim = ref_graph.GraphSimulation(
        length=500, height=500, population_size=n, homemates_avg=3, homeless_rate=0.0005,
        amplitudes={Status.Susceptible: 10, Status.Recovered_Immune: 10, Status.Infected: 10},
        critical_limit=0.01, contagion_rate=.9, incubation_time=5, contagion_time=10, recovering_time=20,
        total_wealth=10000000, total_business=9, minimum_income=900.0, minimum_expense=600.0,
        public_gdp_share=0.1, business_gdp_share=0.5, unemployment_rate=0.12, business_distance=20,
        initial_infected_perc=.01, initial_immune_perc=.01, contagion_distance=1.,
        callbacks={'on_person_move': lockdown})
    sim.initialize()
    sim.execute()
    t0 = time.perf_counter()
    for _ in range(iterations):
        sim.execute()
        sim.get_statistics('all')
    return (time.perf_counter() - t0) / iterations


def reference_path(budget="default"):
    """The dictionary bench.py reports as `cpu_baseline.reference_path`."""
    ref = import_reference()
    if ref is None:
        return {"available": False, "why": "no unmodified reference under baseline/_ref or /root/reference"}
    ref_abs, ref_graph, ref_agents, path = ref
    plan = [(20, 200), (80, 60), (300, 8), (1000, 2), (3000, 1)] if budget == "default" else [(20, 50), (80, 10), (300, 2)]
    t_start = time.perf_counter()
    sizes, secs = [], []
    for n, iters in plan:
        secs.append(time_simulation(ref_abs, n, iters))
        sizes.append(n)
    # t(N) = a + b N + c N^2: the pair loop of abs.py:253-259 is the N^2 term
    A = np.array([[1.0, n, float(n) ** 2] for n in sizes])
    coef, *_ = np.linalg.lstsq(A / np.array(secs)[:, None], np.ones(len(secs)), rcond=None)   # relative-error fit
    a, b, c = [float(v) for v in coef]
    n_big = 1e7
    t_big = a + b * n_big + c * n_big ** 2
    out = {
        "available": True, "source": path, "cores": 1,
        "loop": "sim.execute(); sim.get_statistics('all')  (covid_abs/abs.py:232-314), stock code path",
        "simulation": [{"agents": n, "s_per_iteration": s, "agent_updates_per_s": n / s} for n, s in zip(sizes, secs)],
        "fit_seconds_per_iteration": {"a": a, "b_per_agent": b, "c_per_agent2": c},
        "extrapolated_1e7": {"s_per_iteration": t_big, "agent_updates_per_s": n_big / t_big,
                             "note": "O(N^2) law fitted above; a direct run needs an 8e14-byte matrix (abs.py:249)"},
    }
    graph = []
    for n, iters in ([(300, 10), (1000, 3)] if budget == "default" else [(300, 3)]):
        s = time_graph(ref_graph, ref_agents, n, iters)
        graph.append({"persons": n, "s_per_executed_iteration": s, "agent_updates_per_s": n / s})
    out["graph_lockdown"] = graph
    out["wall_s"] = time.perf_counter() - t_start
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(reference_path(sys.argv[1] if len(sys.argv) > 1 else "default"), indent=1))
