"""Batch driver -- mirror of covid_abs/experiments.py:167-216 (plotting helpers are out of scope).

Same signature, same CSV schema (`Iteration,Metric,Min,Avg,Std,Max`, experiments.py:212), same
"an experiment may abort, the batch continues" contract (experiments.py:182-198).
"""
import numpy as np
import pandas as pd

from covid19_agentbasedsimulation_b200.abs import Simulation


def batch_experiment(experiments, iterations, file, simulation_type=Simulation, **kwargs):
    """`experiments` independent runs of `simulation_type(**kwargs)`, `iterations` steps each; per iteration and
    metric the minimum, mean, population standard deviation and maximum over the runs are written to `file` as CSV and
    returned as a DataFrame (experiments.py:167-216).  `verbose='experiments'|'iterations'` prints progress."""
    verbose = kwargs.get('verbose', None)
    columns = None
    runs = []  # one [iterations_done, n_metrics] array per experiment
    aggregated = getattr(simulation_type, 'run_ensemble_aggregated', None)
    if aggregated is not None and verbose != 'iterations' and experiments > 0:
        # all experiments step concurrently on the device and the table below is reduced there too
        try:
            columns, agg = aggregated(experiments, iterations, **kwargs)
            rows2 = [[it, col] + [float(v) for v in agg[it, k]] for it in range(iterations)
                     for k, col in enumerate(columns)]
            df2 = pd.DataFrame(rows2, columns=['Iteration', 'Metric', 'Min', 'Avg', 'Std', 'Max'])
            df2.to_csv(file, index=False)
            return df2
        except Exception as ex:
            print("Exception occurred in the aggregated ensemble run: {}".format(ex))
    ensemble = getattr(simulation_type, 'run_ensemble', None)
    if ensemble is not None and verbose != 'iterations' and experiments > 0:
        # all experiments step concurrently on the device (one block per simulation), one copy back
        try:
            columns, table = ensemble(experiments, iterations, **kwargs)
            runs = [table[k] for k in range(table.shape[0])]
            experiments = 0
        except Exception as ex:
            print("Exception occurred in the ensemble run: {}".format(ex))
            columns, runs = None, []
    for experiment in range(experiments):
        rows = []
        try:
            if verbose == 'experiments':
                print('Experiment {}'.format(experiment))
            sim = simulation_type(**kwargs)
            sim.initialize()
            if columns is None:
                statistics = sim.get_statistics(kind='all')
                columns = [k for k in statistics.keys()]
            if hasattr(sim, 'run') and verbose != 'iterations':
                # the whole run is enqueued on the GPU; rows come back in one copy
                table = sim.run(iterations)
                keys = list(sim.get_statistics(kind='all').keys())
                rows = [table[:, [keys.index(c) for c in columns]]]
            else:
                for it in range(iterations):
                    if verbose == 'iterations':
                        print('Experiment {}\tIteration {}'.format(experiment, it))
                    sim.execute()
                    statistics = sim.get_statistics(kind='all')
                    rows.append(np.array([[statistics[c] for c in columns]], dtype=np.float64))
        except Exception as ex:
            print("Exception occurred in experiment {}: {}".format(experiment, ex))
        if rows:
            runs.append(np.concatenate(rows, axis=0))

    rows2 = []
    for it in range(iterations):
        try:
            vals = np.array([r[it] for r in runs if len(r) > it], dtype=np.float64)
            for k, col in enumerate(columns):
                v = vals[:, k]
                rows2.append([it, col, v.min(), v.mean(), v.std(), v.max()])  # std: ddof=0, experiments.py:206
        except Exception as ex:
            print(ex)
    df2 = pd.DataFrame(rows2, columns=['Iteration', 'Metric', 'Min', 'Avg', 'Std', 'Max'])
    df2.to_csv(file, index=False)
    return df2
