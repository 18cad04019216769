"""B200-native COVID-ABS agent step behind the covid_abs API.

    from covid19_agentbasedsimulation_b200.abs import Simulation
    from covid19_agentbasedsimulation_b200.experiments import batch_experiment

mirror `covid_abs.abs.Simulation` / `covid_abs.experiments.batch_experiment`; `compat.install()` makes
`import covid_abs` resolve to this package for unmodified user scripts.
"""
__version__ = "0.1.0"
