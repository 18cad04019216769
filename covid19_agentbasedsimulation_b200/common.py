"""Disease and wealth tables used by the step -- same names and values as covid_abs/common.py:13-27.

The device keeps copies in the parameter block of every handle (`cabs_params`, include/covidabs.h); the age tables
are indexed by the reference's shifted age bin (`age // 10 - 1 if age > 10 else 0`, abs.py:204): entry 0 covers ages
0-19, entry k covers 10(k+1) .. 10(k+1)+9, entry 8 covers 90-99.
"""
import numpy as np

# (hospitalisation, severe, death) probability per age bin, one row per bin
_BY_AGE_BIN = (
    (0.001, 0.05, 0.002),
    (0.003, 0.05, 0.00006),
    (0.012, 0.05, 0.0003),
    (0.032, 0.05, 0.0008),
    (0.049, 0.063, 0.0015),
    (0.102, 0.122, 0.006),
    (0.166, 0.274, 0.022),
    (0.243, 0.432, 0.051),
    (0.273, 0.709, 0.093),
)
age_hospitalization_probs, age_severe_probs, age_death_probs = (list(col) for col in zip(*_BY_AGE_BIN))

# share of the total wealth held by each quintile (a Lorenz curve), poorest first
lorenz_curve = [.04, .08, .13, .2, .55]
share = min(lorenz_curve)
# daily income / expense unit of each quintile relative to the poorest one: [1, 2, 3.25, 5, 13.75]
basic_income = np.asarray(lorenz_curve, dtype=np.float64) / share
