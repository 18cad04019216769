"""Constants of the disease and of the wealth distribution -- mirror of covid_abs/common.py:13-27."""
import numpy as np

# Probability of each infection severity by age bin (common.py:13-15)
age_hospitalization_probs = [0.001, 0.003, 0.012, 0.032, 0.049, 0.102, 0.166, 0.243, 0.273]
age_severe_probs = [0.05, 0.05, 0.05, 0.05, 0.063, 0.122, 0.274, 0.432, 0.709]
age_death_probs = [0.002, 0.00006, 0.0003, 0.0008, 0.0015, 0.006, 0.022, 0.051, 0.093]

# Wealth distribution by quintile -- Lorenz curve (common.py:25-27)
lorenz_curve = [.04, .08, .13, .2, .55]
share = np.min(lorenz_curve)
basic_income = np.array(lorenz_curve) / share
