python scripts/diag_config3_stats.py 2>&1 | grep -v Warning | cut -c1-260 | tail -18
