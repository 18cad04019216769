#!/bin/bash
# ncu evidence for one round: launch list of a short bench run + one full capture of the step kernels.
# usage: scripts/profile_step.sh <tag> [extra bench args]   (run under gpurun; outputs land in gpurun_out/)
tag=${1:-r1}; shift
args="--steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --e2e-steps 3 $*"
set -x
python bench.py $args > gpurun_out/plain_$tag.log 2>&1 || { tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $args > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_update_scatter|k_contagion|k_move_count|k_scan_cells" -s 12 -c 4 -o gpurun_out/prof_$tag python bench.py $args > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -8
