set -x
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_update_scatter|k_contagion|k_move_count" -s 9 -c 3 -o gpurun_out/prof_r1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
