#!/bin/bash
# A/B timing of prebuilt library variants (build/variants/*.so): per-kernel device times of the bench workload
export CABS_AB_OLD_LIB=1
cp covid19_agentbasedsimulation_b200/libcovidabs.so /tmp/keep.so
for f in build/variants/*.so; do
  cp "$f" covid19_agentbasedsimulation_b200/libcovidabs.so
  touch covid19_agentbasedsimulation_b200/libcovidabs.so
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-other-configs --e2e-steps 3 "$@" > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - "$f" <<PY
import json, sys
d=json.load(open("gpurun_out/ab.json"))
k=d["roofline"]["kernels"]
print("%-28s step %.4f ms | A %.4f scan %.4f B %.4f C %.4f" % (sys.argv[1].split('/')[-1], d["ms_per_step"], k["k_move_count<true>"]["avg_ms"], k["k_scan_cells"]["avg_ms"], k["k_update_scatter<true>"]["avg_ms"], k["k_contagion"]["avg_ms"]))
PY
done
cp /tmp/keep.so covid19_agentbasedsimulation_b200/libcovidabs.so
touch covid19_agentbasedsimulation_b200/libcovidabs.so
