#!/bin/bash
# effect of the cell edge (via the max_cells cap) on the per-kernel times of the bench workload
for mc in ${SWEEP:-0 4000000 1000000}; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --e2e-steps 3 --max-cells $mc > gpurun_out/sweep_$mc.json 2> gpurun_out/sweep_$mc.err || tail -3 gpurun_out/sweep_$mc.err
  python - $mc <<PY
import json, sys
d=json.load(open("gpurun_out/sweep_%s.json" % sys.argv[1]))
print("max_cells %s edge %d: step %.4f ms" % (sys.argv[1], d["config"]["cell_edge"], d["ms_per_step"]))
for k,v in d["roofline"]["kernels"].items(): print("  %-26s %.4f ms" % (k, v["avg_ms"]))
PY
done
