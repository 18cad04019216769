#!/bin/bash
# tests + short bench with per-kernel table (used during kernel iteration)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print("value %.4g  ms/step %.4f  e2e %.4g  step-frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["whole_step"]["frac"]))
for k,v in d["roofline"]["kernels"].items(): print("  %-26s %.4f ms  %6.0f GB/s" % (k, v["avg_ms"], v.get("achieved_gbs",0)))
PY
tail -3 gpurun_out/bench_quick.err
