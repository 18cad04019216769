python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/gputests_r2_final.log
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err || tail -5 gpurun_out/bench_r2_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err || tail -5 gpurun_out/bench_r2_final_reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_final.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f e2e %.4g launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
d=json.loads(open("gpurun_out/bench_r2_final_reference.json").read().strip().splitlines()[-1])
print("reference arm:", {k: d[k] for k in ("impl","value","unit","n_gpus","steps") if k in d}, d.get("cpu_baseline",{}).get("kind"), d.get("cpu_baseline",{}).get("cores"))
PY
