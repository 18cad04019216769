timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; tail -3 gpurun_out/bench_r2_c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2_c.json"))
print("value %.4g ms/step %.4f e2e %.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print("e2e_python", d["e2e_python"]["ms_per_step"], d["e2e_python"]["fraction_of_value"], "e2e_device", d["e2e_device_inputs"])
for k,v in d["roofline"]["kernels"].items(): print("  %-26s %.4f ms" % (k, v["avg_ms"]))
print({k:(v.get("ms_per_step"), v.get("agent_updates_per_s")) for k,v in d["other_configs"].items() if isinstance(v, dict)})
PY
