timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/gputests_r2_final.log
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err || tail -5 gpurun_out/bench_r2_final.err
timeout 600 python bench.py --workload ensemble > gpurun_out/ens_r2_final.json 2> gpurun_out/ens_r2_final.err || tail -5 gpurun_out/ens_r2_final.err
ncu --set full --clock-control none --import-source on -k regex:k_graph_run --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_graph_r2_final python scripts/graph_phase_times.py --scenario 0 --seeds 296 --days 2 > gpurun_out/ncu_graph_r2_final.log 2>&1
tail -1 gpurun_out/ncu_graph_r2_final.log
python scripts/graph_phase_times.py --scenario 0 --seeds 296 --days 2,10,25 2>&1 | grep "^day"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_final.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
print({k: round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()})
oc=d.get("other_configs", {}); print({k: (oc[k].get("ms"), oc[k].get("ms_per_step"), oc[k].get("agent_updates_per_s")) for k in oc})
d=json.loads(open("gpurun_out/ens_r2_final.json").read().strip().splitlines()[-1])
print("ensemble value %.4g ms/iter %.3f frac80 %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]), d["e2e"]["value"])
PY
