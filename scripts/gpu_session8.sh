run8(){ python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 8 "$@"; }
run8 --no-cpu-baseline > gpurun_out/bench_8gpu_r2_final.json 2> gpurun_out/bench_8gpu_r2_final.err || tail -5 gpurun_out/bench_8gpu_r2_final.err
python - <<'PY'
import json
for f in ("bench_8gpu_r2_final",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.4f frac %s inv %s" % (d["value"], d["ms_per_step"], d["roofline"].get("frac"), d.get("slab_invariance")), d["roofline"].get("whole_step"))
        print("   ", {k: (v.get("ms_per_step"), v.get("agent_updates_per_s"), v.get("whole_step_frac")) for k,v in d.get("other_configs",{}).items() if isinstance(v, dict)})
        print("   ", {k: round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "failed", e)
PY
