timeout 900 python -m pytest tests/test_gpu_slabs.py -q -x 2>&1 | tail -5
run2(){ python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline "$@"; }
run2 > gpurun_out/bench_2gpu_fused.json 2> gpurun_out/bench_2gpu_fused.err || tail -5 gpurun_out/bench_2gpu_fused.err
CABS_SLAB_UNFUSED=1 run2 > gpurun_out/bench_2gpu_unfused.json 2> gpurun_out/bench_2gpu_unfused.err || tail -5 gpurun_out/bench_2gpu_unfused.err
python - <<'PY'
import json
for f in ("fused","unfused"):
    try:
        d=json.loads(open("gpurun_out/bench_2gpu_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f value %.4g invariance %s" % (d["ms_per_step"], d["value"], d.get("slab_invariance")))
        print("   ", {k: round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "failed", e)
PY
