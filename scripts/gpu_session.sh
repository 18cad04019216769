timeout 900 python -m pytest tests/test_gpu_slabs.py -q -x -k "two_process" 2>&1 | tail -3
run2(){ python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline --no-other-configs "$@"; }
run2 > gpurun_out/bench_2gpu_edgefirst.json 2> gpurun_out/bench_2gpu_edgefirst.err || tail -5 gpurun_out/bench_2gpu_edgefirst.err
CABS_SLAB_EDGE_FIRST=0 run2 > gpurun_out/bench_2gpu_noedge.json 2> gpurun_out/bench_2gpu_noedge.err || tail -5 gpurun_out/bench_2gpu_noedge.err
python - <<'PY'
import json
for f in ("edgefirst","noedge"):
    try:
        d=json.loads(open("gpurun_out/bench_2gpu_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f value %.4g invariance %s" % (d["ms_per_step"], d["value"], d.get("slab_invariance")), d.get("slab_invariance_detail"), d.get("final_stats"))
    except Exception as e: print(f, "failed", e)
PY
