run8(){ python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 8 "$@"; }
run8 --workload ensemble > gpurun_out/ens_8gpu_r2_final.json 2> gpurun_out/ens_8gpu_r2_final.err || tail -5 gpurun_out/ens_8gpu_r2_final.err
python - <<'PY'
import json
for f in ("ens_8gpu_r2_final",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.4f frac %s" % (d["value"], d["ms_per_step"], d["roofline"].get("frac")), d["roofline"].get("whole_step"), d["e2e"], d["host"])
    except Exception as e: print(f, "failed", e)
PY
