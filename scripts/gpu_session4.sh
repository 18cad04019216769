python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/bench_4gpu_r2_final.json 2> gpurun_out/bench_4gpu_r2_final.err || tail -5 gpurun_out/bench_4gpu_r2_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_4gpu_r2_final.json").read().strip().splitlines()[-1])
print("4 GPUs: value %.4g ms/step %.4f inv %s" % (d["value"], d["ms_per_step"], d.get("slab_invariance")), d["roofline"].get("whole_step"), {k: (v.get("ms_per_step"), v.get("agent_updates_per_s")) for k,v in d.get("other_configs",{}).items() if isinstance(v, dict)})
PY
