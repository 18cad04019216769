set -x
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_${N}gpu_r2_b.json 2> gpurun_out/bench_${N}gpu_r2_b.err; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_${N}gpu_r2_b.err | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload ensemble > gpurun_out/bench_ens_${N}gpu_r2_b.json 2> gpurun_out/bench_ens_${N}gpu_r2_b.err; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_ens_${N}gpu_r2_b.err | tail -5
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("gpurun_out/bench_%sgpu_r2_b.json"%N, "gpurun_out/bench_ens_%sgpu_r2_b.json"%N):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.4g ms/step %.4f e2e %.4g clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
    print({k:d.get(k) for k in ("slab_invariance","mismatching_statistics_rows","mismatching_agents","agents")})
    print("roofline", {k:v for k,v in d["roofline"].items() if k!="kernels"})
    print("other", d.get("other_configs"), d.get("host"), d.get("final_infected_avg_std_by_scenario"))
PY
