set -x
bash scripts/ab_variants.sh 2>&1 | tee gpurun_out/ab_r2_5.log
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py --workload ensemble > gpurun_out/bench_ens_r2_a.json 2> gpurun_out/bench_ens_r2_a.err; tail -5 gpurun_out/bench_ens_r2_a.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_ens_r2_a.json"))
print(json.dumps({k:v for k,v in d.items() if k not in ("config",)}, indent=None)[:3000])
PY
