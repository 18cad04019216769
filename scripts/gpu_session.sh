timeout 1500 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/gputests_graph.log
CABS_GRAPH_QUEUE=1 timeout 1500 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -2
for s in 0 2; do
scripts/with_variant.sh build/variants_graph/graph_timers.so python scripts/graph_phase_times.py --scenario $s --seeds 296 --days 2,10,14 2>&1 | grep -v Warning
done | tee gpurun_out/graph_phase_times_by_scenario12.log
timeout 600 python bench.py --workload ensemble > gpurun_out/ens_r2_n.json 2> gpurun_out/ens_r2_n.err || tail -5 gpurun_out/ens_r2_n.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ens_r2_n.json").read().strip().splitlines()[-1])
print("ensemble value %.4g ms/iter %.3f frac80 %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]), d["final_infected_avg_std_by_scenario"][:2])
PY
