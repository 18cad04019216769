timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/gputests_graph.log
CABS_GRAPH_QUEUE=1 timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -2
for s in 0 2; do
scripts/with_variant.sh build/variants_graph/graph_timers.so python scripts/graph_phase_times.py --scenario $s --seeds 296 --days 10 2>&1 | grep -v Warning
done | tee gpurun_out/graph_phase_times_by_scenario9.log
