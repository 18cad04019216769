for v in m_131072 m_65536; do echo "== $v"
R="scripts/with_variant.sh build/variants_graph/$v.so"
for s in 2 0 1; do $R python scripts/graph_phase_times.py --scenario $s --seeds 296 --days 2,10 2>&1 | grep "^day"; done
$R python bench.py --workload ensemble 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ensemble value %.4g frac80 %.3f'%(d['value'], d['roofline']['frac']), d['final_infected_avg_std_by_scenario'][:2])"
done | tee gpurun_out/graph_ab_mapbits.log
