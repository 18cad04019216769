timeout 1500 python -m pytest tests/test_gpu_graph.py tests/test_gpu_statistical.py -m gpu -q -x 2>&1 | tail -2
CABS_GRAPH_QUEUE=0 timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -2
python bench.py --workload ensemble 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default: ensemble value %.4g frac80 %.3f'%(d['value'], d['roofline']['frac']))"
scripts/with_variant.sh build/variants_graph/m_524288.so python bench.py --workload ensemble 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('maps 2x64KB: ensemble value %.4g frac80 %.3f'%(d['value'], d['roofline']['frac']))"
python scripts/graph_bench.py 2>&1 | tail -8
