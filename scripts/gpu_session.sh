python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/gputests_r2_final.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: value %.4g ms/step %.4f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['frac']))"
