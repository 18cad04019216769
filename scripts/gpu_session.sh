set -x
timeout 900 python -m pytest tests/test_offlattice.py -m gpu -x -q 2>&1 | tail -30
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_offlattice.py 2>&1 | tail -4
CABS_PDL=1 CABS_DEBUG=1 python bench.py --steps 100 --no-cpu-baseline --no-other-configs --e2e-steps 3 2>&1 >gpurun_out/pdl_dbg.json | grep -i "capture\|error" | head -5
python - <<'PY'
import json
d=json.load(open("gpurun_out/pdl_dbg.json"))
print("CABS_PDL=1 step %.4f ms" % d["ms_per_step"])
PY
