set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pdl in 0 1; do
CABS_NO_PDL=$pdl python bench.py --steps 100 --no-cpu-baseline --no-other-configs --e2e-steps 3 > gpurun_out/pdl_$pdl.json 2>/dev/null
python - $pdl <<'PY'
import json,sys
d=json.load(open("gpurun_out/pdl_%s.json"%sys.argv[1]))
print("CABS_NO_PDL=%s  step %.4f ms  value %.4g  e2e_python %.4f" % (sys.argv[1], d["ms_per_step"], d["value"], d["e2e_python"]["ms_per_step"]))
PY
done
