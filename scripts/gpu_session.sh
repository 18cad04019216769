set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/ab_variants.sh 2>&1 | tee gpurun_out/ab_r2_3.log
python scripts/dense_epidemic.py 2>&1 | tail -5
