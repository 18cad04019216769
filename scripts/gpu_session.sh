bash scripts/ab_variants.sh 2>&1 | tee gpurun_out/ab_r2_7_diag.log
