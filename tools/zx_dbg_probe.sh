#!/bin/bash
# where does k_gemm_zx's time go: 0 = normal, 1 = no output stores, 2 = no weight-ring traffic (results are garbage for 1, 2)
for m in 0 1 2; do
  echo "== C3R_ZX_DBG=$m"
  C3R_ZX_DBG=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"k_gemm_zx" python bench.py --steps 1 --warmup 1 --no_cpu_baseline 2>/dev/null | grep -E "k_gemm_zx" | tail -2 | awk -F'","' '{print "   ", $NF}'
done
