#!/bin/bash
# A/B of the split-precision products of the L4 GEMM (mask: 1 = A_hi*B_hi, 2 = A_lo*B_hi, 4 = A_hi*B_lo)
for t in 7 5 3; do
  echo "== C3R_L4_TERMS=$t"
  C3R_L4_TERMS=$t python tools/prec_probe.py 2>&1 | grep "max|dp|" | grep "sharpen 8" | cut -c1-90
  C3R_L4_TERMS=$t tools/bench_brief.sh --steps 10 2>&1 | head -2
done
