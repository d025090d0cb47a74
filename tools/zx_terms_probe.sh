#!/bin/bash
# A/B of the split-precision term count of the hoisted LSTM2 projection: accuracy against the oracle and pass time
for t in 5 7; do
  echo "== C3R_ZX_TERMS=$t"
  C3R_ZX_TERMS=$t python tools/prec_probe.py 2>&1 | grep "max|dp|" | awk '{print "   ", $1, $2, $3, $6, $7, $8, $9}'
  C3R_ZX_TERMS=$t tools/bench_brief.sh --steps 10 2>&1 | head -2
done
