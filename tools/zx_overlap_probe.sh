#!/bin/bash
# projection tiles run under LSTM1's remainder round: tile pairs per idle SM pair (0 = no overlap)
for t in 0 5 7 10; do
  echo "== C3R_ZX_OVERLAP=$t"
  C3R_ZX_OVERLAP=$t tools/bench_brief.sh --steps 10 2>&1 | head -2
done
