for m in 0 1; do
  C3R_ZX_DBG=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"k_gemm_zx|k_lstm_tc" python bench.py --steps 1 --warmup 1 --no_cpu_baseline 2>/dev/null | grep -E "k_gemm_zx|k_lstm" | tail -3 | awk -F'","' -v m=$m '{print "dbg="m, $5, $NF}'
done
