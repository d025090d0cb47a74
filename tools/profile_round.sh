#!/bin/bash
# ncu evidence of one bench pass (config 2): launch list + one `--set full` capture of the main kernels.
# usage: tools/profile_round.sh TAG      -> gpurun_out/launches_TAG.csv, gpurun_out/prof_TAG.ncu-rep, gpurun_out/b_TAG*.log
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/b_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_lstm_tc|k_gemm_zx|k_gemm_tc|k_heads|k_rows|k_cmp|k_scatter|k_window|k_altinfo|k_xop" -s 12 -c 14 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/b_${TAG}2.log 2>&1
ls -la gpurun_out/launches_$TAG.csv gpurun_out/prof_$TAG.ncu-rep
