#!/bin/bash
# `ncu --set full` of the launches of one timed pass whose names match a regex: gpurun_out/prof_TAG.ncu-rep
# usage: tools/ncu_kernels.sh TAG 'k_cigar|k_cmp|k_rows' [env assignments...]
TAG=$1; RE=$2; shift; shift
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $((2 * $(echo "$RE" | tr '|' '\n' | wc -l))) -c $(echo "$RE" | tr '|' '\n' | wc -l) \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no_cpu_baseline --cfg5_scale 0 > gpurun_out/b_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
