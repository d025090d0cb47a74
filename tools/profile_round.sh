#!/bin/bash
# ncu evidence of one bench pass (config 2): launch list + one `--set full` capture of every kernel of the timed pass.
# usage: tools/profile_round.sh TAG      -> gpurun_out/launches_TAG.csv, gpurun_out/prof_TAG.ncu-rep, gpurun_out/b_TAG*.log
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no_cpu_baseline --cfg5_scale 0 > gpurun_out/b_$TAG.log 2>&1
# the pass has 26 launches (config 2, C = 18, no padding): skip the first submit and the warm-up pass
N=$(python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
body = rows[h + 1:]
ik = rows[h].index('Kernel Name')
s = [i for i, r in enumerate(body) if 'k_cigar' in r[ik]]
print(s[2], s[3] - s[2])
PY
)
set -- $N
ncu --set full --clock-control none --import-source on -s $1 -c $2 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no_cpu_baseline --cfg5_scale 0 > gpurun_out/b_${TAG}2.log 2>&1
ls -la gpurun_out/launches_$TAG.csv gpurun_out/prof_$TAG.ncu-rep
