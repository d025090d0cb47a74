#!/bin/bash
# per-kernel durations of one timed bench pass (config 2), serialised / cold cache: gpurun_out/launches_TAG.csv + table
TAG=${1:-x}; shift
mkdir -p gpurun_out
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no_cpu_baseline --cfg5_scale 0 > gpurun_out/b_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H, body = rows[h], rows[h + 1:]
ik, iv, iu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
starts = [i for i, r in enumerate(body) if 'k_cigar' in r[ik]]
seg = body[starts[2]:starts[3]] if len(starts) > 3 else body[starts[-1]:]
tot = 0
for r in seg:
    t = float(r[iv].replace(',', ''))
    t = t / 1000.0 if r[iu] in ('nsecond', 'ns') else t
    tot += t
    print("%-70s %9.1f us" % (r[ik][:70], t))
print("total %.1f us" % tot)
PY
