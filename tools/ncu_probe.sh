#!/bin/bash
# per-kernel durations (serialised, cold cache) of the network forward at a bench-sized batch
# usage: tools/ncu_probe.sh TAG [env assignments...]
TAG=$1; shift
mkdir -p gpurun_out
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_$TAG.csv \
    python tools/lstm2f_probe.py 14617 > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/ncu_$TAG.csv")) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H, body = rows[h], rows[h + 1:]
ik, iv, iu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
# the last forward of the run: from the last k_xop on
last = max(i for i, r in enumerate(body) if 'k_xop' in r[ik])
for r in body[last:]:
    t = float(r[iv].replace(',', ''))
    t = t / 1000.0 if r[iu] in ('nsecond', 'ns') else t
    print("%-60s %9.1f us" % (r[ik][:60], t))
PY
