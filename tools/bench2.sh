#!/bin/bash
# two-rank bench + reference arm, as the driver launches them
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no_cpu_baseline 2>&1 | grep '^{"metric"' | python -c '
import json, sys
d = json.loads(sys.stdin.read())
print("n_gpus %d value %.0f e2e %.0f ms/step %.4f clocks %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]))'
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --impl reference --steps 1 --warmup 0 2>&1 | grep '^{"impl"' | cut -c1-160
