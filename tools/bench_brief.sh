#!/bin/bash
# prints the headline numbers and stage times of one bench.py run
python bench.py --no_cpu_baseline "$@" 2>&1 | grep '^{"metric"' | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.read())
print("value %.0f  e2e %.0f  ms/step %.4f  launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
print({k: round(v, 4) for k, v in d["stage_ms"].items()})
print("roofline %.4f  count %.4f  clocks %s" % (d["roofline"]["frac"], d["roofline_count"]["frac"], d["clocks"]))
'
