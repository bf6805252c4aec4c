#!/bin/bash
# Sweep one environment switch of the library on one GPU box: tools/env_sweep.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --no-e2e --no-cpu 2>/dev/null > gpurun_out/sweep.log
  python - <<PY
import json
d = json.loads(open("gpurun_out/sweep.log").read().strip().splitlines()[-1])
print("$var=$v", round(d["value"]), {k: round(x["ms_per_step"], 3) for k, x in d["stages"].items()})
PY
done
