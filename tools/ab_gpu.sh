#!/bin/bash
# A/B of several builds of the library on the same GPU box (under gpurun), alternating runs:
#   tools/ab_gpu.sh <rounds> <libA.so> <libB.so> [...]
# prints the per-stage ms of the device-resident bench for every run.
R=$1; shift
o=gpurun_out; mkdir -p $o
for r in $(seq 1 $R); do
  for lib in "$@"; do
    FT8B_LIB_PATH=$PWD/$lib timeout 200 python bench.py --no-e2e --no-cpu > $o/ab.log 2> $o/ab.err || tail -3 $o/ab.err
    python - <<PY
import json
d = json.loads(open("$o/ab.log").read().strip().splitlines()[-1])
print("$r $lib", round(d["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, d["decode_check"])
PY
  done
done
