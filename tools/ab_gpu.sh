#!/bin/bash
# A/B of two builds of the library on the same GPU box (under gpurun), alternating runs:
#   tools/ab_gpu.sh <libA.so> <libB.so> [rounds]
# prints the per-stage ms of the device-resident bench for every run.
A=$1; B=$2; R=${3:-2}
o=gpurun_out; mkdir -p $o
for r in $(seq 1 $R); do
  for tag in A B; do
    lib=$A; [ $tag = B ] && lib=$B
    FT8B_LIB_PATH=$PWD/$lib timeout 200 python bench.py --no-e2e --no-cpu > $o/ab_$tag$r.log 2> $o/ab_$tag$r.err || tail -3 $o/ab_$tag$r.err
    python - <<PY
import json
d = json.loads(open("$o/ab_$tag$r.log").read().strip().splitlines()[-1])
print("$tag$r $lib", round(d["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
PY
  done
done
