#!/bin/bash
# A/B of environment switches of the library on the same GPU box (under gpurun), alternating runs:
#   tools/ab_env.sh <rounds> "VAR=1" "VAR=0 OTHER=2" ...     ("-" = no variable)
R=$1; shift
o=gpurun_out; mkdir -p $o
for r in $(seq 1 $R); do
  for e in "$@"; do
    if [ "$e" = "-" ]; then envs=""; else envs="$e"; fi
    env $envs timeout 200 python bench.py --no-e2e --no-cpu > $o/abe.log 2> $o/abe.err || tail -3 $o/abe.err
    python - <<PY
import json
d = json.loads(open("$o/abe.log").read().strip().splitlines()[-1])
print("$r [$e]", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, d["decode_check"])
PY
  done
done
