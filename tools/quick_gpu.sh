#!/bin/bash
# Quick GPU check of a kernel change (under gpurun): GPU parity tests, then the device-resident bench with per-stage times.
#   tools/quick_gpu.sh <tag> [extra pytest args]
tag=${1:-q}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1
tail -4 $o/${tag}_pytest.log
timeout 200 python bench.py --no-e2e --no-cpu > $o/${tag}_bench.log 2> $o/${tag}_bench.err || tail -5 $o/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("$o/${tag}_bench.log").read().strip().splitlines()[-1])
print(round(d["value"]), "slots/s", round(d["ms_per_step"], 3), "ms/step")
print({k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
PY
