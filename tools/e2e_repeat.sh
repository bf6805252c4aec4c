#!/bin/bash
# e2e leg of bench.py a few times on one box (under gpurun): value, wall and device span of the pinned-host call
#   tools/e2e_repeat.sh <runs>
mkdir -p gpurun_out
for r in $(seq 1 ${1:-2}); do
  timeout 300 python bench.py --no-cpu --steps 3 > gpurun_out/e2e_rep.log 2> gpurun_out/e2e_rep.err || tail -3 gpurun_out/e2e_rep.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/e2e_rep.log").read().strip().splitlines()[-1])
e = d["e2e"]
print("run $r value", round(d["value"]), "e2e", round(e["value"]), "wall ms", round(e["ms_per_step"], 2), "device ms", round(e["cuda_event_ms_per_step"], 2),
      "pageable", round(e["pageable_numpy_input"]["value"]), "torch-stream", round(e["torch_stream_overlap"]["value"]), "pcm16", round(e["pcm16"]["value"]))
PY
done
