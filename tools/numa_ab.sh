#!/bin/bash
# A/B of the host NUMA binding on an N-GPU box (under gpurun --gpus N): e2e slots/s with and without it.
#   tools/numa_ab.sh N
n=${1:-4}
o=gpurun_out
mkdir -p $o
nvidia-smi topo -m > $o/topo_${n}.txt 2>&1
for d in /sys/devices/system/node/node*; do echo "$d $(cat $d/cpulist)"; done >> $o/topo_${n}.txt 2>&1
nproc >> $o/topo_${n}.txt
for mode in 0 1; do
  FT8B_NUMA_BIND=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 29517 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu > $o/numa_${n}_bind${mode}.log 2> $o/numa_${n}_bind${mode}.err
  python - <<PY
import json
try:
    d = json.loads(open("$o/numa_${n}_bind${mode}.log").read().strip().splitlines()[-1])
    print("bind=$mode value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pcm16", round(d["e2e"]["pcm16"]["value"]), d.get("host_numa_binding"))
except Exception as e:
    print("bind=$mode failed", e)
PY
done
tail -3 $o/numa_${n}_bind1.err
cat $o/topo_${n}.txt | head -30
