#!/bin/bash
# A/B of several library builds on the drift-hypothesis bench (under gpurun): tools/ab_drift.sh <rounds> <libA.so> ...
R=$1; shift
for r in $(seq 1 $R); do
  for lib in "$@"; do
    FT8B_LIB_PATH=$PWD/$lib timeout 120 python tools/bench_drift.py 512 28 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$r $lib', round(d['value']), 'slots/s  sync', round(d['sync_drift_us_per_slot'],2), 'us/slot  walk', round(d['sync_drift_walk_us_per_slot'],2), 'blocksum', d['sync_drift_blocksum_us_per_slot'], 'direct', round(d['sync_drift_direct_us_per_slot'],1), d['tiled_vs_direct'], d['slots_with_tx_payload_decoded'])"
  done
done
