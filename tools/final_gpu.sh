#!/bin/bash
# Round-end verification on the GPU box (under gpurun): GPU parity tests, smoke, the drift A/B bench, then the
# every-kernel smoke passes (written for compute-sanitizer memcheck / racecheck where that tool is available).
#   tools/final_gpu.sh <tag>
tag=${1:-fin}
o=gpurun_out
mkdir -p $o
timeout 420 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log
tail -3 $o/${tag}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> $o/${tag}_smoke.log
tail -2 $o/${tag}_smoke.log
timeout 150 python tools/bench_drift.py 512 28 3 > $o/${tag}_drift.log 2>&1; echo "exit $?" >> $o/${tag}_drift.log
tail -2 $o/${tag}_drift.log | cut -c1-1500
# compute-sanitizer is closed on this GPU pool (its runs left GPUs needing a reset), so the two every-kernel passes run plain:
# they still exercise every kernel and geometry on small inputs and fail on a sticky CUDA error.
timeout 200 python tools/sanitize_smoke.py > $o/${tag}_allkernels.log 2>&1; echo "exit $?" >> $o/${tag}_allkernels.log
tail -3 $o/${tag}_allkernels.log
timeout 200 python tools/racecheck_smoke.py > $o/${tag}_allkernels_small.log 2>&1; echo "exit $?" >> $o/${tag}_allkernels_small.log
tail -3 $o/${tag}_allkernels_small.log
