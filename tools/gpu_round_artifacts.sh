#!/bin/bash
# Runs on the GPU box (under gpurun): the evidence set of one round, written to gpurun_out/<tag>_*.
#   tools/gpu_round_artifacts.sh r43
# Order matters: every command runs plain first and is only repeated under ncu when it exited 0.
tag=${1:-rX}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log
timeout 900 python bench.py > $o/${tag}_bench_default.log 2>&1; echo "exit $?" >> $o/${tag}_bench_default.log
timeout 900 python bench.py --impl reference > $o/${tag}_bench_reference.log 2>&1; echo "exit $?" >> $o/${tag}_bench_reference.log
timeout 600 python tools/bench_ldpc.py 1000000 5 > $o/${tag}_ldpc_c5.log 2>&1
K='regex:stft|sync_scores|select_candidates|llr_kernel|decode_kernel'
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $o/${tag}_bench_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $o/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $o/${tag}_ncu_launches.log 2>&1
python tools/profile_step.py 256 2 > $o/${tag}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k 'regex:stft1920|sync_scores|llr_kernel|decode_kernel|select_cand' -s 5 -c 5 \
      -f -o $o/${tag}_prof python tools/profile_step.py 256 2 > $o/${tag}_ncu_full.log 2>&1
tail -2 $o/${tag}_pytest_gpu.log
tail -c 600 $o/${tag}_bench_default.log
