#!/bin/bash
# N-GPU evidence on one box (under gpurun --gpus N): bench.py and the C5 LDPC bench under torchrun.
#   tools/multi_gpu_artifacts.sh N tag
n=${1:-2}; tag=${2:-rX}
o=gpurun_out; mkdir -p $o
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 "$@"; }
timeout 900 bash -c "$(declare -f run); n=$n; run bench.py --gpus $n --steps 3 --warmup 3 --no-cpu" > $o/${tag}_bench_${n}gpu.log 2> $o/${tag}_bench_${n}gpu.err; echo "exit $?" >> $o/${tag}_bench_${n}gpu.err
timeout 600 bash -c "$(declare -f run); n=$n; run tools/bench_ldpc.py 1000000 5" > $o/${tag}_ldpc_${n}gpu.log 2> $o/${tag}_ldpc_${n}gpu.err; echo "exit $?" >> $o/${tag}_ldpc_${n}gpu.err
tail -c 700 $o/${tag}_bench_${n}gpu.log; echo; tail -c 600 $o/${tag}_ldpc_${n}gpu.log; tail -2 $o/${tag}_ldpc_${n}gpu.err
