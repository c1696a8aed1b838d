#!/bin/bash
# N-GPU pass (gpurun --gpus N): the real NCCL sharding test, then the default bench line at N ranks (scene block with canvas hashes).
set -u
N=${1:-2}
T=${2:-r2d}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/smi_${N}gpu_$T.txt
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s 2>&1 | tail -15 > $O/pytest_multirank_${N}gpu_$T.log; echo "multirank rc=$? $(tail -1 $O/pytest_multirank_${N}gpu_$T.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 \
    > $O/bench_default_${N}gpu_$T.json 2> $O/bench_default_${N}gpu_$T.err; echo "bench N=$N rc=$?"
tail -3 $O/bench_default_${N}gpu_$T.err
