#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
echo "== dist check x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check$N.log 2>&1
echo "exit $?"; grep -E "dist x|dist_check|rror" gpurun_out/dist_check$N.log | grep -v ": ok" | head -20
echo "== bench C3 x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 0 --e2e-steps 3 --layers-json gpurun_out/layers_c3_x$N.json > gpurun_out/bench_c3_x$N.json 2> gpurun_out/bench_c3_x$N.err
echo "exit $?"; grep -v "^\s*$" gpurun_out/bench_c3_x$N.err | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -8
