#!/bin/bash
set -u
mkdir -p gpurun_out
N=8
free -g | head -2; nproc
echo "== dist check x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check$N.log 2>&1
echo "exit $?"; grep -E "dist x|dist_check|rror" gpurun_out/dist_check$N.log | head -20
echo "== bench C3 x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 0 --e2e-steps 2 --layers-json gpurun_out/layers_c3_x$N.json > gpurun_out/bench_c3_x$N.json 2> gpurun_out/bench_c3_x$N.err
echo "exit $?"; tail -c 2200 gpurun_out/bench_c3_x$N.json; grep -v "^\s*$" gpurun_out/bench_c3_x$N.err | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -8
echo "== bench C4 x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload C4 --steps 3 --warmup 3 --cpu-seconds 0 --e2e-steps 1 --layers-json gpurun_out/layers_c4_x$N.json > gpurun_out/bench_c4_x$N.json 2> gpurun_out/bench_c4_x$N.err
echo "exit $?"; tail -c 2200 gpurun_out/bench_c4_x$N.json; grep -v "^\s*$" gpurun_out/bench_c4_x$N.err | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -8
