#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-4}
for G in 0 1; do
echo "== bench C3 x$N NO_GUESTS=$G"
GENLIB_NO_GUESTS=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 0 --e2e-steps 1 --layers-json gpurun_out/layers_c3_x${N}_ng$G.json > gpurun_out/bench_c3_x${N}_ng$G.json 2> gpurun_out/bench_c3_x${N}_ng$G.err
echo "exit $?"; grep -v "^\s*$" gpurun_out/bench_c3_x${N}_ng$G.err | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -5
done
