#!/bin/bash
# A/B of GENLIB_PIPE (couple_kernel of one couple group beside cross_kernel of the next) on N GPUs.
set -u
mkdir -p gpurun_out
N=${1:-2}
for P in 0 1; do
  echo "== GENLIB_PIPE=$P: dist check + bench C3 x$N"
  GENLIB_PIPE=$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$P tests/dist_check.py 2>&1 | grep -E "dist_check|MISMATCH|rror" | head -5
  GENLIB_PIPE=$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$P bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 0 --e2e-steps 1 --layers-json gpurun_out/layers_pipe$P.json > gpurun_out/bench_pipe$P.json 2> gpurun_out/bench_pipe$P.err
  echo "exit $?"; grep -v "^\s*$" gpurun_out/bench_pipe$P.err | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -4
done
