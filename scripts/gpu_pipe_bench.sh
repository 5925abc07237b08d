#!/bin/bash
# bench-only A/B of GENLIB_PIPE on N GPUs (parity of both settings is checked by gpu_pipe.sh)
set -u
mkdir -p gpurun_out
N=${1:-4}
for P in 0 1; do
  GENLIB_PIPE=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$P bench.py --gpus $N --steps 10 --warmup 3 --cpu-seconds 0 --e2e-steps 0 --layers-json gpurun_out/layers_pipe$P.json > gpurun_out/bench_pipe$P.json 2> gpurun_out/bench_pipe$P.err
  echo "pipe $P exit $?"
done
