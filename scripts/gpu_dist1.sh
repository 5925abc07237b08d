#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== single-GPU regression"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== dist check x2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py 2>&1 | tail -30 | tee gpurun_out/dist_check2.log
