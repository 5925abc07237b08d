#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dist check x2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check2.log 2>&1
echo "exit $?"
grep -v "^\s*$" gpurun_out/dist_check2.log | head -60
