#!/bin/bash
# 2-GPU box: GPU parity tests, 1-GPU C3 bench, 2-GPU parity check + C3 bench.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench C3 x1"
timeout 900 python bench.py --workload C3 --steps 5 --cpu-seconds 0 --e2e-steps 3 --layers-json gpurun_out/layers_C3.json > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; tail -c 300 gpurun_out/bench_C3.json; tail -3 gpurun_out/bench_C3.err
bash scripts/gpu_dist3.sh 2
