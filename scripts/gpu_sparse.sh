#!/bin/bash
# sparse_phi row (N2) on the GPU + the full GPU suite + a C3 bench to confirm the phi path is unchanged.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench C3"
timeout 900 python bench.py --workload C3 --steps 5 --cpu-seconds 0 --e2e-steps 3 --layers-json gpurun_out/layers_C3.json > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; tail -c 400 gpurun_out/bench_C3.json; tail -3 gpurun_out/bench_C3.err
