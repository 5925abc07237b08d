#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench C3 full"
timeout 900 python bench.py --workload C3 --steps 10 --cpu-seconds 0 --e2e-steps 4 --layers-json gpurun_out/layers_c3.json > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1200 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
echo "== bench genea140"
timeout 300 python bench.py --workload genea140 --steps 20 --cpu-seconds 0 --e2e-steps 4 > gpurun_out/bench_g140.json 2> gpurun_out/bench_g140.err; tail -c 700 gpurun_out/bench_g140.json; tail -3 gpurun_out/bench_g140.err
