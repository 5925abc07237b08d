#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench C5 full"
timeout 900 python bench.py --workload C5 --steps 5 --cpu-seconds 10 --layers-json gpurun_out/layers_c5.json > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
echo "== bench C3 fp64 storage"
timeout 900 python bench.py --workload C3 --numerics fp64 --steps 5 --cpu-seconds 0 > gpurun_out/bench_c3_fp64.json 2> gpurun_out/bench_c3_fp64.err; tail -c 300 gpurun_out/bench_c3_fp64.json; tail -3 gpurun_out/bench_c3_fp64.err
echo "== bench genea140"
timeout 300 python bench.py --workload genea140 --steps 20 --cpu-seconds 30 > gpurun_out/bench_g140.json 2> gpurun_out/bench_g140.err; tail -c 300 gpurun_out/bench_g140.json; tail -3 gpurun_out/bench_g140.err
