#!/bin/bash
# cross_kernel pipeline check: GPU parity tests, then C3 / C5 / fp64 / genea140 benches with per-layer timings.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for W in C3 C5; do
  echo "== bench $W"
  timeout 900 python bench.py --workload $W --steps 5 --cpu-seconds 0 --e2e-steps 3 --layers-json gpurun_out/layers_$W.json > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -c 1500 gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
done
echo "== bench C3 fp64"
timeout 900 python bench.py --workload C3 --numerics fp64 --steps 3 --cpu-seconds 0 --e2e-steps 0 > gpurun_out/bench_c3_fp64.json 2> gpurun_out/bench_c3_fp64.err; tail -c 600 gpurun_out/bench_c3_fp64.json; tail -3 gpurun_out/bench_c3_fp64.err
echo "== bench genea140"
timeout 300 python bench.py --workload genea140 --steps 20 --cpu-seconds 0 > gpurun_out/bench_g140.json 2> gpurun_out/bench_g140.err; tail -c 400 gpurun_out/bench_g140.json; tail -3 gpurun_out/bench_g140.err
