#!/bin/bash
set -u
mkdir -p gpurun_out
for m in 0 1 2 3; do
GENLIB_EXPAND_MODE=$m timeout 600 python bench.py --workload C3 --steps 5 --cpu-seconds 0 --e2e-steps 0 --layers-json gpurun_out/layers_mode$m.json > gpurun_out/bench_mode$m.json 2> gpurun_out/bench_mode$m.err; tail -2 gpurun_out/bench_mode$m.err
done
