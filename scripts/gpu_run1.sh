#!/bin/bash
# First GPU contact: environment, smoke under compute-sanitizer, GPU tests, first bench lines.
set -u
mkdir -p gpurun_out
{
  nvidia-smi -L; nproc; free -g | head -2
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.limit --format=csv
} > gpurun_out/env.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== sanitizer (memcheck) on smoke"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?"; tail -4 gpurun_out/sanitizer.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench C3 scale 0.25"
timeout 600 python bench.py --workload C3 --scale 0.25 --steps 3 --cpu-seconds 5 --layers-json gpurun_out/layers_c3_q.json > gpurun_out/bench_c3_q.json 2> gpurun_out/bench_c3_q.err; tail -c 3000 gpurun_out/bench_c3_q.json; tail -3 gpurun_out/bench_c3_q.err
echo "== bench C3 full"
timeout 900 python bench.py --workload C3 --steps 5 --cpu-seconds 20 --layers-json gpurun_out/layers_c3.json > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 3000 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
echo "== bench genea140"
timeout 300 python bench.py --workload genea140 --steps 5 --cpu-seconds 0 --layers-json gpurun_out/layers_g140.json > gpurun_out/bench_g140.json 2> gpurun_out/bench_g140.err; tail -c 1500 gpurun_out/bench_g140.json; tail -3 gpurun_out/bench_g140.err
