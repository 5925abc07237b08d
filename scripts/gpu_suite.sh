#!/bin/bash
# Single-GPU acceptance run: smoke, GPU tests, default bench and the reference arm.  Every step has its
# own timeout (a hung kernel must not eat the GPU budget).  gpurun -- 'bash scripts/gpu_suite.sh'
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench default"
( time timeout 300 python bench.py --layers-json gpurun_out/layers_c3.json > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
tail -c 600 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm (one full pass of the oracle)"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real
tail -c 900 gpurun_out/bench_reference.json
for w in C5 genea140; do
  echo "== bench $w"; timeout 300 python bench.py --workload $w --steps 10 --cpu-seconds 0 --layers-json gpurun_out/layers_$w.json > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.json
done
