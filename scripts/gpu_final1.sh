#!/bin/bash
# Single-GPU acceptance run: smoke, GPU tests, default bench + reference arm, ncu launch list and full capture.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench default"
( time timeout 900 python bench.py --layers-json gpurun_out/layers_c3.json > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
tail -c 400 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real
tail -c 700 gpurun_out/bench_reference.json
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0"
echo "== ncu launch list (full C3)"
$CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
echo "== ncu full capture (full C3, one launch of each kernel)"
$CMD > gpurun_out/plain2.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'cross_kernel|couple_kernel|expand_kernel' -s 150 -c 3 -o gpurun_out/prof_full $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?"; tail -2 gpurun_out/ncu2.log
