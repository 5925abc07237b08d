#!/bin/bash
# Parity + bench for both expand staging variants, then ncu on the default.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for st in 2 1; do
echo "== bench C3 full stages=$st"
GENLIB_EXPAND_STAGES=$st timeout 900 python bench.py --workload C3 --steps 10 --cpu-seconds 0 --layers-json gpurun_out/layers_c3_s$st.json > gpurun_out/bench_c3_s$st.json 2> gpurun_out/bench_c3_s$st.err; tail -c 1500 gpurun_out/bench_c3_s$st.json; tail -3 gpurun_out/bench_c3_s$st.err
done
echo "== bench genea140"
timeout 300 python bench.py --workload genea140 --steps 20 --cpu-seconds 0 --layers-json gpurun_out/layers_g140.json > gpurun_out/bench_g140.json 2> gpurun_out/bench_g140.err; tail -c 600 gpurun_out/bench_g140.json; tail -3 gpurun_out/bench_g140.err
echo "== bench C5 scale 0.25"
timeout 600 python bench.py --workload C5 --scale 0.25 --steps 5 --cpu-seconds 0 --layers-json gpurun_out/layers_c5q.json > gpurun_out/bench_c5q.json 2> gpurun_out/bench_c5q.err; tail -c 600 gpurun_out/bench_c5q.json; tail -3 gpurun_out/bench_c5q.err
CMD="python bench.py --workload C3 --scale 0.25 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0"
echo "== ncu full capture"
$CMD > gpurun_out/plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'cross_kernel|couple_kernel|expand_kernel' -s 90 -c 3 -o gpurun_out/prof $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?"; tail -3 gpurun_out/ncu2.log
