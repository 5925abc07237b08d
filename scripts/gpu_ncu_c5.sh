#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload C5 --steps 1 --warmup 3 --cpu-seconds 0 --e2e-steps 0"
$CMD > gpurun_out/plain_c5.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'cross_kernel|mirror_kernel' -s 400 -c 2 -o gpurun_out/prof_c5 $CMD > gpurun_out/ncu_c5.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_c5.log
