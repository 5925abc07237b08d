#!/bin/bash
# ncu launch list + full capture of one mid-pedigree launch of each layer kernel (full C3), after a plain run.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'cross_kernel|couple_kernel|expand_kernel' -s 150 -c 3 -o gpurun_out/prof_full2 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?"; tail -2 gpurun_out/ncu2.log
