#!/bin/bash
# N GPUs of one box: tests/dist_check.py (N ranks == 1 GPU == oracle, bit for bit), the one-process
# multi-device call, then the C3 bench under torchrun (and C4 on 8 GPUs).  gpurun --gpus N -- 'bash scripts/gpu_dist.sh N'
set -u
mkdir -p gpurun_out
N=${1:-2}
echo "== dist check x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check$N.log 2>&1
echo "exit $?"; grep -E "dist x|rror" gpurun_out/dist_check$N.log | grep -v ": ok" | head
echo "== one process, $N devices"; timeout 300 python -m pytest tests/test_gpu_parity.py::test_one_process_several_devices tests/test_abi.py::test_plain_c_client_on_gpu -m gpu -x -q 2>&1 | tail -2
run() { W=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $W "$@" --cpu-seconds 0 --layers-json gpurun_out/layers_${W}_x$N.json > gpurun_out/bench_${W}_x$N.json 2> gpurun_out/bench_${W}_x$N.err
  echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_x$N.json'))
print('$W x$N ms_per_step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_call'],1), d['e2e']['breakdown_ms'], 'golden', d['parity']['matches_golden'], 'frac', round(d['roofline']['frac'],3), 'survey-model frac', round(d['roofline']['survey_model']['frac'],3))"; }
echo "== bench C3 x$N"; run C3 --steps 10 --warmup 3 --e2e-steps 3
if [ "$N" = "8" ]; then echo "== bench C4 x$N"; run C4 --steps 3 --warmup 3 --e2e-steps 1; fi
