#!/bin/bash
# Multi-GPU validation: NCCL parity tests, then the strong-scaling bench at the given GPU counts.  Usage: gpu_multi_r2.sh TAG "2 4 8"
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r02m}
NS=${2:-2}
timeout 900 python -m pytest tests/test_gpu_z_multi.py -m gpu -q --timeout 800 > gpurun_out/test_multi_$TAG.log 2>&1
echo "multi tests exit $? $(tail -1 gpurun_out/test_multi_$TAG.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/test_multi_$TAG.log | head -20
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_n1_$TAG.log 2> gpurun_out/bench_n1_$TAG.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-extras > gpurun_out/bench_n${N}_$TAG.log 2> gpurun_out/bench_n${N}_$TAG.err
  fi
  echo "bench N=$N exit $?"; tail -1 gpurun_out/bench_n${N}_$TAG.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print(' value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'parity', d['parallelism']['sharded_equals_single_gpu'], 'launches', d['gpu_launches'], 'kernel ms', round(d['roofline']['kernel_ms_per_step'], 3))
except Exception as e:
    print(' (no json)', e)
"
done
