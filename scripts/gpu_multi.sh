#!/bin/bash
# 2-GPU run: N-GPU == 1-GPU determinism test + scaling bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/_nccl_worker.py > gpurun_out/nccl_worker.log 2>&1
echo "nccl worker exit $?"; tail -3 gpurun_out/nccl_worker.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_n{n}.log").read().strip().splitlines() if l.startswith("{")][-1])
        print(n, "gpus: rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "parse fail", e)
PY
