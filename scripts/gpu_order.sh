#!/bin/bash
# compare the (chunk, tile) pair orders of the CTA-pair kernel (PNR_ORDER 0/1/2)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for o in ${ORDERS:-2 1}; do
  PNR_ORDER=$o timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "field and bf16 or render_bf16 or yolo" > gpurun_out/test_order$o.log 2>&1
  echo "ORDER=$o tests exit $? $(tail -1 gpurun_out/test_order$o.log)"
  PNR_ORDER=$o timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_order$o.log 2>gpurun_out/bench_order$o.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_order$o.log").read().strip().splitlines()[-1])
    print("ORDER=$o rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("ORDER=$o parse fail", e); print(open("gpurun_out/bench_order$o.err").read()[-800:])
PY
  PNR_ORDER=$o PNR_PROF=1 python scripts/profile_field.py 8192 1 2>&1 | grep -E "mma_" | head -5
done
