#!/bin/bash
# compare the synchronous and the st.async epilogue exchange of the CTA-pair kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for a in ${MODES:-1 0}; do
  PNR_ASYNC=$a timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "field or render or edge or full_image or latent_widths or yolo" > gpurun_out/test_async$a.log 2>&1
  echo "ASYNC=$a tests exit $? $(tail -1 gpurun_out/test_async$a.log)"
  grep -E "pnr:|Error|FAILED|assert " gpurun_out/test_async$a.log | head -6
  PNR_ASYNC=$a timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_async$a.log 2>gpurun_out/bench_async$a.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_async$a.log").read().strip().splitlines()[-1])
    print("ASYNC=$a rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("ASYNC=$a parse fail", e); print(open("gpurun_out/bench_async$a.err").read()[-1200:])
PY
  PNR_ASYNC=$a PNR_PROF=1 timeout 300 python scripts/profile_field.py 8192 1 2>&1 | grep -E "mma_|gather_|epi_" | head -12
done
