#!/bin/bash
# how much of the step does the gather cost?  build the pair kernel with the tap loads removed (wrong results, timing only)
cd "$(dirname "$0")/.."
CS=pixel-nerf-yolo_b200/csrc
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('normal   ms/step', round(d['ms_per_step'],2), d['clocks'])"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DPNR_DIAG_NOGATHER -c $CS/mlp_umma_pair.cu -o $CS/build/mlp_umma_pair.o 2>/dev/null
nvcc -shared -o $CS/libpixelnerf_b200.so $CS/build/*.o -gencode arch=compute_100a,code=sm_100a
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-gather ms/step', round(d['ms_per_step'],2), d['clocks'])"
