#!/bin/bash
# is the MMA warp's instruction stream or the tensor pipe the limiter of the active phases?  Same issue loop, MMA N = 128 / 64 / 32
cd "$(dirname "$0")/.."
CS=pixel-nerf-yolo_b200/csrc
for n in 128 64 32; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DPNR_DIAG_N=$n -c $CS/mlp_umma_pair.cu -o $CS/build/mlp_umma_pair.o 2>/dev/null
  nvcc -shared -o $CS/libpixelnerf_b200.so $CS/build/*.o -gencode arch=compute_100a,code=sm_100a
  echo "== MMA N=$n"
  PNR_PROF=1 timeout 300 python scripts/profile_field.py 8192 1 2>&1 | grep -E "mma_total|mma_wait_weights|mma_wait_chunk|mma_issue|mma_wait_gather" | head -5
done
