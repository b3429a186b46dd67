#!/bin/bash
# weight-ring depth experiment: rebuild the pair kernel with 3 / 4 / 5 stages and time the bench workload
cd "$(dirname "$0")/.."
CS=pixel-nerf-yolo_b200/csrc
for n in 3 4 5; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DPNR_STAGES=$n -c $CS/mlp_umma_pair.cu -o $CS/build/mlp_umma_pair.o 2>/dev/null
  nvcc -shared -o $CS/libpixelnerf_b200.so $CS/build/*.o -gencode arch=compute_100a,code=sm_100a
  echo "== stages $n"
  PNR_PROF=1 timeout 300 python scripts/profile_field.py 8192 1 2>&1 | grep -E "mma_total|mma_wait_weights|mma_wait_chunk|mma_issue" | head -4
done
