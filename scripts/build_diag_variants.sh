#!/bin/bash
# Timing-diagnostic builds of the pair kernel (WRONG RESULTS, timing only): each variant removes one cost so that the step time
# without it bounds what a redesign of that part could gain.  Built here (nvcc needs ~2 min per variant), measured with
# scripts/gpu_diag_bounds.sh on the GPU box.  The product library is never replaced in the repository.
cd "$(dirname "$0")/.."
CS=pixel-nerf-yolo_b200/csrc
D=$CS/build/diag
mkdir -p $D
python -m pixel_nerf_yolo_b200.build > /dev/null || exit 1      # product objects (all other translation units)
build_one() {   # name, flags...
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $CS/mlp_umma_pair.cu -o $D/pair_$name.o 2> $D/pair_$name.log || { echo "compile $name failed"; return 1; }
  objs=$(ls $CS/build/*.o | grep -v mlp_umma_pair.o)
  nvcc -shared -o $D/lib_$name.so $objs $D/pair_$name.o -gencode arch=compute_100a,code=sm_100a && echo "built $D/lib_$name.so"
}
if [ -n "$ONLY" ]; then
  for v in $ONLY; do
    case $v in
      n64) build_one n64 -DPNR_DIAG_N=64 & ;;
      n32) build_one n32 -DPNR_DIAG_N=32 & ;;
      noepi) build_one noepi -DPNR_DIAG_NOEPI & ;;
      noring) build_one noring -DPNR_DIAG_NORING & ;;
      noepi_n32) build_one noepi_n32 -DPNR_DIAG_NOEPI -DPNR_DIAG_N=32 & ;;
      mmaonly) build_one mmaonly -DPNR_DIAG_MMAONLY & ;;
      mmaonly_n32) build_one mmaonly_n32 -DPNR_DIAG_MMAONLY -DPNR_DIAG_N=32 & ;;
      ringonly) build_one ringonly -DPNR_DIAG_OFF_WAITS -DPNR_DIAG_OFF_EPI -DPNR_DIAG_OFF_GATHER & ;;
      ringonly_n32) build_one ringonly_n32 -DPNR_DIAG_OFF_WAITS -DPNR_DIAG_OFF_EPI -DPNR_DIAG_OFF_GATHER -DPNR_DIAG_N=32 & ;;
      idle1) build_one idle1 -DPNR_IDLE_WAIT=1 & ;;
      idle2) build_one idle2 -DPNR_IDLE_WAIT=2 & ;;
      idle3) build_one idle3 -DPNR_IDLE_WAIT=3 & ;;
      epi_nostore) build_one epi_nostore -DPNR_DIAG_EPI_NOSTORE & ;;
      epi_noalu) build_one epi_noalu -DPNR_DIAG_EPI_NOALU & ;;
      ring_gather) build_one ring_gather -DPNR_DIAG_OFF_WAITS -DPNR_DIAG_OFF_EPI & ;;
      stages4) build_one stages4 -DPNR_STAGES=4 & ;;
      stages3) build_one stages3 -DPNR_STAGES=3 & ;;
      nogather) build_one nogather -DPNR_DIAG_NOGATHER & ;;
      *) echo "unknown variant $v" ;;
    esac
  done
  wait
  ls -la $D/*.so
  exit 0
fi
build_one noweights -DPNR_DIAG_NOWEIGHTS &
build_one noepi -DPNR_DIAG_NOEPI &
build_one nogather -DPNR_DIAG_NOGATHER &
build_one noweights_noepi -DPNR_DIAG_NOWEIGHTS -DPNR_DIAG_NOEPI &
build_one none3 -DPNR_DIAG_NOWEIGHTS -DPNR_DIAG_NOEPI -DPNR_DIAG_NOGATHER &
wait
build_one noring -DPNR_DIAG_NORING &
build_one noring_noepi -DPNR_DIAG_NORING -DPNR_DIAG_NOEPI &
build_one noring_none3 -DPNR_DIAG_NORING -DPNR_DIAG_NOEPI -DPNR_DIAG_NOGATHER &
wait
ls -la $D/*.so
