// Design-aid micro-benchmarks and the tcgen05 building-block self test.  These are lab equipment: they are declared in
// the INTERNAL header csrc/pnr_lab.h (not in include/pixelnerf_b200.h) and used only by scripts/ and one GPU test.
#include "pnr_common.cuh"
#include "pnr_lab.h"
#include "umma.cuh"
#include <cuda.h>

namespace pnr {
using namespace umma;
namespace lab {
constexpr int kNCol = 64;
constexpr int kOperandKB = kNCol * kRowBytes;
__global__ void selftest_pack_a(const float* __restrict__ a, uint8_t* __restrict__ stages, int K) {
  uint8_t* dst = stages + (size_t)blockIdx.x * kStageBytes;
  for (int i = threadIdx.x; i < kStageRows * kBlockK; i += blockDim.x) {
    int r = i / kBlockK, k = i % kBlockK;
    *reinterpret_cast<__nv_bfloat16*>(dst + swz_offset(r, k)) = __float2bfloat16_rn(a[(size_t)r * K + blockIdx.x * kBlockK + k]);
  }
}

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const uint8_t* __restrict__ stages, const float* __restrict__ b, float* __restrict__ d, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / kBlockK;
  const uint32_t off_b = 2 * kStageBytes;                 // B operand: nkb k-blocks of 64 rows
  const uint32_t off_bar = off_b + 8 * kOperandKB;
  auto bar = [&](int i) -> uint32_t { return sbase + off_bar + 8u * i; };   // 0,1 full; 2,3 empty; 4 done
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(4), 1);
    fence_barrier_init();
  }
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + off_bar + 48);
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("pnr: dynamic smem not 1024-aligned\n"); __trap(); }
  if (warp == 1) tmem_alloc(sbase + off_bar + 48, 64);
  for (int i = threadIdx.x; i < kNCol * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    float v = r < N ? b[(size_t)r * K + k] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(smem + off_b + (k / kBlockK) * kOperandKB + swz_offset(r, k % kBlockK)) = __float2bfloat16_rn(v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    for (int s = 0; s < nkb; ++s) {
      mbar_wait(bar(2 + (s & 1)), ((s >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(s & 1), kStageBytes);
        bulk_g2s(sbase + (s & 1) * kStageBytes, stages + (size_t)s * kStageBytes, kStageBytes, bar(s & 1));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = instr_desc_bf16(N);
    for (int s = 0; s < nkb; ++s) {
      mbar_wait(bar(s & 1), (s >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        mma_kblock(tmem_base, sbase + (s & 1) * kStageBytes, sbase + off_b + s * kOperandKB, idesc, s > 0);
        mma_commit(bar(2 + (s & 1)));
      }
      __syncwarp();
    }
    if (lane == 0) mma_commit(bar(4));
    __syncwarp();
  }
  mbar_wait(bar(4), 0);
  tc_fence_after();
  {
    uint32_t r[kNCol];
    tmem_ld<kNCol>(tmem_base + ((uint32_t)(warp * 32) << 16), r);
    const int row = warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < kNCol; ++c)
      if (c < N) d[(size_t)row * N + c] = __uint_as_float(r[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}
}  // namespace lab
}  // namespace pnr

using namespace pnr;
using namespace pnr::lab;

extern "C" int pnr_umma_selftest(const float* a, const float* b, float* d, void* workspace, int N, int K, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(a && b && d && workspace, PNR_ERR_ARG, "pnr_umma_selftest: null pointer");
  PNR_REQUIRE(N >= 16 && N <= 64 && N % 16 == 0, PNR_ERR_ARG, "pnr_umma_selftest: N=%d", N);
  PNR_REQUIRE(K >= 64 && K <= 512 && K % 64 == 0, PNR_ERR_ARG, "pnr_umma_selftest: K=%d", K);
  PNR_REQUIRE(((uintptr_t)workspace & 1023) == 0, PNR_ERR_ARG, "pnr_umma_selftest: workspace alignment");
  cudaStream_t st = (cudaStream_t)stream;
  selftest_pack_a<<<K / 64, 256, 0, st>>>(a, (uint8_t*)workspace, K);
  PNR_CHECK_LAUNCH("selftest_pack_a");
  const int smem = 2 * kStageBytes + 8 * kOperandKB + 64;
  cudaError_t e = cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  selftest_kernel<<<1, 128, smem, st>>>((const uint8_t*)workspace, b, d, N, K);
  PNR_CHECK_LAUNCH("selftest_kernel");
  return PNR_OK;
}

// ---- micro-benchmark: per-SM bulk-copy ingest rate (design aid) --------------------------------------------
// One `depth`-slot ring of `stage_bytes` stages per CTA, fed by `n_prod` producer warps (warp p issues stages
// s % n_prod == p) and drained by `n_cons` consumer warps (warp c releases stages s % n_cons == c) which release a

// slot as soon as it has landed.  out[blockIdx.x] = elapsed SM cycles.
namespace pnr {
__global__ void __launch_bounds__(512, 1)
ingest_kernel(const uint8_t* __restrict__ src, int src_stages, int n_stages, int depth, long long* __restrict__ out,
              int n_prod, int n_cons, int stage_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const uint32_t bars = sbase + depth * stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * depth; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < n_prod) {
    for (int s = warp; s < n_stages; s += n_prod) {
      const uint32_t slot = s % depth, par = ((s / depth) & 1) ^ 1;
      mbar_wait(bars + 8 * (depth + slot), par);
      if (elect_one()) {
        mbar_arrive_expect_tx(bars + 8 * slot, stage_bytes);
        bulk_g2s(sbase + slot * stage_bytes, src + (size_t)(s % src_stages) * kStageBytes, stage_bytes, bars + 8 * slot);
      }
      __syncwarp();
    }
  } else if (warp < n_prod + n_cons) {
    for (int s = warp - n_prod; s < n_stages; s += n_cons) {
      const uint32_t slot = s % depth, par = (s / depth) & 1;
      mbar_wait(bars + 8 * slot, par);
      if (elect_one()) mbar_arrive(bars + 8 * (depth + slot));
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
}  // namespace pnr

extern "C" int pnr_ingest_bench(const void* src, int src_stages, int n_stages, int depth, int grid, long long* out,
                                int n_prod, int n_cons, int stage_bytes, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(src && out && depth >= 1 && grid >= 1 && n_prod >= 1 && n_cons >= 1 && n_prod + n_cons <= 16, PNR_ERR_ARG,
              "pnr_ingest_bench: bad arguments");
  PNR_REQUIRE(stage_bytes % 1024 == 0 && depth * stage_bytes <= 200 * 1024, PNR_ERR_ARG, "pnr_ingest_bench: smem");
  const int smem = depth * stage_bytes + 16 * depth + 64;
  cudaError_t e = cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  ingest_kernel<<<grid, 512, smem, (cudaStream_t)stream>>>((const uint8_t*)src, src_stages, n_stages, depth, out, n_prod,
                                                           n_cons, stage_bytes);
  PNR_CHECK_LAUNCH("ingest_kernel");
  return PNR_OK;
}

// ---- micro-benchmark 2: tensor-map TMA (cp.async.bulk.tensor.2d) ingest ------------------------------------
#include <cuda.h>
namespace pnr {
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__global__ void __launch_bounds__(64, 1)
ingest_tma_kernel(const __grid_constant__ CUtensorMap tmap, int src_stages, int n_stages, int depth, long long* __restrict__ out,
                  int box_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const uint32_t bars = sbase + depth * kStageBytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * depth; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t slot = 0, par = 1;
    for (int s = 0; s < n_stages; ++s) {
      mbar_wait(bars + 8 * (depth + slot), par);
      if (elect_one()) {
        mbar_arrive_expect_tx(bars + 8 * slot, kStageBytes);
        for (int r = 0; r < kStageRows; r += box_rows)
          tma_load_2d(sbase + slot * kStageBytes + r * kRowBytes, &tmap, 0, (s % src_stages) * kStageRows + r, bars + 8 * slot);
      }
      __syncwarp();
      if (++slot == (uint32_t)depth) { slot = 0; par ^= 1; }
    }
  } else {
    uint32_t slot = 0, par = 0;
    for (int s = 0; s < n_stages; ++s) {
      mbar_wait(bars + 8 * slot, par);
      if (elect_one()) mbar_arrive(bars + 8 * (depth + slot));
      __syncwarp();
      if (++slot == (uint32_t)depth) { slot = 0; par ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
}  // namespace pnr

extern "C" int pnr_ingest_bench_tma(const void* src, int src_stages, int n_stages, int depth, int grid, long long* out,
                                    int box_rows, int swizzle, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(src && out && depth >= 1 && depth <= 13 && grid >= 1, PNR_ERR_ARG, "pnr_ingest_bench_tma: bad arguments");
  PNR_REQUIRE(box_rows >= 8 && box_rows <= 128 && 128 % box_rows == 0, PNR_ERR_ARG, "pnr_ingest_bench_tma: box_rows");
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  PNR_REQUIRE(e == cudaSuccess && fn, PNR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {64, (cuuint64_t)src_stages * 128};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)src, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PNR_REQUIRE(r == CUDA_SUCCESS, PNR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  const int smem = depth * kStageBytes + 16 * depth + 64;
  e = cudaFuncSetAttribute(ingest_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  ingest_tma_kernel<<<grid, 64, smem, (cudaStream_t)stream>>>(tmap, src_stages, n_stages, depth, out, box_rows);
  PNR_CHECK_LAUNCH("ingest_tma_kernel");
  return PNR_OK;
}

// ---- micro-benchmark 2b: latency of the weight stream's loads (design aid) -----------------------------------------
// Cluster of 2; per iteration both CTAs synchronise, thread 0 of each issues `k` 16 KiB loads back to back and the time from the
// first issue to the last completion is accumulated.  mode 0: 1-D bulk copy, own barrier; 1: 2-D tensor-map box (64 x 128 bf16,
// SWIZZLE_128B), own barrier; 2: the field kernel's form -- cta_group::2 tensor-map loads of BOTH CTAs completing on the LEADER's
// barrier (32 KiB expected per slot), timed by the leader.  out[pair * 2 + rank] = cycles summed over the iterations.
namespace pnr {
__global__ void __launch_bounds__(64, 1)
tma_latency_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ src, int src_stages, int iters, int mode, int k,
                   long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t crank = cluster_ctarank();
  const uint32_t bars = sbase + 8 * kStageBytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();
  long long acc = 0;
  uint32_t par = 0;
  const int pair = blockIdx.x >> 1;
  for (int it = 0; it < iters; ++it) {
    cluster_sync_all();
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int j = 0; j < k; ++j) {
        const int st = (pair * 131 + it * k + j) % src_stages;
        const uint32_t dst = sbase + j * kStageBytes, b = bars + 8 * j;
        if (mode == 0) {
          mbar_arrive_expect_tx(b, kStageBytes);
          bulk_g2s(dst, src + (size_t)st * kStageBytes, kStageBytes, b);
        } else if (mode == 1) {
          mbar_arrive_expect_tx(b, kStageBytes);
          tma_load_2d(dst, &tmap, 0, st * kStageRows, b);
        } else {
          if (crank == 0) mbar_arrive_expect_tx(b, 2 * kStageBytes);
          tma_load_2d_2sm(dst, &tmap, 0, ((st & ~1) + (int)crank) * kStageRows, b);
        }
      }
      if (mode != 2 || crank == 0) {
        for (int j = 0; j < k; ++j) mbar_wait(bars + 8 * j, par);
        acc += clock64() - t0;
      }
    }
    par ^= 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
  cluster_sync_all();
}
}  // namespace pnr

extern "C" int pnr_tma_latency_bench(const void* src, int src_stages, int iters, int mode, int k, int pairs, long long* out, void* stream) {
  using namespace pnr;
  reset_launch_count();
  PNR_REQUIRE(src && out && iters > 0 && mode >= 0 && mode <= 2 && k >= 1 && k <= 8 && pairs >= 1 && src_stages >= 2, PNR_ERR_ARG,
              "pnr_tma_latency_bench: bad arguments");
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  PNR_REQUIRE(e == cudaSuccess && fn, PNR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {64, (cuuint64_t)src_stages * 128};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)src, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PNR_REQUIRE(r == CUDA_SUCCESS, PNR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  const int smem = 8 * kStageBytes + 256;
  e = cudaFuncSetAttribute(tma_latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, tma_latency_kernel, tmap, (const uint8_t*)src, src_stages, iters, mode, k, out);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "tma_latency launch: %s", cudaGetErrorString(e));
  PNR_CHECK_LAUNCH("tma_latency_kernel");
  return PNR_OK;
}

// ---- micro-benchmark 3: tcgen05.mma issue/execute cost vs shape (design aid) ---------------------------------
namespace pnr {
__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(int M, int N, int iters, int a_stride_kb, long long* __restrict__ out, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const uint32_t bar0 = sbase + 160 * 1024;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(sbase + 160 * 1024 + 64, 512);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 160 * 1024 + 64);
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t a0 = smem_desc(sbase), b0 = smem_desc(sbase + 64 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // commit_every: 0 = never, 4 = after each group of 4 MMAs (the fused kernel's per-stage pattern),
      // 104 = same but each group under its own elect/syncwarp (exactly the fused kernel's step()).
      if (commit_every == 104) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < 4; ++u) mma_bf16(tmem_base + g * 256, a0 + (uint64_t)(g * 1024) + 2 * u, b0 + 2 * u, idesc, 1);
            mma_commit(bar0 + 8);
          }
          __syncwarp();
        }
      } else if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          mma_bf16(tmem_base + (u >> 2) * 256, a0 + (uint64_t)((u >> 2) * 1024) + 2 * (u & 3), b0 + 2 * (u & 3), idesc, 1);
          if (commit_every == 4 && (u & 3) == 3) mma_commit(bar0 + 8);   // a barrier nobody waits on
        }
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(bar0);
    __syncwarp();
    mbar_wait(bar0, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}
}  // namespace pnr

extern "C" int pnr_umma_bench(int M, int N, int iters, int a_stride_kb, int grid, long long* out, int commit_every,
                              void* stream) {
  reset_launch_count();
  PNR_REQUIRE(out && (M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0, PNR_ERR_ARG, "pnr_umma_bench: bad arguments");
  const int smem = 160 * 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  umma_bench_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(M, N, iters, a_stride_kb < 1 ? 1 : a_stride_kb, out, commit_every);
  PNR_CHECK_LAUNCH("umma_bench_kernel");
  return PNR_OK;
}


// ---- micro-benchmark 3b: the CTA-PAIR MMA (cta_group::2, M = 256) in isolation, optionally under shared-memory traffic ----
// Leader thread issues iters x 8 MMAs (two groups of 4 = two "stages": A from a ring of 16 KiB slots, B from 8 KiB k-blocks), a
// multicast commit after each group when commit_every == 4.  bg: bit 0 = warps 2, 3 of each CTA stream 16 KiB bulk copies from
// global memory into a 4-slot ring (the weight stream's shared-memory writes), bit 1 = warps 4..7 keep writing 16-byte
// st.shared rows (the epilogue / gather stores).  out[pair] = cycles of the MMA loop; out[pairs + 4 * cta + w] = 8 KiB bulk copies
// loader warp w of that CTA completed meanwhile (bg bit 0; bit 2 = no MMAs at all, the loaders' baseline).
namespace pnr {
__global__ void __launch_bounds__(256, 1)
umma2_bench_kernel(int N, int iters, int commit_every, int a_slots, int bg, const uint8_t* __restrict__ src, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t crank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kA = 0, kB = 96 * 1024, kBg = 128 * 1024, kSt = 192 * 1024, kBar = 200 * 1024;
  const uint32_t bar0 = sbase + kBar;
  volatile int* done = reinterpret_cast<volatile int*>(smem + kBar + 256);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) mbar_init(bar0 + 8 * i, 1);   // 0, 1: MMA; 2..9: loader slots
    *done = 0;
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 0) tmem_alloc_2sm(sbase + kBar + 128, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kBar + 128);
  if (warp == 0) {
    if (crank == 0) {
      const uint32_t idesc = instr_desc_bf16_2sm(N);
      long long t0 = 0, t1 = 0;
      if (elect_one()) {
        t0 = clock64();
        int slot = 0;
        if (bg & 4) {                                    // baseline for the background traffic: no MMAs, same duration
          while (clock64() - t0 < (long long)iters * 8 * (N / 2)) {}
        } else
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint64_t a = smem_desc(sbase + kA + slot * 16384), b = smem_desc(sbase + kB + g * 8192);
#pragma unroll
            for (int u = 0; u < 4; ++u) mma_bf16_2sm(tmem_base + ((it * 2 + g) & 1) * 256, a + 2 * u, b + 2 * u, idesc, 1);
            if (commit_every == 4) mma_commit_2sm(bar0 + 8, 3);     // nobody waits on it
            if (++slot >= a_slots) slot = 0;
          }
        }
        mma_commit_2sm(bar0, 1);
        mbar_wait(bar0, 0);
        t1 = clock64();
        out[blockIdx.x >> 1] = t1 - t0;
      }
      __syncwarp();
      if (lane == 0) {                                   // stop the background warps of both CTAs
        *done = 1;
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa_u32(sbase + kBar + 256, 1)), "r"(1) : "memory");
      }
    }
  } else if ((bg & 1) && (warp == 2 || warp == 3 || warp == 6 || warp == 7)) {
    // four loader warps, each keeps two 8 KiB bulk copies in flight all the time (64 KiB in flight per CTA)
    if (elect_one()) {
      const int w = warp < 4 ? warp - 2 : warp - 4;        // 0..3
      uint32_t par[2] = {0, 0};
      const uint8_t* s0 = src + (size_t)w * (32 * 16384);      // every CTA streams the same 2 MiB: L2 hits, like the weight stream
      long long n = 0;
      for (int sl = 0; sl < 2; ++sl) {
        const uint32_t b = bar0 + 8 * (2 + w * 2 + sl);
        mbar_arrive_expect_tx(b, 8192);
        bulk_g2s(sbase + kBg + (w * 2 + sl) * 8192, s0 + (size_t)sl * 8192, 8192, b);
      }
      for (int i = 2; !*done; ++i) {
        const int sl = i & 1;
        const uint32_t b = bar0 + 8 * (2 + w * 2 + sl);
        mbar_wait(b, par[sl]); par[sl] ^= 1; ++n;
        mbar_arrive_expect_tx(b, 8192);
        bulk_g2s(sbase + kBg + (w * 2 + sl) * 8192, s0 + (size_t)(i & 63) * 8192, 8192, b);
      }
      for (int sl = 0; sl < 2; ++sl) mbar_wait(bar0 + 8 * (2 + w * 2 + sl), par[sl]);      // the last copies have landed
      out[(gridDim.x >> 1) + blockIdx.x * 4 + w] = n;      // 8 KiB copies completed while the MMAs ran
    }
    __syncwarp();
  } else if ((bg & 2) && (warp == 4 || warp == 5)) {
    uint32_t a = sbase + kSt + ((warp - 4) * 32 + lane) * 16;
    while (!*done) {
#pragma unroll
      for (int r = 0; r < 8; ++r) st_shared_v4(a + ((r * 2048) & 8191), lane, r, warp, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2sm(tmem_base, 512);
}
}  // namespace pnr

extern "C" int pnr_umma2_bench(int N, int iters, int commit_every, int a_slots, int bg, int pairs, const void* src, long long* out,
                               void* stream) {
  using namespace pnr;
  reset_launch_count();
  PNR_REQUIRE(out && N >= 32 && N <= 256 && N % 32 == 0 && iters > 0 && a_slots >= 1 && a_slots <= 6 && pairs >= 1, PNR_ERR_ARG,
              "pnr_umma2_bench: bad arguments");
  PNR_REQUIRE(!(bg & 1) || src, PNR_ERR_ARG, "pnr_umma2_bench: bg bit 0 needs a source buffer of pairs * 4 MiB");
  const int smem = 200 * 1024 + 512;
  cudaError_t e = cudaFuncSetAttribute(umma2_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, umma2_bench_kernel, N, iters, commit_every, a_slots, bg, (const uint8_t*)src, out);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "umma2_bench launch: %s", cudaGetErrorString(e));
  PNR_CHECK_LAUNCH("umma2_bench_kernel");
  return PNR_OK;
}

// ---- micro-benchmark: DSMEM ping-pong of `bytes` between the two CTAs of a cluster (design aid) ------------------
// mode 0: st.shared::cluster.v4 by `warps` warps + fence.proxy.async.shared::cluster + relaxed remote arrive
// mode 1: one cp.async.bulk.shared::cluster.shared::cta (TMA engine) completing on the peer's mbarrier
namespace pnr {
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(256, 1) dsmem_pingpong_kernel(int mode, int bytes, int iters, int warps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = sbase + 2 * 65536;            // [0,64K) send buffer, [64K,128K) receive buffer
  if (threadIdx.x == 0) { mbar_init(bar, mode == 0 ? warps : 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  __syncthreads();
  cluster_sync_all();
  const uint32_t peer_recv = mapa_u32(sbase + 65536, peer), peer_bar = mapa_u32(bar, peer);
  const long long t0 = clock64();
  uint32_t par = 0;
  for (int it = 0; it < iters; ++it) {
    const bool my_turn = ((it & 1) == (int)rank);    // rank 0 sends on even iterations, rank 1 on odd ones
    if (my_turn) {
      if (mode == 0) {
        if (warp < warps) {
          for (int off = (warp * 32 + lane) * 16; off < bytes; off += warps * 32 * 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(smem + off);
            st_cluster_v4(peer_recv + off, v.x, v.y, v.z, v.w);
          }
          fence_proxy_async_cluster();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(peer_bar);
        }
      } else if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx_cluster(peer_bar, bytes);
        bulk_s2s(peer_recv, sbase, bytes, peer_bar);
      }
    } else {
      mbar_wait_cluster(bar, par);
      par ^= 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  cluster_sync_all();
}
}  // namespace pnr

extern "C" int pnr_dsmem_bench(int mode, int bytes, int iters, int warps, long long* out, void* stream) {
  using namespace pnr;
  reset_launch_count();
  PNR_REQUIRE(out && bytes >= 1024 && bytes <= 65536 && bytes % 512 == 0 && warps >= 1 && warps <= 8 && iters > 0, PNR_ERR_ARG,
              "pnr_dsmem_bench: bad arguments");
  const int smem = 2 * 65536 + 64;
  cudaError_t e = cudaFuncSetAttribute(dsmem_pingpong_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, dsmem_pingpong_kernel, mode, bytes, iters, warps, out);
  PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "dsmem_pingpong_kernel launch: %s", cudaGetErrorString(e));
  PNR_CHECK_LAUNCH("dsmem_pingpong_kernel");
  return PNR_OK;
}


// ---- stand-alone entry points of the tcgen05 training GEMMs (train_umma.cu), for tests/test_gpu_train.py ---------------------
#include "train_umma.cuh"
namespace pnr {
__global__ void lab_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int cols, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const long long r = i / ld;
  const int c = (int)(i - r * ld);
  dst[i] = __float2bfloat16_rn(c < cols ? src[r * cols + c] : 0.f);
}
}  // namespace pnr

extern "C" size_t pnr_lab_gemm_workspace_bytes(long long M, int N, int K) {
  const size_t ldk = (size_t)((K + 63) / 64 * 64), ldn = (size_t)((N + 63) / 64 * 64);
  return (size_t)M * ldk * 2 + (size_t)M * ldn * 2 + pnr::tg::packed_rowgemm_bytes(N, K) + 4096;
}

// out[M, N] = epilogue(A[M, K] W[N, K]^T): A, W fp32 device inputs rounded to bf16; optional bias [N], mask_src [M, N] (fp32,
// rounded to bf16), res_in [M, N]; out_f32 and / or out_bf16_as_f32 ([M, N] fp32 copy of the bf16 output) may be null.
extern "C" int pnr_lab_rowgemm(const float* A, const float* W, const float* bias, const float* mask_src, const float* res_in,
                               float* out_f32, void* out_bf16, long long M, int N, int K, int relu_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  using namespace pnr;
  reset_launch_count();
  PNR_REQUIRE(A && W && workspace && workspace_bytes >= pnr_lab_gemm_workspace_bytes(M, N, K), PNR_ERR_ARG, "pnr_lab_rowgemm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int ldk = (K + 63) / 64 * 64, ldn = (N + 63) / 64 * 64;
  uint8_t* p = (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* A16 = (__nv_bfloat16*)p; p += ((size_t)M * ldk * 2 + 1023) & ~(size_t)1023;
  __nv_bfloat16* M16 = (__nv_bfloat16*)p; p += ((size_t)M * ldn * 2 + 1023) & ~(size_t)1023;
  uint8_t* Wp = p;
  lab_to_bf16_kernel<<<(unsigned)((M * ldk + 255) / 256), 256, 0, st>>>(A, A16, M, K, ldk);
  if (mask_src) lab_to_bf16_kernel<<<(unsigned)((M * ldn + 255) / 256), 256, 0, st>>>(mask_src, M16, M, N, ldn);
  int rc = tg::pack_rowgemm(W, N, K, K, 0, Wp, st);
  if (rc) return rc;
  tg::RowGemmArgs g = {};
  g.M = M; g.n_valid = N; g.bias = bias; g.mask_src = mask_src ? M16 : nullptr; g.ld_mask = ldn; g.res_in = res_in; g.ld_res = N;
  g.out_f32 = out_f32; g.ld_f32 = N; g.out_bf16 = (__nv_bfloat16*)out_bf16; g.ld_bf16 = N; g.relu_out = relu_out;
  tg::RowGemmSrc s0 = {A16, ldk, ldk, Wp};
  return tg::rowgemm(s0, nullptr, g, st);
}

// dW[N, K] += dY[M, N]^T X[M, K] (fp32 device inputs rounded to bf16)
extern "C" int pnr_lab_wgrad(const float* dY, const float* X, float* dW, long long M, int N, int K, void* workspace,
                             size_t workspace_bytes, void* stream) {
  using namespace pnr;
  reset_launch_count();
  PNR_REQUIRE(dY && X && dW && workspace && workspace_bytes >= pnr_lab_gemm_workspace_bytes(M, N, K), PNR_ERR_ARG, "pnr_lab_wgrad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int ldk = (K + 63) / 64 * 64, ldn = (N + 63) / 64 * 64;
  uint8_t* p = (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* X16 = (__nv_bfloat16*)p; p += ((size_t)M * ldk * 2 + 1023) & ~(size_t)1023;
  __nv_bfloat16* Y16 = (__nv_bfloat16*)p;
  lab_to_bf16_kernel<<<(unsigned)((M * ldk + 255) / 256), 256, 0, st>>>(X, X16, M, K, ldk);
  lab_to_bf16_kernel<<<(unsigned)((M * ldn + 255) / 256), 256, 0, st>>>(dY, Y16, M, N, ldn);
  return tg::wgrad(Y16, ldn, X16, ldk, dW, K, M, N, K, st);
}
