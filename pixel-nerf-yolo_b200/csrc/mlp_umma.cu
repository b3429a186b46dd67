// F2+F3 fused: PixelNeRFNet.forward (src/model/models.py:153-318) as ONE persistent sm_100a kernel.
//
// Tile = 64 "columns" = PP points x NS source views (PP = 64/NS).  The MLP is evaluated feature-major:
//     D^T[512 features, 64 columns] (+)= W[512, K] * act^T[K, 64]
// so the nn.Linear weight matrix (out x in, row-major) is the K-major A operand as it is stored, and an
// activation row (one point-view, K contiguous) is the K-major B operand.  tcgen05.mma M=128 (4 feature
// tiles), N=64 (pre-combine) or N=round16(PP) (after the view mean), K=16, bf16 in, fp32 accumulate.
//
//   TMEM   cols [0,256)   x^T : the fp32 residual stream of the tile, 4 feature tiles x 64 columns.
//                               `x += lin_z(latent)` and `x += fc_1(relu(h))` are MMAs that accumulate in place
//                               (resnetfc.py:176-182, 53-62) -- the residual add costs nothing.
//          cols [256,512) h^T : the four 128x64 fc_0 accumulators.
//   SMEM   weight ring 5 x 16 KiB  <- cp.async.bulk (TMA engine) from the packed, pre-swizzled stream
//          relu(x) K-chunk ring 2 x 16 KiB, relu(h) K-chunk ring 2 x 16 KiB: both 512x512 layers run K-chunk-outer,
//          so a 128-feature chunk of the operand is consumed by the tensor core while the epilogue warps convert
//          the next one (only the first chunk of a layer is exposed);
//          latent operand 64 KiB (gathered ONCE per tile, reused by lin_z[0..2]), z-feature operand 8 KiB.
//   Warps  0,12-14: weight producers   1: MMA issuer (+TMEM alloc)   4-7: epilogue (TMEM -> bias/ReLU/bf16 -> smem,
//          view mean, final sigmoid/relu)   2,3,8-11: project + 4-tap gather + positional encoding.
//
// Measured facts that shaped this (scripts/ingest_bench.py, B200): an mbarrier-mediated hand-off costs a warp
// ~350-540 cycles per stage as a dependent chain (try_wait -> expect_tx -> bulk copy: 538; try_wait -> arrive: 353)
// whatever the stage size up to 32 KiB, while hand-offs of different warps overlap; one producer + one consumer
// therefore top out near 30 B/clk/SM, far below what the tensor core eats at N=64 (128 B/clk).  Hence several
// producer warps, and an MMA warp that probes all ring slots at once (one lane per slot + ballot).
//
// Biases never touch the tensor pipe: the TMEM accumulator holds x minus the biases added so far, and the
// epilogue adds the cumulative bias vector (precomputed at pack time) when it forms relu(x).
#include "pnr_common.cuh"
#include "umma.cuh"
#include <stdlib.h>
#include <vector>

namespace pnr {
using namespace umma;

int validate_scene_points(const pnr_scene* sc, const pnr_points* q, const char* who);

constexpr int kNCol = 64;
constexpr int kMTiles = kHidden / 128;       // 4
constexpr int kKBlocksH = kHidden / kBlockK; // 8
constexpr int kStages = 5;                   // weight ring slots (16 KiB each)
constexpr int kProducers = 4;                // producer warps: 0, 12, 13, 14
constexpr int kThreads = 512;
constexpr int kGatherWarps = 6;
constexpr int kOperandKB = kNCol * kRowBytes;   // bytes of one 64-row k-block = 8 KiB
constexpr int kChunkBytes = 2 * kOperandKB;     // a 128-feature K-chunk of an activation operand = 16 KiB
constexpr int kTmemCols = 512;
constexpr int kHCol = 256;                   // first h accumulator column
constexpr uint32_t kPackMagic = 0x504e5232u; // "PNR2"
constexpr size_t kPackHeader = 1024;

// ---- weight-stream schedule (shared by the packer and the MMA issuer) --------------------------------
struct Sched {
  int n_blocks;   // ResnetFC blocks
  int CL;         // block index before which the view mean happens (combine_layer), < n_blocks
  int n_linz;     // min(combine_layer, n_blocks)
  int KBz;        // d_latent / 64
};
enum { MAT_LIN_IN = 0, MAT_LINZ = 1, MAT_FC0 = 2, MAT_FC1 = 3, MAT_LIN_OUT = 4 };
// lin_z consumes the latent operand in passes of <= 8 k-blocks (512 channels = the shared-memory latent tile):
// one pass when d_latent <= 512 (the tile stays resident for lin_z[0..2]), several otherwise (re-gathered per pass).
struct ZPos { int pass, mt, kbi, kp; };   // pass index, feature tile, k-block inside the pass, k-blocks in this pass
__host__ __device__ inline int z_passes(const Sched& s);
__host__ __device__ inline ZPos z_position(int KBz, int t) {
  ZPos z;
  z.pass = t / (kMTiles * 8);
  const int tt = t - z.pass * kMTiles * 8;
  z.kp = KBz - z.pass * 8 < 8 ? KBz - z.pass * 8 : 8;
  z.mt = tt / z.kp;
  z.kbi = tt - z.mt * z.kp;
  return z;
}
struct StageSrc { int mat, blk, row0, k0; };

__host__ __device__ inline int z_passes(const Sched& s) { return (s.KBz + 7) / 8; }
__host__ __device__ inline int sched_total(const Sched& s) {
  return kMTiles + s.n_linz * kMTiles * s.KBz + s.n_blocks * 64 + kKBlocksH;
}
// Stage s of the per-tile stream -> segment (which layer) and position inside it.  Order = MMA issue order:
//   lin_in | lin_z[0] | per block b: fc_0 (K-chunk outer) | lin_z[b+1] (covers the h-epilogue latency) | fc_1 | lin_out
struct Seg { int kind, blk, t, len; };   // kind = MAT_*; t = stage index inside the segment; len = segment length
__host__ __device__ inline Seg walk_stage(const Sched& sc, int s) {
  Seg g;
  const int s1 = kMTiles * sc.KBz;
  if (s < kMTiles) { g.kind = MAT_LIN_IN; g.blk = 0; g.t = s; g.len = kMTiles; return g; }
  s -= kMTiles;
  if (s < s1) { g.kind = MAT_LINZ; g.blk = 0; g.t = s; g.len = s1; return g; }
  s -= s1;
  for (int b = 0; b < sc.n_blocks; ++b) {
    const bool has_z = b + 1 < sc.n_linz;
    if (s < 32) { g.kind = MAT_FC0; g.blk = b; g.t = s; g.len = 32; return g; }
    s -= 32;
    if (has_z) {
      if (s < s1) { g.kind = MAT_LINZ; g.blk = b + 1; g.t = s; g.len = s1; return g; }
      s -= s1;
    }
    if (s < 32) { g.kind = MAT_FC1; g.blk = b; g.t = s; g.len = 32; return g; }
    s -= 32;
  }
  g.kind = MAT_LIN_OUT; g.blk = 0; g.t = s; g.len = kKBlocksH;
  return g;
}
// ... and which 128x64 slab of which weight matrix that stage carries (used by the packer).
__host__ __device__ inline StageSrc decode_stage(const Sched& sc, int s) {
  const Seg g = walk_stage(sc, s);
  StageSrc r;
  r.mat = g.kind; r.blk = g.blk;
  switch (g.kind) {
    case MAT_LIN_IN: r.row0 = g.t * 128; r.k0 = 0; break;
    case MAT_LINZ: { const ZPos z = z_position(sc.KBz, g.t); r.row0 = z.mt * 128; r.k0 = (z.pass * 8 + z.kbi) * 64; } break;
    case MAT_FC0:
    case MAT_FC1: r.row0 = ((g.t % 8) / 2) * 128; r.k0 = (g.t / 8) * 128 + (g.t % 2) * 64; break;   // chunk kc=t/8, tile (t%8)/2, half t%2
    default: r.row0 = 0; r.k0 = g.t * 64; break;
  }
  return r;
}

// CTA-pair kernel (mlp_umma_pair.cu)
size_t pair_stream_bytes(const pnr_mlp_params* p, int proj);
int pair_pack(const pnr_mlp_params* p, uint8_t* stream, int proj, cudaStream_t st);
size_t pair_workspace_bytes();
int field_forward_pair(const pnr_scene* sc, const pnr_points* q, const pnr_mlp_params* mp, const uint8_t* stream,
                       const float* bx, const float* bh, const float* bo, float* out, void* ws, size_t ws_bytes,
                       int num_freqs, float freq_factor, int raw, int proj, cudaStream_t st);

struct PackOffsets { size_t stages, bias_x, bias_h, bias_out, pair_stream, total; };
static PackOffsets pack_offsets(const Sched& s, const pnr_mlp_params* p, int proj = 0) {
  PackOffsets o;
  o.stages = kPackHeader;
  o.bias_x = o.stages + (size_t)sched_total(s) * kStageBytes;
  o.bias_h = o.bias_x + (size_t)(s.n_blocks + 1) * kHidden * sizeof(float);
  o.bias_out = o.bias_h + (size_t)s.n_blocks * kHidden * sizeof(float);
  o.pair_stream = (o.bias_out + 128 * sizeof(float) + 1023) & ~(size_t)1023;     // second copy of the weights, CTA-pair order
  o.total = o.pair_stream + pair_stream_bytes(p, proj);
  return o;
}
// PNR_PAIR=0 selects the single-CTA kernel, anything else (default) the CTA-pair kernel.
static bool use_pair_kernel() {
  static int cached = -1;
  if (cached < 0) { const char* e = getenv("PNR_PAIR"); cached = (e && atoi(e) == 0) ? 0 : 1; }
  return cached == 1;
}

__global__ void pack_stages_kernel(pnr_mlp_params mp, Sched sc, uint8_t* __restrict__ stages) {
  const int s = blockIdx.x;
  const StageSrc src = decode_stage(sc, s);
  const float* W; int rows, cols;
  switch (src.mat) {
    case MAT_LIN_IN: W = mp.lin_in_w; rows = mp.d_hidden; cols = mp.d_in; break;
    case MAT_LINZ: W = mp.linz_w[src.blk]; rows = mp.d_hidden; cols = mp.d_latent; break;
    case MAT_FC0: W = mp.fc0_w[src.blk]; rows = mp.d_hidden; cols = mp.d_hidden; break;
    case MAT_FC1: W = mp.fc1_w[src.blk]; rows = mp.d_hidden; cols = mp.d_hidden; break;
    default: W = mp.lin_out_w; rows = mp.d_out; cols = mp.d_hidden; break;
  }
  uint8_t* dst = stages + (size_t)s * kStageBytes;
  for (int i = threadIdx.x; i < kStageRows * kBlockK; i += blockDim.x) {
    const int r = i / kBlockK, k = i % kBlockK;
    const int gr = src.row0 + r, gk = src.k0 + k;
    const float v = (gr < rows && gk < cols) ? W[(size_t)gr * cols + gk] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(dst + swz_offset(r, k)) = __float2bfloat16_rn(v);
  }
}

__global__ void pack_bias_kernel(pnr_mlp_params mp, Sched sc, float* __restrict__ bias_x, float* __restrict__ bias_h,
                                 float* __restrict__ bias_out, uint32_t* __restrict__ header) {
  const int f = threadIdx.x;   // 512 threads
  float cum = mp.lin_in_b[f] + (sc.n_linz > 0 ? mp.linz_b[0][f] : 0.f);
  bias_x[f] = cum;
  for (int e = 1; e <= sc.n_blocks; ++e) {
    if (e - 1 == sc.CL) cum = 0.f;                 // the view-mean epilogue materialises the bias into TMEM
    cum += mp.fc1_b[e - 1][f] + (e < sc.n_linz ? mp.linz_b[e][f] : 0.f);
    bias_x[e * kHidden + f] = cum;
  }
  for (int b = 0; b < sc.n_blocks; ++b) bias_h[b * kHidden + f] = mp.fc0_b[b][f];
  if (f < 128) bias_out[f] = f < mp.d_out ? mp.lin_out_b[f] : 0.f;
  if (f == 0) {
    header[0] = kPackMagic; header[1] = mp.d_in; header[2] = mp.d_latent; header[3] = mp.d_hidden;
    header[4] = mp.d_out; header[5] = mp.n_blocks; header[6] = mp.combine_layer; header[7] = sched_total(sc);
  }
}

// ---- shared-memory map ----------------------------------------------------------------------------
struct Smem {
  static constexpr uint32_t w = 0;                                        // kStages x 16 KiB weight ring
  static constexpr uint32_t ax = w + kStages * kStageBytes;               // relu(x) K-chunk ring, 2 x 16 KiB
  static constexpr uint32_t ah = ax + 2 * kChunkBytes;                    // relu(h) K-chunk ring, 2 x 16 KiB
  static constexpr uint32_t lat = ah + 2 * kChunkBytes;                   // latent operand, 8 k-blocks (C=512)
  static constexpr uint32_t zf = lat + kKBlocksH * kOperandKB;            // z-feature operand, 1 k-block
  static constexpr uint32_t bars = zf + kOperandKB;
  static constexpr uint32_t prof = bars + 256;                            // 32 x 8 B debug counters
  static constexpr uint32_t prog = bars + 512;                            // per-stage MMA program, 8 B x kMaxStages
  static constexpr uint32_t total = prog + 8 * 1024;
};
constexpr int kMaxStages = 1024;
enum {
  B_W_FULL = 0, B_W_EMPTY = B_W_FULL + kStages, B_IN_READY = B_W_EMPTY + kStages, B_IN_FREE, B_X_FULL, B_H_FULL,
  B_AX_READY, B_AX_FREE = B_AX_READY + 2, B_AH_READY = B_AX_FREE + 2, B_AH_FREE = B_AH_READY + 2,
  B_COUNT = B_AH_FREE + 2
};
static_assert(B_COUNT <= 30, "barrier parity bits live in one 32-bit word");

// One entry of the MMA issuer's program: what stage s multiplies, where it accumulates, what it must wait for first
// and which barriers its completion signals.  Built once per CTA in shared memory so that the issue loop -- the
// critical serial path of the kernel -- is one small flat loop instead of a large unrolled schedule.
//   w0: [0,14) B-operand descriptor address field | [14,23) D column | 23 accumulate | 24 post-combine N
//       | [25,30) barrier to wait for + 1        w1: [0,5) commit #1 + 1 | [5,10) commit #2 + 1
struct ProgEntry { uint32_t w0, w1; };
__device__ inline ProgEntry make_prog(const Sched& sc, int s, uint32_t sbase) {
  const Seg g = walk_stage(sc, s);
  uint32_t b_addr = 0, dcol = 0, acc = 1, post = 0, wait_id = 0, c1 = 0, c2 = 0;
  const bool last = g.t == g.len - 1;
  switch (g.kind) {
    case MAT_LIN_IN:
      b_addr = sbase + Smem::zf; dcol = g.t * kNCol; acc = 0;
      if (g.t == 0) wait_id = B_IN_READY + 1;
      break;
    case MAT_LINZ: {
      const ZPos z = z_position(sc.KBz, g.t);
      const bool streaming = z_passes(sc) > 1;
      const bool pass_first = z.mt == 0 && z.kbi == 0, pass_last = z.mt == kMTiles - 1 && z.kbi == z.kp - 1;
      b_addr = sbase + Smem::lat + z.kbi * kOperandKB; dcol = z.mt * kNCol;
      // resident latent tile: gathered once per tile (lin_in waits for it, the last lin_z frees it);
      // streamed latent: every (lin_z, pass) waits for its own gather and frees the tile afterwards
      if (streaming && pass_first && !(g.blk == 0 && z.pass == 0)) wait_id = B_IN_READY + 1;
      if (pass_last && (streaming || g.blk == sc.n_linz - 1)) c1 = B_IN_FREE + 1;
      if (last && g.blk == 0) c2 = B_X_FULL + 1;
    } break;
    case MAT_FC0: {
      const int kc = g.t / 8, j = (g.t % 8) / 2, kk = g.t % 2;
      b_addr = sbase + Smem::ax + ((kc & 1) * 2 + kk) * kOperandKB; dcol = kHCol + j * kNCol; acc = (kc > 0 || kk > 0);
      post = g.blk >= sc.CL;
      if (g.t % 8 == 0) wait_id = B_AX_READY + (kc & 1) + 1;
      if (g.t % 8 == 7) c1 = B_AX_FREE + (kc & 1) + 1;
      if (last) c2 = B_H_FULL + 1;
    } break;
    case MAT_FC1: {
      const int kc = g.t / 8, mt = (g.t % 8) / 2, kk = g.t % 2;
      b_addr = sbase + Smem::ah + ((kc & 1) * 2 + kk) * kOperandKB; dcol = mt * kNCol; post = g.blk >= sc.CL;
      if (g.t % 8 == 0) wait_id = B_AH_READY + (kc & 1) + 1;
      if (g.t % 8 == 7) c1 = B_AH_FREE + (kc & 1) + 1;
      if (last) c2 = B_X_FULL + 1;
    } break;
    default: {
      const int kc = g.t / 2, kk = g.t % 2;
      b_addr = sbase + Smem::ax + ((kc & 1) * 2 + kk) * kOperandKB; dcol = kHCol; acc = g.t > 0; post = 1;
      if (kk == 0) wait_id = B_AX_READY + (kc & 1) + 1;
      if (kk == 1) c1 = B_AX_FREE + (kc & 1) + 1;
      if (last) c2 = B_H_FULL + 1;
    } break;
  }
  ProgEntry e;
  e.w0 = ((b_addr >> 4) & 0x3FFFu) | (dcol << 14) | (acc << 23) | (post << 24) | (wait_id << 25);
  e.w1 = c1 | (c2 << 5);
  return e;
}
static_assert(Smem::total <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void store_bf16(uint8_t* base, uint32_t off, float v) {
  *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16_rn(v);
}

// Optional per-role wait-time counters (cycles), filled when the host sets PNR_PROF=1 (debug aid, off by default):
// [0] MMA total, [1] MMA wait weights, [2] wait gather, [4] wait relu(x) chunk, [5] wait relu(h) chunk,
// [8] epilogue total, [9] wait x_full, [10] wait h_full, [11] wait ax_free, [12] wait ah_free,
// [16] gather total, [17] wait in_free, [20] producer0 total, [21] producer0 wait empty.
__device__ long long* g_prof = nullptr;
#define PROF_T0() const long long t0__ = prof ? clock64() : 0
#define PROF_ADD(slot) do { if (prof && lane == 0) prof[slot] += clock64() - t0__; } while (0)

template <int NS>
__global__ void __launch_bounds__(kThreads, 1)
field_umma_kernel(const pnr_scene sc, const pnr_points q, const uint8_t* __restrict__ stages,
                  const float* __restrict__ bias_x, const float* __restrict__ bias_h,
                  const float* __restrict__ bias_out, float* __restrict__ out, const Sched sch, const int num_freqs,
                  const float freq_factor, const int tiles_per_obj, const int n_tiles, const int d_out,
                  const int raw_out) {
  // Cluster of CS CTAs: every CTA works on its own tile, all CTAs consume the SAME weight stream in lockstep, and
  // each weight stage is fetched from L2 once per cluster (every CTA loads 1/CS of it and multicasts it to all).
  const uint32_t CS = cluster_nctarank();
  const uint32_t crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
  const int n_groups = (n_tiles + (int)CS - 1) / (int)CS;          // tile groups, one tile per CTA of a cluster
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;
  constexpr int PP = kNCol / NS;                       // points per tile
  constexpr int NPOST = ((PP + 15) / 16) * 16;         // UMMA N after the view mean
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  volatile uint32_t* tmem_base_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::bars + 8 * B_COUNT);
  long long* const prof = g_prof ? reinterpret_cast<long long*>(smem + Smem::prof) : nullptr;   // smem counters
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  auto bar = [&](int i) -> uint32_t { return sbase + Smem::bars + 8u * i; };
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("pnr: dynamic smem not 1024-aligned\n"); __trap(); }

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar(B_W_FULL + i), 1); mbar_init(bar(B_W_EMPTY + i), CS); }
    mbar_init(bar(B_IN_READY), kGatherWarps);
    mbar_init(bar(B_IN_FREE), 1);
    mbar_init(bar(B_X_FULL), 1);
    mbar_init(bar(B_H_FULL), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_AX_READY + i), 4); mbar_init(bar(B_AX_FREE + i), 1);
      mbar_init(bar(B_AH_READY + i), 4); mbar_init(bar(B_AH_FREE + i), 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 32 && prof) prof[threadIdx.x] = 0;
  const int n_stages = sched_total(sch);
  for (int i = threadIdx.x; i < n_stages; i += kThreads) {
    ProgEntry pe = make_prog(sch, i, sbase);
    // run length: stages from i on that need no further operand wait (so they can be issued back to back)
    int run = 1;
    while (run < 8 && i + run < n_stages && (make_prog(sch, i + run, sbase).w0 >> 25) == 0) ++run;
    pe.w1 |= (uint32_t)run << 10;
    reinterpret_cast<ProgEntry*>(smem + Smem::prog)[i] = pe;
  }
  if (warp == 1) tmem_alloc(sbase + Smem::bars + 8 * B_COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();        // every CTA's barriers are initialised before any remote signal can arrive
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
  const uint32_t idesc_pre = instr_desc_bf16(kNCol), idesc_post = instr_desc_bf16(NPOST);

  if (warp == 0 || (warp >= 12 && warp < 11 + kProducers)) {
    // ===================== weight producers: warp p streams stages p, p+kProducers, ... ===================
    const int pid = warp == 0 ? 0 : warp - 11;
    const long long t_role0 = prof ? clock64() : 0;
    uint32_t slot = pid, par = 1;                       // "empty"-type wait: first pass through the ring is free
    const uint32_t part = kStageBytes / CS, off = crank * part;
    int carry = pid;                                    // first stage of this producer in the current tile
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      int st = carry;
      for (; st < n_stages; st += kProducers) {
        {
          PROF_T0();
          mbar_wait(bar(B_W_EMPTY + slot), par);
          if (pid == 0) PROF_ADD(21);
        }
        if (elect_one()) {
          const uint8_t* src = stages + (size_t)st * kStageBytes + off;
          const uint32_t dst = sbase + Smem::w + slot * kStageBytes + off;
          mbar_arrive_expect_tx(bar(B_W_FULL + slot), kStageBytes);
          if (CS == 1) bulk_g2s(dst, src, kStageBytes, bar(B_W_FULL + slot));
          else bulk_g2s_multicast(dst, src, part, bar(B_W_FULL + slot), cmask);
        }
        __syncwarp();
        slot += kProducers;
        if (slot >= kStages) { slot -= kStages; par ^= 1; }
      }
      carry = st - n_stages;                            // the ring position continues across tiles
    }
    if (prof && pid == 0 && lane == 0) prof[20] += clock64() - t_role0;
  } else if (warp == 1) {
    // ===================== MMA issuer: one flat loop over the per-tile stage program ========================
    // The whole warp walks the (warp-uniform) program, one elected lane issues.  Operands are computed in uniform
    // control flow from uniform values so that ptxas keeps the tcgen05 operands in uniform registers.
    uint32_t slot = 0, wpar = 0;  // weight ring position / parity
    uint32_t ready = 0;           // consecutive ring slots, starting at `slot`, already known to have landed
    uint32_t ph = 0;              // parity bits of the operand barriers this warp waits on (bit = barrier id)
    const long long t_role0 = prof ? clock64() : 0;
    const uint64_t wdesc0 = smem_desc(sbase + Smem::w);
    const uint64_t bdesc_hi = smem_desc(0);                                        // descriptor with a zero address field
    constexpr uint64_t kStageStep = kStageBytes >> 4;                              // descriptor address units (16 B)
    const uint2* prog = reinterpret_cast<const uint2*>(smem + Smem::prog);
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      for (int st = 0; st < n_stages;) {
        const uint2 cur = prog[st];
        const uint32_t wait_id = cur.x >> 25;
        if (wait_id) {                                                             // operand chunk / tile inputs ready?
          PROF_T0();
          const uint32_t id = wait_id - 1;
          mbar_wait(bar(id), (ph >> id) & 1u);
          ph ^= (1u << id);
          tc_fence_after();
          PROF_ADD(id == B_IN_READY ? 2 : (id >= B_AX_READY && id < B_AX_READY + 2) ? 4 : 5);
        }
        if (ready == 0) {
          // One probe round for the whole ring: lane i tests slot+i; a ballot tells how many stages have landed.
          PROF_T0();
          uint32_t my = slot + lane, mypar = wpar;
          if (my >= kStages) { my -= kStages; mypar ^= 1; }
          const bool ok = lane < kStages ? mbar_test_wait(bar(B_W_FULL + my), mypar) : false;
          const uint32_t m = __ballot_sync(0xffffffffu, ok);
          ready = __ffs(~m) - 1;
          if (ready == 0) { mbar_wait(bar(B_W_FULL + slot), wpar); ready = 1; }
          tc_fence_after();
          PROF_ADD(1);
        }
        // Issue every landed stage of this run back to back: tcgen05.mma issue blocks while the tensor core is busy,
        // so all per-stage bookkeeping outside this branch would add serially to the MMA time.
        const uint32_t run = (cur.y >> 10) & 15u;
        const uint32_t batch = ready < run ? ready : run;
        const long long t_issue0 = prof ? clock64() : 0;
        if (elect_one()) {
          uint2 en = cur;
          uint32_t sl = slot;
          for (uint32_t r = 0; r < batch; ++r) {
            const uint2 ecur = en;
            if (r + 1 < batch) en = prog[st + r + 1];
            const uint64_t b_desc = bdesc_hi | (uint64_t)(ecur.x & 0x3FFFu);
            const uint32_t d_col = (ecur.x >> 14) & 0x1FFu;
            const uint32_t idesc = (ecur.x & (1u << 24)) ? idesc_post : idesc_pre;
            mma_kblock_desc(tmem_base + d_col, wdesc0 + sl * kStageStep, b_desc, idesc, (ecur.x >> 23) & 1u);
            if (CS == 1) mma_commit(bar(B_W_EMPTY + sl));
            else mma_commit_multicast(bar(B_W_EMPTY + sl), cmask);     // frees the slot in every CTA of the cluster
            const uint32_t c1 = ecur.y & 31u, c2 = (ecur.y >> 5) & 31u;
            if (c1) mma_commit(bar(c1 - 1));
            if (c2) mma_commit(bar(c2 - 1));
            sl = sl + 1 == kStages ? 0 : sl + 1;
          }
        }
        __syncwarp();
        if (prof && lane == 0) prof[6] += clock64() - t_issue0;
        ready -= batch;
        st += batch;
        slot += batch;
        if (slot >= kStages) { slot -= kStages; wpar ^= 1; }
      }
    }
    if (prof && lane == 0) prof[0] += clock64() - t_role0;
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue warps ==========================================================
    const int qd = warp - 4;                       // TMEM lane quadrant this warp may access
    const int fl = qd * 32 + lane;                 // feature row inside a 128-row tile
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int kbl = fl >> 6, kk = fl & 63;
    uint32_t ph = 0;
    ph |= (1u << (B_AX_FREE + 0)) | (1u << (B_AX_FREE + 1)) | (1u << (B_AH_FREE + 0)) | (1u << (B_AH_FREE + 1));
    const bool prof_warp = qd == 0;
    auto wait = [&](int id) {
      PROF_T0();
      mbar_wait(bar(id), (ph >> id) & 1u); ph ^= (1u << id); tc_fence_after();
      if (prof_warp) PROF_ADD(id == B_X_FULL ? 9 : id == B_H_FULL ? 10 : (id >= B_AX_FREE && id < B_AX_FREE + 2) ? 11 : 12);
    };
    const long long t_role0 = prof ? clock64() : 0;
    auto arrive = [&](int id) { __syncwarp(); if (lane == 0) mbar_arrive(bar(id)); };
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      const int tile = grp * (int)CS + (int)crank;   // tile >= n_tiles: padding tile, computes on zeros, writes nothing
      const bool live = tile < n_tiles;
      const int obj = tile / tiles_per_obj;
      const int p0 = (tile - obj * tiles_per_obj) * PP;
      for (int e = 0; e <= sch.n_blocks; ++e) {
        // ---- x epilogue e: relu(x + cumulative bias) -> bf16 K-chunks (and the view mean at e == CL)
        wait(B_X_FULL);
        for (int mt = 0; mt < kMTiles; ++mt) {
          const float bias = bias_x[e * kHidden + mt * 128 + fl];
          uint8_t* dst = smem + Smem::ax + ((mt & 1) * 2 + kbl) * kOperandKB;
          if (e < sch.CL) {
            uint32_t r[kNCol];
            tmem_ld<kNCol>(tlane + mt * kNCol, r);
            wait(B_AX_FREE + (mt & 1));
#pragma unroll
            for (int c = 0; c < kNCol; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
          } else if (e == sch.CL) {
            uint32_t r[kNCol];
            tmem_ld<kNCol>(tlane + mt * kNCol, r);
            wait(B_AX_FREE + (mt & 1));
            uint32_t m[NPOST];
#pragma unroll
            for (int p = 0; p < NPOST; ++p) {
              float acc = 0.f;
              if (p < PP) {
#pragma unroll
                for (int v = 0; v < NS; ++v) acc += __uint_as_float(r[v * PP + p]);   // combine_interleaved mean
                acc = __fdiv_rn(acc, (float)NS) + bias;
              }
              m[p] = __float_as_uint(acc);
              store_bf16(dst, swz_offset(p, kk), fmaxf(acc, 0.f));
            }
            tmem_st<NPOST>(tlane + mt * kNCol, m);      // x-bar (bias included) is the new residual stream
          } else {
            uint32_t r[NPOST];
            tmem_ld<NPOST>(tlane + mt * kNCol, r);
            wait(B_AX_FREE + (mt & 1));
#pragma unroll
            for (int c = 0; c < NPOST; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
          }
          tc_fence_before();
          fence_proxy_async();
          arrive(B_AX_READY + (mt & 1));
        }
        if (e < sch.n_blocks) {
          // ---- h epilogue of block e: relu(fc_0 out + b) -> bf16 K-chunks for fc_1
          wait(B_H_FULL);
          for (int j = 0; j < kMTiles; ++j) {
            const float bias = bias_h[e * kHidden + j * 128 + fl];
            uint8_t* dst = smem + Smem::ah + ((j & 1) * 2 + kbl) * kOperandKB;
            if (e < sch.CL) {
              uint32_t r[kNCol];
              tmem_ld<kNCol>(tlane + kHCol + j * kNCol, r);
              wait(B_AH_FREE + (j & 1));
#pragma unroll
              for (int c = 0; c < kNCol; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
            } else {
              uint32_t r[NPOST];
              tmem_ld<NPOST>(tlane + kHCol + j * kNCol, r);
              wait(B_AH_FREE + (j & 1));
#pragma unroll
              for (int c = 0; c < NPOST; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
            }
            tc_fence_before();
            fence_proxy_async();
            arrive(B_AH_READY + (j & 1));
          }
        }
      }
      // ---- output epilogue: lin_out rows live in TMEM lanes 0..d_out-1 of h tile 0
      wait(B_H_FULL);
      {
        uint32_t r[NPOST];
        tmem_ld<NPOST>(tlane + kHCol, r);
        tc_fence_before();
        if (live && qd == 0 && lane < d_out) {
          const float bias = bias_out[lane];
#pragma unroll
          for (int p = 0; p < PP; ++p) {
            if (p0 + p < q.P) {
              float v = __uint_as_float(r[p]) + bias;
              if (!raw_out) v = lane < 3 ? 1.0f / (1.0f + expf(-v)) : fmaxf(v, 0.f);   // models.py:312-317
              out[((size_t)obj * q.P + p0 + p) * d_out + lane] = v;
            }
          }
        }
      }
    }
    if (prof && prof_warp && lane == 0) prof[8] += clock64() - t_role0;
  } else if (warp == 2 || warp == 3 || (warp >= 8 && warp < 12)) {
    // ===================== gather warps: project + 4-tap bilinear gather + positional encoding =======
    const int gw = warp < 4 ? warp - 2 : warp - 6;   // 0..5
    uint32_t par_free = 1;                           // "empty"-type barrier: first wait passes
    const long long t_role0 = prof ? clock64() : 0;
    const int n_pass = z_passes(sch);
    const int fills = n_pass > 1 ? sch.n_linz * n_pass : 1;      // latent tile fills per tile (1 = resident)
    for (int grp = cluster_id; grp < n_groups; grp += n_clusters) {
      const int tile = grp * (int)CS + (int)crank;
      const int obj = tile / tiles_per_obj;
      const int p0 = (tile - obj * tiles_per_obj) * PP;
      for (int fill = 0; fill < fills; ++fill) {
        const int pass = fill % n_pass;
        const int kp = sch.KBz - pass * 8 < 8 ? sch.KBz - pass * 8 : 8;     // k-blocks (64 channels each) in this pass
        const int ch0 = pass * 512;
        {
          PROF_T0();
          mbar_wait(bar(B_IN_FREE), par_free);
          if (gw == 0) PROF_ADD(17);
        }
        par_free ^= 1;
        for (int c = gw; c < kNCol; c += kGatherWarps) {
          const int v = c / PP, p = c - v * PP;
          const bool valid = (tile < n_tiles) && (v < NS) && (p0 + p < q.P);
          Projection pr;
          Taps tp;
          const int view = obj * NS + (v < NS ? v : 0);
          if (valid) {
            float px, py, pz, vx, vy, vz;
            fetch_point(q, (long long)obj * q.P + p0 + p, px, py, pz, vx, vy, vz);
            pr = project_point(sc, view, px, py, pz, vx, vy, vz);
            tp = make_taps(pr.ix, pr.iy, sc.Hl, sc.Wl, sc.C);
          }
          // z-feature row (first fill of the tile only): 64 bf16 (d_in <= 64 used), two per lane
          if (fill == 0) {
            const int j0 = lane * 2;
            float a = 0.f, b = 0.f;
            const int d_in = 6 * num_freqs + 6;
            if (valid) {
              if (j0 < d_in) a = zfeat_value(pr, j0, num_freqs, freq_factor);
              if (j0 + 1 < d_in) b = zfeat_value(pr, j0 + 1, num_freqs, freq_factor);
            }
            *reinterpret_cast<__nv_bfloat162*>(smem + Smem::zf + swz_offset(c, j0)) = __floats2bfloat162_rn(a, b);
          }
          // latent row: channels [ch0, ch0 + 64*kp); lane covers 16 of them = two 16-byte chunks of k-block lane/4
          if ((lane >> 2) < kp) {
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            if (valid) {
              const __nv_bfloat16* fmap = (const __nv_bfloat16*)sc.feat + (size_t)view * sc.Hl * sc.Wl * sc.C + ch0 + lane * 16;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (tp.off[k] < 0) continue;
                const uint4* src = reinterpret_cast<const uint4*>(fmap + tp.off[k]);
                const uint4 r0 = __ldg(src), r1 = __ldg(src + 1);
                const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
                const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float2 f0 = __bfloat1622float2(h0[i]), f1 = __bfloat1622float2(h1[i]);
                  acc[2 * i] += tp.w[k] * f0.x; acc[2 * i + 1] += tp.w[k] * f0.y;
                  acc[8 + 2 * i] += tp.w[k] * f1.x; acc[8 + 2 * i + 1] += tp.w[k] * f1.y;
                }
              }
            }
            uint4 o0, o1;
            __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
            __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              q0[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
              q1[i] = __floats2bfloat162_rn(acc[8 + 2 * i], acc[8 + 2 * i + 1]);
            }
            uint8_t* kb_base = smem + Smem::lat + (lane >> 2) * kOperandKB;
            const int k_in = (lane & 3) * 16;
            *reinterpret_cast<uint4*>(kb_base + swz_offset(c, k_in)) = o0;
            *reinterpret_cast<uint4*>(kb_base + swz_offset(c, k_in + 8)) = o1;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_IN_READY));
      }
    }
    if (prof && gw == 0 && lane == 0) prof[16] += clock64() - t_role0;
  }

  tc_fence_before();
  __syncthreads();
  if (g_prof && threadIdx.x < 32) g_prof[(size_t)blockIdx.x * 32 + threadIdx.x] = prof[threadIdx.x];
  if (CS > 1) cluster_sync_all();        // no CTA exits while a peer may still multicast into its smem / barriers
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- building-block self test ------------------------------------------------------------------------
// D(128 x N) = A(128 x K) * B(N x K)^T : A packed on the device into pre-swizzled 16 KiB stages and streamed
// with cp.async.bulk through a 2-deep ring, B written by the threads into the swizzled operand layout.
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

static int make_sched(const pnr_mlp_params* p, Sched* s, const char* who) {
  PNR_REQUIRE(p, PNR_ERR_ARG, "%s: null params", who);
  PNR_REQUIRE(p->d_hidden == kHidden, PNR_ERR_UNSUPPORTED, "%s: d_hidden=%d (tcgen05 path is built for %d)", who, p->d_hidden, kHidden);
  PNR_REQUIRE(p->d_in > 0 && p->d_in <= 64, PNR_ERR_UNSUPPORTED, "%s: d_in=%d must be in (0,64]", who, p->d_in);
  PNR_REQUIRE(p->d_latent > 0 && p->d_latent % 64 == 0, PNR_ERR_UNSUPPORTED, "%s: d_latent=%d must be a positive multiple of 64", who, p->d_latent);
  PNR_REQUIRE(p->n_blocks >= 1 && p->n_blocks <= 8, PNR_ERR_UNSUPPORTED, "%s: n_blocks=%d", who, p->n_blocks);
  PNR_REQUIRE(p->combine_layer >= 1 && p->combine_layer < p->n_blocks, PNR_ERR_UNSUPPORTED,
              "%s: combine_layer=%d must be in [1, n_blocks) for the fused path", who, p->combine_layer);
  PNR_REQUIRE(p->d_out >= 1 && p->d_out <= 32, PNR_ERR_UNSUPPORTED, "%s: d_out=%d", who, p->d_out);
  s->n_blocks = p->n_blocks; s->CL = p->combine_layer; s->n_linz = p->combine_layer; s->KBz = p->d_latent / 64;
  PNR_REQUIRE(sched_total(*s) <= kMaxStages, PNR_ERR_UNSUPPORTED, "%s: %d weight stages per tile exceed the %d-entry stage program", who, sched_total(*s), kMaxStages);
  return PNR_OK;
}

// Cluster size of the weight multicast: PNR_CLUSTER env (1, 2, 4 or 8) overrides the default of 1.  Measured on
// B200: L2 is not the limiter of this kernel (the per-SM shared-memory fill + operand reads are), so multicast buys
// nothing at the moment and clusters of 4/8 strand SMs (132/120 of 148 usable).
static int cluster_size_setting() {
  static int cached = 0;
  if (cached == 0) {
    int v = 1;
    const char* e = getenv("PNR_CLUSTER");
    if (e) v = atoi(e);
    if (v != 1 && v != 2 && v != 4 && v != 8) v = 1;
    cached = v;
  }
  return cached;
}

size_t field_workspace_umma() { return use_pair_kernel() ? pair_workspace_bytes() : 0; }

int field_forward_umma(const pnr_scene* sc, const pnr_points* q, const pnr_mlp_params* mp, const void* packed,
                       float* out, void* ws, size_t ws_bytes, int num_freqs, float freq_factor, int raw,
                       cudaStream_t st) {
  Sched sch;
  int rc = make_sched(mp, &sch, "field_forward_umma");
  if (rc) return rc;
  PNR_REQUIRE(packed, PNR_ERR_ARG, "field_forward_umma: packed weights missing (call pnr_mlp_pack)");
  PNR_REQUIRE(((uintptr_t)packed & 1023) == 0, PNR_ERR_ARG, "field_forward_umma: packed blob must be 1024-byte aligned");
  PNR_REQUIRE(!sc->feat_fp32, PNR_ERR_ARG, "field_forward_umma: needs bf16 channels-last feature maps");
  const int proj = (sc->flags & PNR_SCENE_PROJECTED) ? 1 : 0;
  if (proj) {
    PNR_REQUIRE(use_pair_kernel(), PNR_ERR_UNSUPPORTED, "field_forward_umma: projected feature maps need the CTA-pair kernel (PNR_PAIR=0 is set)");
    PNR_REQUIRE(sc->C == sch.n_linz * kHidden, PNR_ERR_ARG, "field_forward_umma: projected maps must have n_linz x %d = %d channels (got %d)",
                kHidden, sch.n_linz * kHidden, sc->C);
    sch.KBz = kHidden / 64;              // blob from pnr_mlp_pack_projected: identity lin_z over 512 channels
  } else {
    PNR_REQUIRE(sc->C == mp->d_latent, PNR_ERR_ARG, "field_forward_umma: feature maps have %d channels, the MLP expects %d", sc->C, mp->d_latent);
  }
  PNR_REQUIRE(mp->d_in == 6 * num_freqs + 6, PNR_ERR_ARG, "field_forward_umma: d_in/num_freqs mismatch");
  PNR_REQUIRE(sc->NS >= 1 && sc->NS <= 8 && sc->NS != 7, PNR_ERR_UNSUPPORTED, "field_forward_umma: NS=%d source views", sc->NS);
  if ((long long)sc->SB * q->P == 0) return PNR_OK;
  const int PP = kNCol / sc->NS;
  const int tiles_per_obj = (q->P + PP - 1) / PP;
  const long long n_tiles_ll = (long long)tiles_per_obj * sc->SB;
  PNR_REQUIRE(n_tiles_ll < (1LL << 31), PNR_ERR_ARG, "field_forward_umma: too many tiles");
  const int n_tiles = (int)n_tiles_ll;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int cs = cluster_size_setting();
  while (cs > 1 && n_tiles < cs * 2) cs >>= 1;          // tiny problems: do not pad most of a cluster with dummy tiles
  long long* prof_dev = nullptr;
  if (getenv("PNR_PROF") && !use_pair_kernel()) {
    cudaMalloc(&prof_dev, (size_t)4096 * 32 * sizeof(long long));
    cudaMemset(prof_dev, 0, (size_t)4096 * 32 * sizeof(long long));
    cudaMemcpyToSymbol(g_prof, &prof_dev, sizeof(prof_dev));
  }
  const PackOffsets po = pack_offsets(sch, mp, proj);
  const uint8_t* blob = (const uint8_t*)packed;
  const uint8_t* stages = blob + po.stages;
  const float* bx = (const float*)(blob + po.bias_x);
  const float* bh = (const float*)(blob + po.bias_h);
  const float* bo = (const float*)(blob + po.bias_out);
  if (use_pair_kernel())
    return field_forward_pair(sc, q, mp, blob + po.pair_stream, bx, bh, bo, out, ws, ws_bytes, num_freqs, freq_factor, raw, proj, st);
#define PNR_LAUNCH_NS(NSV)                                                                                   \
  case NSV: {                                                                                                \
    auto kern = field_umma_kernel<NSV>;                                                                      \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem::total); \
    PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));         \
    cudaLaunchConfig_t cfg = {};                                                                             \
    cudaLaunchAttribute attr[1];                                                                             \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                        \
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;               \
    cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Smem::total; cfg.stream = st;                      \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                      \
    int max_clusters = sms / cs;                                                                             \
    if (cs > 1) {                                                                                            \
      cfg.gridDim = dim3(sms / cs * cs);                                                                     \
      int mc = 0;                                                                                            \
      if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) == cudaSuccess && mc > 0) max_clusters = mc;      \
    }                                                                                                        \
    const int n_groups = (n_tiles + cs - 1) / cs;                                                            \
    const int n_clusters = n_groups < max_clusters ? n_groups : max_clusters;                                \
    cfg.gridDim = dim3(n_clusters * cs);                                                                     \
    e = cudaLaunchKernelEx(&cfg, kern, *sc, *q, stages, bx, bh, bo, out, sch, num_freqs, freq_factor,        \
                           tiles_per_obj, n_tiles, (int)mp->d_out, raw);                                     \
    PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "field_umma_kernel launch: %s", cudaGetErrorString(e));      \
  } break;
  switch (sc->NS) {
    PNR_LAUNCH_NS(1) PNR_LAUNCH_NS(2) PNR_LAUNCH_NS(3) PNR_LAUNCH_NS(4) PNR_LAUNCH_NS(5) PNR_LAUNCH_NS(6) PNR_LAUNCH_NS(8)
    default: PNR_REQUIRE(false, PNR_ERR_UNSUPPORTED, "field_forward_umma: NS=%d", sc->NS);
  }
#undef PNR_LAUNCH_NS
  PNR_CHECK_LAUNCH("field_umma_kernel");
  if (prof_dev) {   // debug: dump the per-role wait breakdown (synchronises; never on in production)
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)4096 * 32);
    cudaMemcpy(h.data(), prof_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[32] = {"mma_total", "mma_wait_weights", "mma_wait_gather", 0, "mma_wait_relu_x", "mma_wait_relu_h", "mma_issue_block", 0,
                             "epi_total", "epi_wait_x_full", "epi_wait_h_full", "epi_wait_ax_free", "epi_wait_ah_free", 0, 0, 0,
                             "gather_total", "gather_wait_in_free", 0, 0, "producer0_total", "producer0_wait_empty"};
    int ctas = 0; double sum[32] = {0};
    for (int b = 0; b < 4096; ++b) { if (h[(size_t)b * 32] == 0) continue; ++ctas; for (int k = 0; k < 32; ++k) sum[k] += (double)h[(size_t)b * 32 + k]; }
    const double tiles_per_cta = (double)n_tiles / (ctas ? ctas : 1);
    fprintf(stderr, "[pnr prof] tiles=%d ctas=%d tiles/cta=%.1f  (cycles per tile, mean over CTAs)\n", n_tiles, ctas, tiles_per_cta);
    for (int k = 0; k < 32; ++k) if (names[k]) fprintf(stderr, "[pnr prof]   %-22s %10.0f\n", names[k], sum[k] / (ctas ? ctas : 1) / tiles_per_cta);
    long long* null_ptr = nullptr;
    cudaMemcpyToSymbol(g_prof, &null_ptr, sizeof(null_ptr));
    cudaFree(prof_dev);
  }
  return PNR_OK;
}

}  // namespace pnr

using namespace pnr;

static size_t mlp_pack_bytes(const pnr_mlp_params* p, int proj) {
  Sched s;
  if (make_sched(p, &s, "pnr_mlp_pack_bytes")) return 0;
  if (proj) s.KBz = kHidden / 64;
  return pack_offsets(s, p, proj).total;
}
extern "C" size_t pnr_mlp_pack_bytes(const pnr_mlp_params* p) { return mlp_pack_bytes(p, 0); }
extern "C" size_t pnr_mlp_pack_projected_bytes(const pnr_mlp_params* p) { return mlp_pack_bytes(p, 1); }

static int mlp_pack(const pnr_mlp_params* p, void* packed, int proj, void* stream);
extern "C" int pnr_mlp_pack(const pnr_mlp_params* p, void* packed, void* stream) { return mlp_pack(p, packed, 0, stream); }
extern "C" int pnr_mlp_pack_projected(const pnr_mlp_params* p, void* packed, void* stream) { return mlp_pack(p, packed, 1, stream); }

static int mlp_pack(const pnr_mlp_params* p, void* packed, int proj, void* stream) {
  reset_launch_count();
  Sched s;
  int rc = make_sched(p, &s, "pnr_mlp_pack");
  if (rc) return rc;
  if (proj) s.KBz = kHidden / 64;
  PNR_REQUIRE(packed && ((uintptr_t)packed & 1023) == 0, PNR_ERR_ARG, "pnr_mlp_pack: packed must be non-null and 1024-byte aligned");
  PNR_REQUIRE(p->lin_in_w && p->lin_in_b && p->lin_out_w && p->lin_out_b, PNR_ERR_ARG, "pnr_mlp_pack: null lin_in/lin_out");
  for (int b = 0; b < p->n_blocks; ++b)
    PNR_REQUIRE(p->fc0_w[b] && p->fc0_b[b] && p->fc1_w[b] && p->fc1_b[b], PNR_ERR_ARG, "pnr_mlp_pack: null block %d", b);
  for (int b = 0; b < s.n_linz; ++b) PNR_REQUIRE(p->linz_w[b] && p->linz_b[b], PNR_ERR_ARG, "pnr_mlp_pack: null lin_z %d", b);
  const PackOffsets po = pack_offsets(s, p, proj);
  uint8_t* blob = (uint8_t*)packed;
  cudaStream_t st = (cudaStream_t)stream;
  rc = pair_pack(p, blob + po.pair_stream, proj, st);
  if (rc) return rc;
  if (!proj) {      // the single-CTA kernel's stream (PNR_PAIR=0) has no projected mode; its slot in the blob stays unused
    pack_stages_kernel<<<sched_total(s), 256, 0, st>>>(*p, s, blob + po.stages);
    PNR_CHECK_LAUNCH("pack_stages_kernel");
  }
  pack_bias_kernel<<<1, kHidden, 0, st>>>(*p, s, (float*)(blob + po.bias_x), (float*)(blob + po.bias_h),
                                          (float*)(blob + po.bias_out), (uint32_t*)blob);
  PNR_CHECK_LAUNCH("pack_bias_kernel");
  return PNR_OK;
}

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
