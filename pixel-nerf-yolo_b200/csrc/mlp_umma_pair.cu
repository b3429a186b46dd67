// F2+F3 fused, CTA-pair version: PixelNeRFNet.forward (src/model/models.py:153-318) on tcgen05.mma.cta_group::2.
//
// Same feature-major formulation as mlp_umma.cu (D^T[feat, col] (+)= W[feat, K] * act^T[K, col]) but one MMA now spans
// the two SMs of a cluster of 2:  M = 256 features (128 per CTA), N = 128 columns (64 per CTA: each CTA gathers and
// owns its own tile of 64 columns = PP points x NS views), K = 16.
//   * per MMA each CTA reads 4 KiB of weights + 2 KiB of activations from its own shared memory for 4x the MACs of
//     the single-CTA N=64 MMA -- the shared-memory traffic per MAC that bounds mlp_umma.cu drops by 4, and N=128
//     runs the tensor pipe at its full rate (scripts/umma_bench.py);
//   * each CTA streams only ITS half (128 rows) of every 256-row weight slab (tensor-map TMA, cta_group::2, completion
//     on the leader's mbarrier), so the weight fill per SM per column halves as well;
//   * TMEM per CTA: x^T = 2 feature tiles x 128 columns (cols [0,256)), h^T likewise (cols [256,512)).
// The price: D rows are FEATURES (CTA c holds features 256*mt + 128*c + lane of all 128 columns) while the B operand
// needs COLUMNS (CTA c holds the rows of its own 64 columns for all K).  So every epilogue unit sends half of its
// bf16 output to the peer CTA's shared memory: values are transposed 8x8 across lanes (shuffles) into 16-byte chunks
// = 8 consecutive K elements of one operand row, and stored with st.shared / st.shared::cluster.  Epilogue warp
// (quadrant qd, half hs) handles the 64 columns of CTA hs, so a warp's stores all go to one CTA.
//
// Hand-offs (L = leader CTA 0; "mc" = tcgen05.commit.cta_group::2 multicast to both CTAs):
//   W_FULL[s]  @L   : leader producer's expect_tx(32 KiB) + TMA bytes of both CTAs     W_EMPTY[s] @both : mc
//   IN_READY   @L   : gather warps of both CTAs (remote arrive)                        IN_FREE    @both : mc
//   AX/AH_READY[i] @L : the 8 epilogue warps of CTA i (ring slot i is always written by CTA i)   *_FREE[i] @both : mc
//   X_FULL, H_FULL @both : mc
#include "pnr_common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <vector>

namespace pnr {
using namespace umma;

namespace pair {

constexpr int kNCol = 64;                    // columns per CTA tile
constexpr int kMT = 2;                       // 256-feature tiles per 512-wide layer
constexpr int kKBlocksH = kHidden / kBlockK; // 8
constexpr int kStages = 5;
constexpr int kThreads = 512;
constexpr int kProducers = 3;                // warps 0,12,13
constexpr int kGatherWarps = 4;              // warps 2,3,14,15
constexpr int kEpiWarps = 8;                 // warps 4..11
constexpr int kOperandKB = kNCol * kRowBytes;   // 8 KiB
constexpr int kChunkBytes = 2 * kOperandKB;     // 16 KiB: a 128-feature K-chunk of one CTA's 64 operand rows
constexpr int kTmemCols = 512;
constexpr int kHCol = 256;
constexpr int kMaxStages = 1024;

struct Sched { int n_blocks, CL, n_linz, KBz; };
enum { MAT_LIN_IN = 0, MAT_LINZ = 1, MAT_FC0 = 2, MAT_FC1 = 3, MAT_LIN_OUT = 4 };
struct Seg { int kind, blk, t, len; };
struct StageSrc { int mat, blk, row0, k0; };   // row0 = first row of the 256-row slab

__host__ __device__ inline int z_passes(const Sched& s) { return (s.KBz + 7) / 8; }
__host__ __device__ inline int sched_total(const Sched& s) {
  return kMT + s.n_linz * kMT * s.KBz + s.n_blocks * 32 + kKBlocksH;
}
// per-tile stage order = MMA issue order:
//   lin_in (2) | lin_z[0] | per block: fc_0 (K-chunk outer: kc(4) x mt(2) x kk(2)) | lin_z[b+1] | fc_1 (same) | lin_out (8)
__host__ __device__ inline Seg walk_stage(const Sched& sc, int s) {
  Seg g;
  const int s1 = kMT * sc.KBz;
  if (s < kMT) { g.kind = MAT_LIN_IN; g.blk = 0; g.t = s; g.len = kMT; return g; }
  s -= kMT;
  if (s < s1) { g.kind = MAT_LINZ; g.blk = 0; g.t = s; g.len = s1; return g; }
  s -= s1;
  for (int b = 0; b < sc.n_blocks; ++b) {
    const bool has_z = b + 1 < sc.n_linz;
    if (s < 16) { g.kind = MAT_FC0; g.blk = b; g.t = s; g.len = 16; return g; }
    s -= 16;
    if (has_z) {
      if (s < s1) { g.kind = MAT_LINZ; g.blk = b + 1; g.t = s; g.len = s1; return g; }
      s -= s1;
    }
    if (s < 16) { g.kind = MAT_FC1; g.blk = b; g.t = s; g.len = 16; return g; }
    s -= 16;
  }
  g.kind = MAT_LIN_OUT; g.blk = 0; g.t = s; g.len = kKBlocksH;
  return g;
}
struct ZPos { int pass, mt, kbi, kp; };
__host__ __device__ inline ZPos z_position(int KBz, int t) {
  ZPos z;
  z.pass = t / (kMT * 8);
  const int tt = t - z.pass * kMT * 8;
  z.kp = KBz - z.pass * 8 < 8 ? KBz - z.pass * 8 : 8;
  z.mt = tt / z.kp;
  z.kbi = tt - z.mt * z.kp;
  return z;
}
__host__ __device__ inline StageSrc decode_stage(const Sched& sc, int s) {
  const Seg g = walk_stage(sc, s);
  StageSrc r;
  r.mat = g.kind; r.blk = g.blk;
  switch (g.kind) {
    case MAT_LIN_IN: r.row0 = g.t * 256; r.k0 = 0; break;
    case MAT_LINZ: { const ZPos z = z_position(sc.KBz, g.t); r.row0 = z.mt * 256; r.k0 = (z.pass * 8 + z.kbi) * 64; } break;
    case MAT_FC0:
    case MAT_FC1: r.row0 = ((g.t % 4) / 2) * 256; r.k0 = (g.t / 4) * 128 + (g.t % 2) * 64; break;   // chunk kc=t/4, tile (t%4)/2, half t%2
    default: r.row0 = 0; r.k0 = g.t * 64; break;
  }
  return r;
}

// Packed pair stream: stage s = [CTA0 half: 128 rows x 64 k][CTA1 half], both pre-swizzled 16 KiB k-blocks.
__global__ void pack_stages_kernel(pnr_mlp_params mp, Sched sc, uint8_t* __restrict__ stages) {
  const int s = blockIdx.x >> 1, half = blockIdx.x & 1;
  const StageSrc src = decode_stage(sc, s);
  const float* W; int rows, cols;
  switch (src.mat) {
    case MAT_LIN_IN: W = mp.lin_in_w; rows = mp.d_hidden; cols = mp.d_in; break;
    case MAT_LINZ: W = mp.linz_w[src.blk]; rows = mp.d_hidden; cols = mp.d_latent; break;
    case MAT_FC0: W = mp.fc0_w[src.blk]; rows = mp.d_hidden; cols = mp.d_hidden; break;
    case MAT_FC1: W = mp.fc1_w[src.blk]; rows = mp.d_hidden; cols = mp.d_hidden; break;
    default: W = mp.lin_out_w; rows = mp.d_out; cols = mp.d_hidden; break;
  }
  uint8_t* dst = stages + ((size_t)s * 2 + half) * kStageBytes;
  for (int i = threadIdx.x; i < kStageRows * kBlockK; i += blockDim.x) {
    const int r = i / kBlockK, k = i % kBlockK;
    const int gr = src.row0 + half * 128 + r, gk = src.k0 + k;
    const float v = (gr < rows && gk < cols) ? W[(size_t)gr * cols + gk] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(dst + swz_offset(r, k)) = __float2bfloat16_rn(v);
  }
}

struct Smem {
  static constexpr uint32_t w = 0;
  static constexpr uint32_t ring = w + kStages * kStageBytes;             // 4 K-chunk buffers (16 KiB each): chunk kc -> buffer kc
  static constexpr uint32_t lat = ring + 4 * kChunkBytes;
  static constexpr uint32_t zf = lat + kKBlocksH * kOperandKB;
  static constexpr uint32_t bars = zf + kOperandKB;
  static constexpr uint32_t prog = bars + 512;
  static constexpr uint32_t total = prog + 8 * kMaxStages;
};
enum {
  B_W_FULL = 0, B_W_EMPTY = B_W_FULL + kStages, B_IN_READY = B_W_EMPTY + kStages, B_IN_FREE, B_X_FULL, B_H_FULL,
  B_RDY, B_COUNT = B_RDY + 4
};
static_assert(B_COUNT <= 30, "barrier parity bits live in one 32-bit word");
static_assert(Smem::total <= 227 * 1024, "shared memory budget");

struct ProgEntry { uint32_t w0, w1; };
__device__ inline ProgEntry make_prog(const Sched& sc, int s, uint32_t sbase) {
  const Seg g = walk_stage(sc, s);
  uint32_t b_addr = 0, dcol = 0, acc = 1, post = 0, wait_id = 0, c1 = 0, c2 = 0;
  const bool last = g.t == g.len - 1;
  switch (g.kind) {
    case MAT_LIN_IN:
      b_addr = sbase + Smem::zf; dcol = g.t * 128; acc = 0;
      if (g.t == 0) wait_id = B_IN_READY + 1;
      break;
    case MAT_LINZ: {
      const ZPos z = z_position(sc.KBz, g.t);
      const bool streaming = z_passes(sc) > 1;
      const bool pass_first = z.mt == 0 && z.kbi == 0, pass_last = z.mt == kMT - 1 && z.kbi == z.kp - 1;
      b_addr = sbase + Smem::lat + z.kbi * kOperandKB; dcol = z.mt * 128;
      if (streaming && pass_first && !(g.blk == 0 && z.pass == 0)) wait_id = B_IN_READY + 1;
      if (pass_last && (streaming || g.blk == sc.n_linz - 1)) c1 = B_IN_FREE + 1;
      if (last && g.blk == 0) c2 = B_X_FULL + 1;
    } break;
    case MAT_FC0: {
      const int kc = g.t / 4, mt = (g.t % 4) / 2, kk = g.t % 2;
      b_addr = sbase + Smem::ring + (kc * 2 + kk) * kOperandKB; dcol = kHCol + mt * 128; acc = (kc > 0 || kk > 0);
      post = g.blk >= sc.CL;
      if (g.t % 4 == 0) wait_id = B_RDY + kc + 1;
      if (last) c2 = B_H_FULL + 1;
    } break;
    case MAT_FC1: {
      const int kc = g.t / 4, mt = (g.t % 4) / 2, kk = g.t % 2;
      b_addr = sbase + Smem::ring + (kc * 2 + kk) * kOperandKB; dcol = mt * 128; post = g.blk >= sc.CL;
      if (g.t % 4 == 0) wait_id = B_RDY + kc + 1;
      if (last) c2 = B_X_FULL + 1;
    } break;
    default: {
      const int kc = g.t / 2, kk = g.t % 2;
      b_addr = sbase + Smem::ring + (kc * 2 + kk) * kOperandKB; dcol = kHCol; acc = g.t > 0; post = 1;
      if (kk == 0) wait_id = B_RDY + kc + 1;
      if (last) c2 = B_H_FULL + 1;
    } break;
  }
  ProgEntry e;
  e.w0 = ((b_addr >> 4) & 0x3FFFu) | (dcol << 14) | (acc << 23) | (post << 24) | (wait_id << 25);
  e.w1 = c1 | (c2 << 5);
  return e;
}

// Store an epilogue warp's [32 features x NC columns] block as operand rows (row = column, 16-byte chunk = 8 consecutive
// features).  vals[c] = fp32 bits of this lane's feature at column c; bias/ReLU/bf16 rounding happen here.  Per 16
// columns: pack column pairs to bf16x2, one 8x8 transpose of 32-bit registers across the 8 lanes of a lane group, then
// split low/high halves -> every lane owns two complete 16-byte chunks (columns c0+2i and c0+2i+1).
// `base` = address of the K-chunk (two k-blocks) in the destination CTA; kq = qd*32 = this warp's feature offset in it.
// relu + round-to-nearest bf16 of two floats in one instruction: low half = lo, high half = hi
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <int NC, bool REMOTE, bool ADD_BIAS>
__device__ __forceinline__ void store_transposed(uint32_t base, const uint32_t* vals, float bias, int lane, int kq, int row0) {
  const int g = lane >> 3, i = lane & 7;
  const int k = kq + 8 * g;                                     // first of the 8 features this lane stores, within the chunk
  const uint32_t kb_off = (uint32_t)(k >> 6) * kOperandKB;
#pragma unroll
  for (int c0 = 0; c0 < NC; c0 += 16) {
    // pair column j with column j+8: after the transpose lane i owns rows c0+i and c0+8+i, whose (row & 7) differ
    // across the 8 lanes of a group -> the swizzled 16-byte stores of a quarter warp hit 8 different bank groups
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = __uint_as_float(vals[c0 + j]), y = __uint_as_float(vals[c0 + 8 + j]);
      if (ADD_BIAS) { x += bias; y += bias; }
      p[j] = pack_relu_bf16x2(x, y);
    }
    transpose8x8(p, lane);                                      // p[f] = (feature 8g+f at column c0+i, at column c0+8+i)
    const uint32_t e0 = __byte_perm(p[0], p[1], 0x5410), e1 = __byte_perm(p[2], p[3], 0x5410);
    const uint32_t e2 = __byte_perm(p[4], p[5], 0x5410), e3 = __byte_perm(p[6], p[7], 0x5410);
    const uint32_t o0 = __byte_perm(p[0], p[1], 0x7632), o1 = __byte_perm(p[2], p[3], 0x7632);
    const uint32_t o2 = __byte_perm(p[4], p[5], 0x7632), o3 = __byte_perm(p[6], p[7], 0x7632);
    const uint32_t addr_e = base + kb_off + swz_offset(row0 + c0 + i, k & 63);
    const uint32_t addr_o = base + kb_off + swz_offset(row0 + c0 + 8 + i, k & 63);
    if (REMOTE) { st_cluster_v4(addr_e, e0, e1, e2, e3); st_cluster_v4(addr_o, o0, o1, o2, o3); }
    else { st_shared_v4(addr_e, e0, e1, e2, e3); st_shared_v4(addr_o, o0, o1, o2, o3); }
  }
}

// optional per-role wait counters (PNR_PROF=1): [0] MMA total [1] wait weights [2] wait gather [4] wait relu(x) chunk
// [5] wait relu(h) chunk [6] blocked in issue | [8] epilogue total [9] wait x_full [10] wait h_full [11] wait ax_free
// [12] wait ah_free | [16] gather total [17] wait in_free    (leader CTA's MMA warp; warp 4 and warp 2 of every CTA)
__device__ long long* g_prof_pair = nullptr;
#define PPROF_T0() const long long t0__ = prof ? clock64() : 0
#define PPROF_ADD(slot) do { if (prof && lane == 0) prof[slot] += clock64() - t0__; } while (0)

template <int NS>
__global__ void __launch_bounds__(kThreads, 1)
field_pair_kernel(const pnr_scene sc, const pnr_points q, const __grid_constant__ CUtensorMap wmap,
                  const float* __restrict__ bias_x, const float* __restrict__ bias_h,
                  const float* __restrict__ bias_out, float* __restrict__ out, const Sched sch, const int num_freqs,
                  const float freq_factor, const int tiles_per_obj, const int n_tiles, const int d_out,
                  const int raw_out) {
  const uint32_t crank = cluster_ctarank();            // 0 = leader (issues every MMA of the pair)
  const int n_groups = (n_tiles + 1) / 2;              // one tile per CTA of the pair
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  constexpr int PP = kNCol / NS;
  constexpr int NPOST = ((PP + 15) / 16) * 16;         // operand rows per CTA after the view mean
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  volatile uint32_t* tmem_base_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::bars + 8 * B_COUNT);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  long long* const prof = g_prof_pair ? reinterpret_cast<long long*>(smem + Smem::bars + 256) : nullptr;
  if (prof && threadIdx.x < 32) prof[threadIdx.x] = 0;
  auto bar = [&](int i) -> uint32_t { return sbase + Smem::bars + 8u * i; };
  auto lbar = [&](int i) -> uint32_t { return mapa_u32(sbase + Smem::bars + 8u * i, 0); };   // the leader's copy
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("pnr: dynamic smem not 1024-aligned\n"); __trap(); }

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar(B_W_FULL + i), 1); mbar_init(bar(B_W_EMPTY + i), 1); }
    mbar_init(bar(B_IN_READY), 2 * kGatherWarps);
    mbar_init(bar(B_IN_FREE), 1);
    mbar_init(bar(B_X_FULL), 1);
    mbar_init(bar(B_H_FULL), 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar(B_RDY + i), kEpiWarps);
    fence_barrier_init();
  }
  const int n_stages = sched_total(sch);
  for (int i = threadIdx.x; i < n_stages; i += kThreads) {
    ProgEntry pe = make_prog(sch, i, sbase);
    int run = 1;
    while (run < 8 && i + run < n_stages && (make_prog(sch, i + run, sbase).w0 >> 25) == 0) ++run;
    pe.w1 |= (uint32_t)run << 10;
    reinterpret_cast<ProgEntry*>(smem + Smem::prog)[i] = pe;
  }
  if (warp == 1) tmem_alloc_2sm(sbase + Smem::bars + 8 * B_COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // both CTAs' barriers initialised and TMEM allocated before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
  const uint32_t idesc_pre = instr_desc_bf16_2sm(2 * kNCol), idesc_post = instr_desc_bf16_2sm(2 * NPOST);

  if (warp == 0 || warp == 12 || warp == 13) {
    // ===================== weight producers: warp p streams global stages g = p (mod kProducers); each CTA loads its
    // 128-row half of every 256-row slab, both halves complete on the LEADER's W_FULL barrier
    const int pid = warp == 0 ? 0 : warp - 11;
    uint32_t slot = pid, par = 1;
    int carry = pid;                                    // first stage of this producer inside the current tile
    for (int grp = pair_id; grp < n_groups; grp += n_pairs) {
      int st = carry;
      for (; st < n_stages; st += kProducers) {
        mbar_wait_cluster(bar(B_W_EMPTY + slot), par);
        if (elect_one()) {
          if (crank == 0) mbar_arrive_expect_tx(bar(B_W_FULL + slot), 2 * kStageBytes);
          tma_load_2d_2sm(sbase + Smem::w + slot * kStageBytes, &wmap, 0, (st * 2 + (int)crank) * kStageRows, bar(B_W_FULL + slot));
        }
        __syncwarp();
        slot += kProducers;
        if (slot >= kStages) { slot -= kStages; par ^= 1; }
      }
      carry = st - n_stages;                            // the ring position continues across tiles
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ===================== MMA issuer (leader CTA only): flat loop over the per-tile stage program ============
      uint32_t slot = 0, wpar = 0, ready = 0, ph = 0;
      const long long t_role0 = prof ? clock64() : 0;
      const uint64_t wdesc0 = smem_desc(sbase + Smem::w);
      const uint64_t bdesc_hi = smem_desc(0);
      constexpr uint64_t kStageStep = kStageBytes >> 4;
      const uint2* prog = reinterpret_cast<const uint2*>(smem + Smem::prog);
      for (int grp = pair_id; grp < n_groups; grp += n_pairs) {
        for (int st = 0; st < n_stages;) {
          const uint2 cur = prog[st];
          const uint32_t wait_id = cur.x >> 25;
          if (wait_id) {
            PPROF_T0();
            const uint32_t id = wait_id - 1;
            mbar_wait_cluster(bar(id), (ph >> id) & 1u);
            ph ^= (1u << id);
            tc_fence_after();
            PPROF_ADD(id == B_IN_READY ? 2 : 4);
          }
          if (ready == 0) {
            PPROF_T0();
            uint32_t my = slot + lane, mypar = wpar;
            if (my >= kStages) { my -= kStages; mypar ^= 1; }
            const bool ok = lane < kStages ? mbar_test_wait(bar(B_W_FULL + my), mypar) : false;
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            ready = __ffs(~m) - 1;
            if (ready == 0) { mbar_wait(bar(B_W_FULL + slot), wpar); ready = 1; }
            tc_fence_after();
            PPROF_ADD(1);
          }
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
              mma_kblock_desc_2sm(tmem_base + d_col, wdesc0 + sl * kStageStep, b_desc, idesc, (ecur.x >> 23) & 1u);
              mma_commit_2sm(bar(B_W_EMPTY + sl), 3);
              const uint32_t c1 = ecur.y & 31u, c2 = (ecur.y >> 5) & 31u;
              if (c1) mma_commit_2sm(bar(c1 - 1), 3);
              if (c2) mma_commit_2sm(bar(c2 - 1), 3);
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
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue warps ===================================================================
    // 8 warps = 4 TMEM lane quadrants (qd) x 2.  Unit = this CTA's 128 features of feature tile mt x all 128 columns of
    // the pair = K-chunk kc = 2*mt + crank, written to chunk buffer kc of BOTH CTAs (rows = that CTA's own columns).
    //  * regular epilogues: warp (qd, part) converts columns [part*NC/2, (part+1)*NC/2) of BOTH tiles, the peer's first
    //    (remote st.shared::cluster stores are in flight while the local half is computed), one publish per unit;
    //  * view-mean epilogue: warp (qd, hs) needs all 64 columns of tile hs (the three views of a point) in one thread.
    const int qd = warp & 3;                       // TMEM lane quadrant (warp % 4)
    const int hs = (warp - 4) >> 2;                // second index: column part, or tile in the mean / output epilogues
    const int fl = qd * 32 + lane;                 // feature row inside this CTA's 128-row half
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    uint32_t ph = 0;
    const bool prof_warp = warp == 4;
    auto wait = [&](int id) {
      PPROF_T0();
      mbar_wait_cluster(bar(id), (ph >> id) & 1u); ph ^= (1u << id); tc_fence_after();
      if (prof_warp) PPROF_ADD(id == B_X_FULL ? 9 : 10);
    };
    const long long t_role0 = prof ? clock64() : 0;
    const uint32_t peer = crank ^ 1u;
    auto publish = [&](int kc) {                   // operand rows of chunk kc written -> tell the leader's MMA warp
      const long long tp0 = (prof && prof_warp) ? clock64() : 0;
      fence_proxy_async_cluster();                 // waits until this warp's local AND remote rows are performed
      tc_fence_before();
      __syncwarp();
      // the fence already made the rows visible where the tensor cores read them: no cluster-scope release needed
      if (lane == 0) { if (crank == 0) mbar_arrive(bar(B_RDY + kc)); else mbar_arrive_cluster_relaxed(lbar(B_RDY + kc)); }
      if (prof && prof_warp && lane == 0) prof[18] += clock64() - tp0;
    };
    // both halves of a unit: columns [part*NC/2 ...) of tile `peer` (remote rows) then of tile `crank` (local rows)
    auto convert_unit = [&](uint32_t tcol, int nch, int kc, float bias, bool post) {
      const uint32_t loc = sbase + Smem::ring + kc * kChunkBytes;
      const uint32_t rem = mapa_u32(loc, peer);
      if (!post) {
        uint32_t v[kNCol / 2];
        tmem_ld<kNCol / 2>(tlane + tcol + peer * kNCol + hs * (kNCol / 2), v);
        store_transposed<kNCol / 2, true, true>(rem, v, bias, lane, qd * 32, hs * (kNCol / 2));
        tmem_ld<kNCol / 2>(tlane + tcol + crank * kNCol + hs * (kNCol / 2), v);
        store_transposed<kNCol / 2, false, true>(loc, v, bias, lane, qd * 32, hs * (kNCol / 2));
      } else if constexpr (NPOST >= 32) {
        uint32_t v[NPOST / 2];
        tmem_ld<NPOST / 2>(tlane + tcol + peer * NPOST + hs * (NPOST / 2), v);
        store_transposed<NPOST / 2, true, true>(rem, v, bias, lane, qd * 32, hs * (NPOST / 2));
        tmem_ld<NPOST / 2>(tlane + tcol + crank * NPOST + hs * (NPOST / 2), v);
        store_transposed<NPOST / 2, false, true>(loc, v, bias, lane, qd * 32, hs * (NPOST / 2));
      } else if (hs == 0) {      // 16 post-combine columns per tile: too few to split, warp (qd,0) converts both tiles
        uint32_t v[NPOST];
        tmem_ld<NPOST>(tlane + tcol + peer * NPOST, v);
        store_transposed<NPOST, true, true>(rem, v, bias, lane, qd * 32, 0);
        tmem_ld<NPOST>(tlane + tcol + crank * NPOST, v);
        store_transposed<NPOST, false, true>(loc, v, bias, lane, qd * 32, 0);
      }
      (void)nch;
    };
    for (int grp = pair_id; grp < n_groups; grp += n_pairs) {
      const int tile = grp * 2 + hs;               // mean / output epilogues: the tile whose columns this warp handles
      const bool live = tile < n_tiles;
      const int obj = tile / tiles_per_obj;
      const int p0 = (tile - obj * tiles_per_obj) * PP;
      for (int e = 0; e <= sch.n_blocks; ++e) {
        wait(B_X_FULL);
        for (int mt = 0; mt < kMT; ++mt) {
          const int kc = 2 * mt + (int)crank;
          const float bias = bias_x[e * kHidden + mt * 256 + crank * 128 + fl];
          if (e != sch.CL) {
            const long long ts0 = (prof && prof_warp) ? clock64() : 0;
            convert_unit(mt * 128, 0, kc, bias, e > sch.CL);
            if (prof && prof_warp && lane == 0) prof[14] += clock64() - ts0;
          } else {
            // view mean (combine_interleaved): warp (qd, hs) owns tile hs
            const bool remote = (uint32_t)hs != crank;
            uint32_t m[NPOST];
            {
              uint32_t v[kNCol];
              tmem_ld<kNCol>(tlane + mt * 128 + hs * kNCol, v);
              // the post-combine layout packs NPOST columns per CTA: warp (qd,1) writes columns that warp (qd,0) reads
              asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
#pragma unroll
              for (int p = 0; p < NPOST; ++p) {
                float acc = 0.f;
                if (p < PP) {
#pragma unroll
                  for (int vw = 0; vw < NS; ++vw) acc += __uint_as_float(v[vw * PP + p]);
                  acc = __fdiv_rn(acc, (float)NS) + bias;
                }
                m[p] = __float_as_uint(acc);
              }
            }
            tmem_st<NPOST>(tlane + mt * 128 + hs * NPOST, m);      // x-bar (bias included), post-combine column layout
            const uint32_t loc = sbase + Smem::ring + kc * kChunkBytes;
            if (remote) store_transposed<NPOST, true, false>(mapa_u32(loc, (uint32_t)hs), m, 0.f, lane, qd * 32, 0);
            else store_transposed<NPOST, false, false>(loc, m, 0.f, lane, qd * 32, 0);
          }
          publish(kc);
        }
        if (e < sch.n_blocks) {
          wait(B_H_FULL);
          for (int mt = 0; mt < kMT; ++mt) {
            const int kc = 2 * mt + (int)crank;
            const float bias = bias_h[e * kHidden + mt * 256 + crank * 128 + fl];
            convert_unit(kHCol + mt * 128, 0, kc, bias, e >= sch.CL);
            publish(kc);
          }
        }
      }
      // ---- output: lin_out rows are features 0..d_out-1 -> leader CTA, TMEM lanes 0..d_out-1 of h tile 0
      wait(B_H_FULL);
      {
        uint32_t r[NPOST];
        tmem_ld<NPOST>(tlane + kHCol + hs * NPOST, r);
        tc_fence_before();
        if (crank == 0 && live && qd == 0 && lane < d_out) {
          const float bias = bias_out[lane];
#pragma unroll
          for (int p = 0; p < PP; ++p) {
            if (p0 + p < q.P) {
              float val = __uint_as_float(r[p]) + bias;
              if (!raw_out) val = lane < 3 ? 1.0f / (1.0f + expf(-val)) : fmaxf(val, 0.f);   // models.py:312-317
              out[((size_t)obj * q.P + p0 + p) * d_out + lane] = val;
            }
          }
        }
      }
    }
    if (prof && prof_warp && lane == 0) prof[8] += clock64() - t_role0;
  } else if (warp == 2 || warp == 3 || warp == 14 || warp == 15) {
    // ===================== gather warps: this CTA's 64 columns -> its own z-feature / latent operand rows =======
    const int gw = warp < 4 ? warp - 2 : warp - 12;
    uint32_t par_free = 1;
    const long long t_role0 = prof ? clock64() : 0;
    const int n_pass = z_passes(sch);
    const int fills = n_pass > 1 ? sch.n_linz * n_pass : 1;
    for (int grp = pair_id; grp < n_groups; grp += n_pairs) {
      const int tile = grp * 2 + (int)crank;
      const int obj = tile / tiles_per_obj;
      const int p0 = (tile - obj * tiles_per_obj) * PP;
      for (int fill = 0; fill < fills; ++fill) {
        const int pass = fill % n_pass;
        const int kp = sch.KBz - pass * 8 < 8 ? sch.KBz - pass * 8 : 8;
        const int ch0 = pass * 512;
        {
          PPROF_T0();
          mbar_wait_cluster(bar(B_IN_FREE), par_free);
          if (gw == 0) PPROF_ADD(17);
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
        fence_proxy_async();                 // the gather writes this CTA's own shared memory only
        __syncwarp();
        if (lane == 0) { if (crank == 0) mbar_arrive(bar(B_IN_READY)); else mbar_arrive_cluster_relaxed(lbar(B_IN_READY)); }
      }
    }
    if (prof && gw == 0 && lane == 0) prof[16] += clock64() - t_role0;
  }

  tc_fence_before();
  __syncthreads();
  if (g_prof_pair && threadIdx.x < 32) g_prof_pair[(size_t)blockIdx.x * 32 + threadIdx.x] = prof[threadIdx.x];
  cluster_sync_all();                    // no CTA exits (or frees TMEM) while the pair may still touch it
  if (warp == 1) tmem_dealloc_2sm(tmem_base, kTmemCols);
}

}  // namespace pair

// ---- host side ---------------------------------------------------------------------------------------------
size_t pair_stream_bytes(const pnr_mlp_params* p) {
  pair::Sched s{p->n_blocks, p->combine_layer, p->combine_layer, p->d_latent / 64};
  return (size_t)pair::sched_total(s) * 2 * kStageBytes;
}
int pair_stages(const pnr_mlp_params* p) {
  pair::Sched s{p->n_blocks, p->combine_layer, p->combine_layer, p->d_latent / 64};
  return pair::sched_total(s);
}
int pair_pack(const pnr_mlp_params* p, uint8_t* stream, cudaStream_t st) {
  pair::Sched s{p->n_blocks, p->combine_layer, p->combine_layer, p->d_latent / 64};
  PNR_REQUIRE(pair::sched_total(s) <= pair::kMaxStages, PNR_ERR_UNSUPPORTED, "pair_pack: %d stages exceed the stage program", pair::sched_total(s));
  pair::pack_stages_kernel<<<pair::sched_total(s) * 2, 256, 0, st>>>(*p, s, stream);
  PNR_CHECK_LAUNCH("pair::pack_stages_kernel");
  return PNR_OK;
}

int field_forward_pair(const pnr_scene* sc, const pnr_points* q, const pnr_mlp_params* mp, const uint8_t* stream,
                       const float* bx, const float* bh, const float* bo, float* out, int num_freqs, float freq_factor,
                       int raw, cudaStream_t st) {
  pair::Sched sch{mp->n_blocks, mp->combine_layer, mp->combine_layer, mp->d_latent / 64};
  const int PP = pair::kNCol / sc->NS;
  const int tiles_per_obj = (q->P + PP - 1) / PP;
  const long long n_tiles_ll = (long long)tiles_per_obj * sc->SB;
  PNR_REQUIRE(n_tiles_ll < (1LL << 31), PNR_ERR_ARG, "field_forward_pair: too many tiles");
  const int n_tiles = (int)n_tiles_ll;
  // tensor map over the packed pair stream: rows = stages x 256, 64 bf16 per row (128 B), box = 64 x 128, no swizzle
  // (the stream is already stored in the swizzled operand layout)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    PNR_REQUIRE(e == cudaSuccess && fn, PNR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    encode = (EncodeFn)fn;
  }
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {64, (cuuint64_t)pair::sched_total(sch) * 256};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)stream, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PNR_REQUIRE(r == CUDA_SUCCESS, PNR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_groups = (n_tiles + 1) / 2;
  long long* prof_dev = nullptr;
  if (getenv("PNR_PROF")) {
    cudaMalloc(&prof_dev, (size_t)4096 * 32 * sizeof(long long));
    cudaMemset(prof_dev, 0, (size_t)4096 * 32 * sizeof(long long));
    cudaMemcpyToSymbol(pair::g_prof_pair, &prof_dev, sizeof(prof_dev));
  }
#define PNR_LAUNCH_PAIR(NSV)                                                                                     \
  case NSV: {                                                                                                    \
    auto kern = pair::field_pair_kernel<NSV>;                                                                    \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair::Smem::total); \
    PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));             \
    cudaLaunchConfig_t cfg = {};                                                                                 \
    cudaLaunchAttribute attr[1];                                                                                 \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                            \
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;                    \
    cfg.blockDim = dim3(pair::kThreads); cfg.dynamicSmemBytes = pair::Smem::total; cfg.stream = st;              \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                          \
    cfg.gridDim = dim3(sms / 2 * 2);                                                                             \
    int max_pairs = sms / 2, mc = 0;                                                                             \
    if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) == cudaSuccess && mc > 0) max_pairs = mc;               \
    const int n_pairs = n_groups < max_pairs ? n_groups : max_pairs;                                             \
    cfg.gridDim = dim3(n_pairs * 2);                                                                             \
    e = cudaLaunchKernelEx(&cfg, kern, *sc, *q, tmap, bx, bh, bo, out, sch, num_freqs, freq_factor, tiles_per_obj, \
                           n_tiles, (int)mp->d_out, raw);                                                        \
    PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "field_pair_kernel launch: %s", cudaGetErrorString(e));          \
  } break;
  switch (sc->NS) {
    PNR_LAUNCH_PAIR(1) PNR_LAUNCH_PAIR(2) PNR_LAUNCH_PAIR(3) PNR_LAUNCH_PAIR(4) PNR_LAUNCH_PAIR(5) PNR_LAUNCH_PAIR(6) PNR_LAUNCH_PAIR(8)
    default: PNR_REQUIRE(false, PNR_ERR_UNSUPPORTED, "field_forward_pair: NS=%d", sc->NS);
  }
#undef PNR_LAUNCH_PAIR
  PNR_CHECK_LAUNCH("field_pair_kernel");
  if (prof_dev) {   // debug only: synchronises and prints the per-role wait breakdown
    cudaStreamSynchronize(st);
    std::vector<long long> h((size_t)4096 * 32);
    cudaMemcpy(h.data(), prof_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[32] = {"mma_total", "mma_wait_weights", "mma_wait_gather", 0, "mma_wait_chunk", 0, "mma_issue_block", 0,
                             "epi_total", "epi_wait_x_full", "epi_wait_h_full", 0, 0, 0, "epi_convert_x(x10 units)", 0,
                             "gather_total", "gather_wait_in_free", "epi_publish(x22)"};
    int ctas = 0, leaders = 0; double sum[32] = {0};
    for (int b = 0; b < 4096; ++b) {
      if (h[(size_t)b * 32 + 8] == 0) continue;
      ++ctas; if (h[(size_t)b * 32] != 0) ++leaders;
      for (int k = 0; k < 32; ++k) sum[k] += (double)h[(size_t)b * 32 + k];
    }
    const double groups_per_pair = (double)n_groups / (leaders ? leaders : 1);
    fprintf(stderr, "[pnr pair prof] tiles=%d pairs=%d tile-pairs/pair=%.1f  (cycles per tile-pair)\n", n_tiles, leaders, groups_per_pair);
    for (int k = 0; k < 32; ++k) if (names[k]) fprintf(stderr, "[pnr pair prof]   %-22s %10.0f\n", names[k], sum[k] / ((k < 8) ? (leaders ? leaders : 1) : (ctas ? ctas : 1)) / groups_per_pair);
    long long* null_ptr = nullptr;
    cudaMemcpyToSymbol(pair::g_prof_pair, &null_ptr, sizeof(null_ptr));
    cudaFree(prof_dev);
  }
  return PNR_OK;
}

}  // namespace pnr

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
