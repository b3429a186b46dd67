// F2+F3 fused: PixelNeRFNet.forward (src/model/models.py:153-318) as ONE persistent sm_100a kernel.
//
// Tile = 64 "columns" = PP points x NS source views (PP = 64/NS).  The MLP is evaluated feature-major:
//     D^T[512 features, 64 columns] (+)= W[512, K] * act^T[K, 64]
// so the nn.Linear weight matrix (out x in, row-major) is the K-major A operand as it is stored, and an
// activation row (one point-view, K contiguous) is the K-major B operand.  tcgen05.mma M=128 (4 feature
// tiles), N=64 (pre-combine) or N=round16(PP) (after the view mean), K=16, bf16 in, fp32 accumulate.
//
//   TMEM   cols [0,256)   x^T   : the fp32 residual stream of the tile, 4 feature tiles x 64 columns.
//                                 `x += lin_z(latent)` and `x += fc_1(relu(h))` are MMAs that accumulate in
//                                 place (resnetfc.py:176-182, 53-62) -- the residual add costs nothing.
//          cols [256,384) h^T   : two 128x64 fc_0 accumulators (double buffer).
//   SMEM   weight ring (3 x 16 KiB)  <- cp.async.bulk (TMA engine) from the packed, pre-swizzled stream
//          relu(x) operand 64 KiB, latent operand 64 KiB (gathered ONCE per tile, reused by lin_z[0..2]),
//          relu(h) chunk x2 32 KiB (fc_0 tile j is consumed by fc_1 as K-chunk j right away), z-feature 8 KiB.
//   Warps  0: weight producer   1: MMA issuer (+TMEM alloc)   4-7: epilogue (TMEM->bias/ReLU/bf16->smem,
//          view mean, final sigmoid/relu)   2,3,8-11: project + 4-tap gather + positional encoding.
//
// Biases never touch the tensor pipe: the TMEM accumulator holds x minus the biases added so far, and the
// epilogue adds the cumulative bias vector (precomputed at pack time) when it forms relu(x).
#include "pnr_common.cuh"
#include "umma.cuh"

namespace pnr {
using namespace umma;

int validate_scene_points(const pnr_scene* sc, const pnr_points* q, const char* who);

constexpr int kNCol = 64;
constexpr int kMTiles = kHidden / 128;       // 4
constexpr int kKBlocksH = kHidden / kBlockK; // 8
constexpr int kStages = 3;
constexpr int kThreads = 384;
constexpr int kGatherWarps = 6;
constexpr int kOperandKB = kNCol * kRowBytes;   // bytes of one 64-row k-block = 8 KiB
constexpr int kTmemCols = 512;
constexpr int kHCol = 256;                   // first h accumulator column
constexpr uint32_t kPackMagic = 0x504e5231u; // "PNR1"
constexpr size_t kPackHeader = 1024;

// ---- weight-stream schedule (shared by the packer and the MMA issuer) --------------------------------
struct Sched {
  int n_blocks;   // ResnetFC blocks
  int CL;         // block index before which the view mean happens (combine_layer), < n_blocks
  int n_linz;     // min(combine_layer, n_blocks)
  int KBz;        // d_latent / 64
};
enum { MAT_LIN_IN = 0, MAT_LINZ = 1, MAT_FC0 = 2, MAT_FC1 = 3, MAT_LIN_OUT = 4 };
struct StageSrc { int mat, blk, row0, k0; };

__host__ __device__ inline int sched_total(const Sched& s) {
  return kMTiles + s.n_linz * kMTiles * s.KBz + s.n_blocks * 64 + kKBlocksH;
}
// Stage s of the per-tile stream -> which 128x64 slab of which matrix.  Order = MMA issue order.
__host__ __device__ inline StageSrc decode_stage(const Sched& sc, int s) {
  StageSrc r;
  if (s < kMTiles) { r.mat = MAT_LIN_IN; r.blk = 0; r.row0 = s * 128; r.k0 = 0; return r; }
  s -= kMTiles;
  const int s1 = kMTiles * sc.KBz;
  if (s < s1) { r.mat = MAT_LINZ; r.blk = 0; r.row0 = (s / sc.KBz) * 128; r.k0 = (s % sc.KBz) * 64; return r; }
  s -= s1;
  for (int b = 0; b < sc.n_blocks; ++b) {
    const int len = 64 + ((b + 1 < sc.n_linz) ? s1 : 0);
    if (s < len) {
      if (s >= 64) { s -= 64; r.mat = MAT_LINZ; r.blk = b + 1; r.row0 = (s / sc.KBz) * 128; r.k0 = (s % sc.KBz) * 64; return r; }
      // interleave: S2(0) | S2(1) S3p(0) | S2(2) S3p(1) | S2(3) S3p(2) | S3p(3)
      int j, t; bool is_s2;
      if (s < 8) { is_s2 = true; j = 0; t = s; }
      else if (s >= 56) { is_s2 = false; j = 3; t = s - 56; }
      else { int u = s - 8; int grp = u / 16; int w = u % 16; is_s2 = w < 8; t = w % 8; j = is_s2 ? grp + 1 : grp; }
      if (is_s2) { r.mat = MAT_FC0; r.blk = b; r.row0 = j * 128; r.k0 = t * 64; }
      else { r.mat = MAT_FC1; r.blk = b; r.row0 = (t / 2) * 128; r.k0 = j * 128 + (t % 2) * 64; }
      return r;
    }
    s -= len;
  }
  r.mat = MAT_LIN_OUT; r.blk = 0; r.row0 = 0; r.k0 = s * 64;
  return r;
}

struct PackOffsets { size_t stages, bias_x, bias_h, bias_out, total; };
static PackOffsets pack_offsets(const Sched& s) {
  PackOffsets o;
  o.stages = kPackHeader;
  o.bias_x = o.stages + (size_t)sched_total(s) * kStageBytes;
  o.bias_h = o.bias_x + (size_t)(s.n_blocks + 1) * kHidden * sizeof(float);
  o.bias_out = o.bias_h + (size_t)s.n_blocks * kHidden * sizeof(float);
  o.total = o.bias_out + 128 * sizeof(float);
  return o;
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
  static constexpr uint32_t w = 0;                                        // kStages x 16 KiB
  static constexpr uint32_t ax = w + kStages * kStageBytes;               // relu(x) operand, 8 k-blocks
  static constexpr uint32_t lat = ax + kKBlocksH * kOperandKB;            // latent operand, 8 k-blocks (C=512)
  static constexpr uint32_t ah = lat + kKBlocksH * kOperandKB;            // relu(h) chunks, 2 x 2 k-blocks
  static constexpr uint32_t zf = ah + 2 * 2 * kOperandKB;                 // z-feature operand, 1 k-block
  static constexpr uint32_t bars = zf + kOperandKB;
  static constexpr uint32_t total = bars + 256;
};
enum {
  B_W_FULL = 0, B_W_EMPTY = B_W_FULL + kStages, B_IN_READY = B_W_EMPTY + kStages, B_IN_FREE, B_X_FULL,
  B_AX_READY, B_H_FULL = B_AX_READY + kMTiles, B_H_EMPTY = B_H_FULL + 2, B_AH_READY = B_H_EMPTY + 2,
  B_AH_FREE = B_AH_READY + 2, B_COUNT = B_AH_FREE + 2
};

__device__ __forceinline__ void store_bf16(uint8_t* base, uint32_t off, float v) {
  *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16_rn(v);
}

template <int NS>
__global__ void __launch_bounds__(kThreads, 1)
field_umma_kernel(const pnr_scene sc, const pnr_points q, const uint8_t* __restrict__ stages,
                  const float* __restrict__ bias_x, const float* __restrict__ bias_h,
                  const float* __restrict__ bias_out, float* __restrict__ out, const Sched sch, const int num_freqs,
                  const float freq_factor, const int tiles_per_obj, const int n_tiles, const int d_out,
                  const int raw_out) {
  constexpr int PP = kNCol / NS;                       // points per tile
  constexpr int NPOST = ((PP + 15) / 16) * 16;         // UMMA N after the view mean
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  volatile uint32_t* tmem_base_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::bars + 8 * B_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) -> uint32_t { return sbase + Smem::bars + 8u * i; };
  if ((sbase & 1023u) != 0) { if (threadIdx.x == 0) printf("pnr: dynamic smem not 1024-aligned\n"); __trap(); }

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar(B_W_FULL + i), 1); mbar_init(bar(B_W_EMPTY + i), 1); }
    mbar_init(bar(B_IN_READY), kGatherWarps);
    mbar_init(bar(B_IN_FREE), 1);
    mbar_init(bar(B_X_FULL), 1);
    for (int i = 0; i < kMTiles; ++i) mbar_init(bar(B_AX_READY + i), 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_H_FULL + i), 1); mbar_init(bar(B_H_EMPTY + i), 4);
      mbar_init(bar(B_AH_READY + i), 4); mbar_init(bar(B_AH_FREE + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(sbase + Smem::bars + 8 * B_COUNT, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int n_stages = sched_total(sch);
  const uint32_t idesc_pre = instr_desc_bf16(kNCol), idesc_post = instr_desc_bf16(NPOST);

  if (warp == 0) {
    // ===================== weight producer: stream the packed stages through the smem ring ==========
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int s = 0; s < n_stages; ++s, ++g) {
        const uint32_t slot = g % kStages;
        mbar_wait(bar(B_W_EMPTY + slot), ((g / kStages) & 1) ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar(B_W_FULL + slot), kStageBytes);
          bulk_g2s(sbase + Smem::w + slot * kStageBytes, stages + (size_t)s * kStageBytes, kStageBytes,
                   bar(B_W_FULL + slot));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===============================================================
    uint32_t gc = 0;            // weight stages consumed
    uint32_t ph = 0;            // parity bits of the barriers this warp waits on (bit = barrier id)
    ph |= (1u << (B_H_EMPTY + 0)) | (1u << (B_H_EMPTY + 1));    // "empty"-type: first wait passes
    auto wait = [&](int id) { mbar_wait(bar(id), (ph >> id) & 1u); ph ^= (1u << id); tc_fence_after(); };
    // consume one weight stage: D[d_col] (+)= W_stage * B_kblock^T
    auto step = [&](uint32_t b_addr, uint32_t d_col, uint32_t idesc, bool acc) {
      const uint32_t slot = gc % kStages;
      mbar_wait(bar(B_W_FULL + slot), (gc / kStages) & 1);
      tc_fence_after();
      if (lane == 0) {
        mma_kblock(tmem_base + d_col, sbase + Smem::w + slot * kStageBytes, b_addr, idesc, acc);
        mma_commit(bar(B_W_EMPTY + slot));
      }
      __syncwarp();
      ++gc;
    };
    auto commit = [&](int id) { if (lane == 0) mma_commit(bar(id)); __syncwarp(); };
    auto issue_linz = [&](int l) {
      for (int mt = 0; mt < kMTiles; ++mt)
        for (int kb = 0; kb < sch.KBz; ++kb) step(sbase + Smem::lat + kb * kOperandKB, mt * kNCol, idesc_pre, true);
      if (l == sch.n_linz - 1) commit(B_IN_FREE);     // latent / z-feature buffers may be refilled
    };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      wait(B_IN_READY);
      for (int mt = 0; mt < kMTiles; ++mt) step(sbase + Smem::zf, mt * kNCol, idesc_pre, false);   // lin_in
      issue_linz(0);
      commit(B_X_FULL);
      for (int b = 0; b < sch.n_blocks; ++b) {
        const uint32_t idesc = b < sch.CL ? idesc_pre : idesc_post;
        for (int j = 0; j <= kMTiles; ++j) {
          if (j < kMTiles) {                           // S2(j): h[j&1] = fc_0 rows [128j,128j+128) * relu(x)
            wait(B_H_EMPTY + (j & 1));
            for (int kb = 0; kb < kKBlocksH; ++kb) {
              if (j == 0 && (kb & 1) == 0) wait(B_AX_READY + (kb >> 1));
              step(sbase + Smem::ax + kb * kOperandKB, kHCol + (j & 1) * kNCol, idesc, kb > 0);
            }
            commit(B_H_FULL + (j & 1));
          }
          if (j >= 1) {                                // S3p(j-1): x[mt] += fc_1[:, chunk j-1] * relu(h chunk)
            const int jj = j - 1;
            wait(B_AH_READY + (jj & 1));
            for (int mt = 0; mt < kMTiles; ++mt)
              for (int kk = 0; kk < 2; ++kk)
                step(sbase + Smem::ah + ((jj & 1) * 2 + kk) * kOperandKB, mt * kNCol, idesc, true);
            commit(B_AH_FREE + (jj & 1));
          }
        }
        if (b + 1 < sch.n_linz) issue_linz(b + 1);
        commit(B_X_FULL);
      }
      // lin_out: rows 0..d_out-1 of a zero-padded 128-row tile -> h[0]
      wait(B_H_EMPTY + 0);
      for (int kb = 0; kb < kKBlocksH; ++kb) {
        if ((kb & 1) == 0) wait(B_AX_READY + (kb >> 1));
        step(sbase + Smem::ax + kb * kOperandKB, kHCol, idesc_post, kb > 0);
      }
      commit(B_H_FULL + 0);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue warps ==========================================================
    const int qd = warp - 4;                       // TMEM lane quadrant this warp may access
    const int fl = qd * 32 + lane;                 // feature row inside a 128-row tile
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int kbl = fl >> 6, kk = fl & 63;
    uint32_t ph = 0;
    ph |= (1u << (B_AH_FREE + 0)) | (1u << (B_AH_FREE + 1));
    auto wait = [&](int id) { mbar_wait(bar(id), (ph >> id) & 1u); ph ^= (1u << id); tc_fence_after(); };
    auto arrive = [&](int id) { __syncwarp(); if (lane == 0) mbar_arrive(bar(id)); };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int obj = tile / tiles_per_obj;
      const int p0 = (tile - obj * tiles_per_obj) * PP;
      for (int e = 0; e <= sch.n_blocks; ++e) {
        // ---- x epilogue e: relu(x + cumulative bias) -> bf16 operand (and the view mean at e == CL)
        wait(B_X_FULL);
        for (int mt = 0; mt < kMTiles; ++mt) {
          const float bias = bias_x[e * kHidden + mt * 128 + fl];
          uint8_t* dst = smem + Smem::ax + (2 * mt + kbl) * kOperandKB;
          if (e < sch.CL) {
            uint32_t r[kNCol];
            tmem_ld<kNCol>(tlane + mt * kNCol, r);
#pragma unroll
            for (int c = 0; c < kNCol; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
          } else if (e == sch.CL) {
            uint32_t r[kNCol];
            tmem_ld<kNCol>(tlane + mt * kNCol, r);
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
#pragma unroll
            for (int c = 0; c < NPOST; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
          }
          tc_fence_before();
          fence_proxy_async();
          arrive(B_AX_READY + mt);
        }
        if (e < sch.n_blocks) {
          // ---- h epilogues of block e: relu(fc_0 out + b) -> bf16 K-chunk for fc_1
          for (int j = 0; j < kMTiles; ++j) {
            const int i = j & 1;
            const float bias = bias_h[e * kHidden + j * 128 + fl];
            uint8_t* dst = smem + Smem::ah + (i * 2 + kbl) * kOperandKB;
            wait(B_H_FULL + i);
            if (e < sch.CL) {
              uint32_t r[kNCol];
              tmem_ld<kNCol>(tlane + kHCol + i * kNCol, r);
              tc_fence_before();
              arrive(B_H_EMPTY + i);
              wait(B_AH_FREE + i);
#pragma unroll
              for (int c = 0; c < kNCol; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
            } else {
              uint32_t r[NPOST];
              tmem_ld<NPOST>(tlane + kHCol + i * kNCol, r);
              tc_fence_before();
              arrive(B_H_EMPTY + i);
              wait(B_AH_FREE + i);
#pragma unroll
              for (int c = 0; c < NPOST; ++c) store_bf16(dst, swz_offset(c, kk), fmaxf(__uint_as_float(r[c]) + bias, 0.f));
            }
            fence_proxy_async();
            arrive(B_AH_READY + i);
          }
        }
      }
      // ---- output epilogue: lin_out rows live in TMEM lanes 0..d_out-1 of h[0]
      wait(B_H_FULL + 0);
      {
        uint32_t r[NPOST];
        tmem_ld<NPOST>(tlane + kHCol, r);
        tc_fence_before();
        arrive(B_H_EMPTY + 0);
        if (qd == 0 && lane < d_out) {
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
  } else {
    // ===================== gather warps: project + 4-tap bilinear gather + positional encoding =======
    const int gw = warp < 4 ? warp - 2 : warp - 6;   // 0..5
    uint32_t par_free = 1;                           // "empty"-type barrier: first wait passes
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int obj = tile / tiles_per_obj;
      const int p0 = (tile - obj * tiles_per_obj) * PP;
      mbar_wait(bar(B_IN_FREE), par_free);
      par_free ^= 1;
      for (int c = gw; c < kNCol; c += kGatherWarps) {
        const int v = c / PP, p = c - v * PP;
        const bool valid = (v < NS) && (p0 + p < q.P);
        Projection pr;
        Taps tp;
        const int view = obj * NS + (v < NS ? v : 0);
        if (valid) {
          float px, py, pz, vx, vy, vz;
          fetch_point(q, (long long)obj * q.P + p0 + p, px, py, pz, vx, vy, vz);
          pr = project_point(sc, view, px, py, pz, vx, vy, vz);
          tp = make_taps(pr.ix, pr.iy, sc.Hl, sc.Wl, sc.C);
        }
        // z-feature row: 64 bf16 (d_in <= 64 used), two per lane
        {
          const int j0 = lane * 2;
          float a = 0.f, b = 0.f;
          const int d_in = 6 * num_freqs + 6;
          if (valid) {
            if (j0 < d_in) a = zfeat_value(pr, j0, num_freqs, freq_factor);
            if (j0 + 1 < d_in) b = zfeat_value(pr, j0 + 1, num_freqs, freq_factor);
          }
          *reinterpret_cast<__nv_bfloat162*>(smem + Smem::zf + swz_offset(c, j0)) = __floats2bfloat162_rn(a, b);
        }
        // latent row: C = 512 channels, lane covers [16*lane, 16*lane+16) = two 16-byte chunks of k-block lane/4
        {
          float acc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = 0.f;
          if (valid) {
            const __nv_bfloat16* fmap = (const __nv_bfloat16*)sc.feat + (size_t)view * sc.Hl * sc.Wl * sc.C + lane * 16;
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

  tc_fence_before();
  __syncthreads();
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
  return PNR_OK;
}

int field_forward_umma(const pnr_scene* sc, const pnr_points* q, const pnr_mlp_params* mp, const void* packed,
                       float* out, int num_freqs, float freq_factor, int raw, cudaStream_t st) {
  Sched sch;
  int rc = make_sched(mp, &sch, "field_forward_umma");
  if (rc) return rc;
  PNR_REQUIRE(packed, PNR_ERR_ARG, "field_forward_umma: packed weights missing (call pnr_mlp_pack)");
  PNR_REQUIRE(((uintptr_t)packed & 1023) == 0, PNR_ERR_ARG, "field_forward_umma: packed blob must be 1024-byte aligned");
  PNR_REQUIRE(!sc->feat_fp32, PNR_ERR_ARG, "field_forward_umma: needs bf16 channels-last feature maps");
  PNR_REQUIRE(sc->C == 512 && mp->d_latent == 512, PNR_ERR_UNSUPPORTED,
              "field_forward_umma: latent size %d (this build keeps a 512-channel latent tile resident)", sc->C);
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
  const int grid = n_tiles < sms ? n_tiles : sms;
  const PackOffsets po = pack_offsets(sch);
  const uint8_t* blob = (const uint8_t*)packed;
  const uint8_t* stages = blob + po.stages;
  const float* bx = (const float*)(blob + po.bias_x);
  const float* bh = (const float*)(blob + po.bias_h);
  const float* bo = (const float*)(blob + po.bias_out);
#define PNR_LAUNCH_NS(NSV)                                                                                   \
  case NSV: {                                                                                                \
    cudaError_t e = cudaFuncSetAttribute(field_umma_kernel<NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)Smem::total);                                                  \
    PNR_REQUIRE(e == cudaSuccess, PNR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));         \
    field_umma_kernel<NSV><<<grid, kThreads, Smem::total, st>>>(*sc, *q, stages, bx, bh, bo, out, sch,      \
                                                                num_freqs, freq_factor, tiles_per_obj,      \
                                                                n_tiles, mp->d_out, raw);                   \
  } break;
  switch (sc->NS) {
    PNR_LAUNCH_NS(1) PNR_LAUNCH_NS(2) PNR_LAUNCH_NS(3) PNR_LAUNCH_NS(4) PNR_LAUNCH_NS(5) PNR_LAUNCH_NS(6) PNR_LAUNCH_NS(8)
    default: PNR_REQUIRE(false, PNR_ERR_UNSUPPORTED, "field_forward_umma: NS=%d", sc->NS);
  }
#undef PNR_LAUNCH_NS
  PNR_CHECK_LAUNCH("field_umma_kernel");
  return PNR_OK;
}

}  // namespace pnr

using namespace pnr;

extern "C" size_t pnr_mlp_pack_bytes(const pnr_mlp_params* p) {
  Sched s;
  if (make_sched(p, &s, "pnr_mlp_pack_bytes")) return 0;
  return pack_offsets(s).total;
}

extern "C" int pnr_mlp_pack(const pnr_mlp_params* p, void* packed, void* stream) {
  reset_launch_count();
  Sched s;
  int rc = make_sched(p, &s, "pnr_mlp_pack");
  if (rc) return rc;
  PNR_REQUIRE(packed && ((uintptr_t)packed & 1023) == 0, PNR_ERR_ARG, "pnr_mlp_pack: packed must be non-null and 1024-byte aligned");
  PNR_REQUIRE(p->lin_in_w && p->lin_in_b && p->lin_out_w && p->lin_out_b, PNR_ERR_ARG, "pnr_mlp_pack: null lin_in/lin_out");
  for (int b = 0; b < p->n_blocks; ++b)
    PNR_REQUIRE(p->fc0_w[b] && p->fc0_b[b] && p->fc1_w[b] && p->fc1_b[b], PNR_ERR_ARG, "pnr_mlp_pack: null block %d", b);
  for (int b = 0; b < s.n_linz; ++b) PNR_REQUIRE(p->linz_w[b] && p->linz_b[b], PNR_ERR_ARG, "pnr_mlp_pack: null lin_z %d", b);
  const PackOffsets po = pack_offsets(s);
  uint8_t* blob = (uint8_t*)packed;
  cudaStream_t st = (cudaStream_t)stream;
  pack_stages_kernel<<<sched_total(s), 256, 0, st>>>(*p, s, blob + po.stages);
  PNR_CHECK_LAUNCH("pack_stages_kernel");
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
