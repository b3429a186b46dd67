// The two steps either side of the rendering hot path:
//   * pnr_pyramid_pack: SpatialEncoder.forward's tail (src/model/encoder.py:159-168) -- bilinear upsampling
//     (align_corners=True) of every pyramid level to the first level's size, channel concat, and the repack the hot
//     path wants (NCHW fp32 -> channels-last bf16/fp32) -- in ONE pass: the fp32 NCHW `latent` (25 MB at 128^2,
//     629 MB at 640^2) is never materialised, each level is read once and the channels-last maps are written once.
//   * pnr_gen_rays: util.gen_rays / util.unproj_map (src/util/util.py:115-145, 240-278) -- per-pixel camera ray
//     [origin, direction, near, far], optionally only for a list of selected pixels (the trainer's ray sampling,
//     train/trainlib/PixelNerfTrainer.py:100-117).
// Both are HBM-bound streaming kernels: coalesced loads along W of the source level, a shared-memory transpose, and
// 16-byte channels-last stores.
#include "pnr_common.cuh"

namespace pnr {

struct PyramidLevels {
  const float* src[8];     // (N, C_l, H_l, W_l) fp32
  int C[8], H[8], W[8];
  int c0[8];               // first output channel of the level
  int n_levels;
};

// Block = 32 output pixels of one row (x0..x0+31) x 32 channels of one level; threads (32, 8).
// Phase 1: thread (tx, ty) computes channel c = cb + ty + 8 r at pixel x0 + tx (source reads coalesced along W).
// Phase 2: transposed write, lane -> channel, so the channels-last store is contiguous.
template <typename OutT>
__global__ void __launch_bounds__(256)
pyramid_pack_kernel(const PyramidLevels lv, OutT* __restrict__ dst, int Ho, int Wo, int Ctot) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int y = blockIdx.y;
  // blockIdx.x enumerates (level, channel block, x block)
  int rem = blockIdx.x, l = 0;
  const int xblocks = (Wo + 31) / 32;
  for (; l < lv.n_levels; ++l) {
    const int nb = ((lv.C[l] + 31) / 32) * xblocks;
    if (rem < nb) break;
    rem -= nb;
  }
  const int cb = (rem / xblocks) * 32, x0 = (rem % xblocks) * 32;
  const int Hl = lv.H[l], Wl = lv.W[l], Cl = lv.C[l];
  const float* src = lv.src[l] + (size_t)n * Cl * Hl * Wl;
  // F.interpolate(mode="bilinear", align_corners=True): src = dst * (in - 1) / (out - 1)       (ATen UpSample.h)
  const float sy = Ho > 1 ? (float)(Hl - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wl - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sy * (float)y;
  const int y0 = (int)fy;
  const int y1 = y0 + (y0 < Hl - 1 ? 1 : 0);
  const float ly = fy - (float)y0;
  const int x = x0 + threadIdx.x;
  const float fx = sx * (float)x;
  const int xa = (int)fx;
  const int xb = xa + (xa < Wl - 1 ? 1 : 0);
  const float lx = fx - (float)xa;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = cb + r;
    float v = 0.f;
    if (c < Cl && x < Wo) {
      const float* p = src + (size_t)c * Hl * Wl;
      const float a = p[y0 * Wl + xa], b = p[y0 * Wl + xb], cc = p[y1 * Wl + xa], d = p[y1 * Wl + xb];
      v = (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * cc + lx * d);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int xo = x0 + r, c = cb + threadIdx.x;
    if (xo < Wo && c < Cl) {
      const float v = tile[threadIdx.x][r];
      const size_t o = (((size_t)n * Ho + y) * Wo + xo) * Ctot + lv.c0[l] + c;
      if constexpr (sizeof(OutT) == 2) dst[o] = __float2bfloat16_rn(v);
      else dst[o] = v;
    }
  }
}

// util.py:115-145 + 240-278.  One thread per output ray.
__global__ void gen_rays_kernel(const float* __restrict__ poses, const long long* __restrict__ pix_inds,
                                float* __restrict__ rays, long long n_out, int N, int H, int W, float fx, float fy,
                                float cx, float cy, float z_near, float z_far) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const long long pix = pix_inds ? pix_inds[i] : i;          // flat index into (N, H, W)
  const int n = (int)(pix / ((long long)H * W));
  const int yx = (int)(pix - (long long)n * H * W);
  const int y = yx / W, x = yx - y * W;
  // unproj_map: X = (x - cx) / fx, Y = (y - cy) / fy, d = (X, -Y, -1) / |d|
  const float X = __fdiv_rn(__fsub_rn((float)x, cx), fx);
  const float Y = __fdiv_rn(__fsub_rn((float)y, cy), fy);
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(X, X), __fmul_rn(Y, Y)), 1.0f));
  const float dx = __fdiv_rn(X, nrm), dy = __fdiv_rn(-Y, nrm), dz = __fdiv_rn(-1.0f, nrm);
  const float* P = poses + (size_t)n * 16;                  // camera-to-world (4, 4)
  float* o = rays + i * 8;
  o[0] = P[3]; o[1] = P[7]; o[2] = P[11];
  o[3] = P[0] * dx + P[1] * dy + P[2] * dz;                 // R d
  o[4] = P[4] * dx + P[5] * dy + P[6] * dz;
  o[5] = P[8] * dx + P[9] * dy + P[10] * dz;
  o[6] = z_near; o[7] = z_far;
}


// Output side.
// eval/eval.py:283-290: depth -> (depth - z_near) / (z_far - z_near); rgb -> uint8(clamp(rgb, 0, 1) * 255).
__global__ void image_output_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, uint8_t* __restrict__ rgb_u8,
                                    float* __restrict__ depth_norm, long long B, float z_near, float z_far) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  if (rgb_u8) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fminf(fmaxf(rgb[i * 3 + c], 0.0f), 1.0f);           // torch.clamp (NaN propagates -> 0 below)
      rgb_u8[i * 3 + c] = (uint8_t)(int)__fmul_rn(v, 255.0f);             // numpy astype(uint8): truncation
    }
  }
  if (depth_norm) depth_norm[i] = __fdiv_rn(__fsub_rn(depth[i], z_near), __fsub_rn(z_far, z_near));
}

// PixelNerfTrainer.py:147-157 + src/model/loss.py:92-104: mean squared (or absolute) error between rendered and ground
// truth rgb, and its gradient, in one pass.  loss is accumulated with one atomic per block (double partial sums).
__global__ void __launch_bounds__(256)
rgb_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, float* __restrict__ loss,
                float* __restrict__ d_rgb, long long n, int use_l1, float grad_scale) {
  __shared__ double part[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(rgb[i], gt[i]);
    acc += use_l1 ? (double)fabsf(d) : (double)d * (double)d;
    if (d_rgb) d_rgb[i] = use_l1 ? (d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f)) : 2.0f * d * grad_scale;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += part[w];
    atomicAdd(loss, (float)(t / (double)n));
  }
}

// util.gen_rays_yolo (src/util/util.py:808-876): one thread per ray.  pixel = (x + 0.49, y + 0.49, 1) (grid of cell indices,
// :826-835), camera-space direction = inv(K) . pixel (:838), world direction = inv(E)[:3,:3] . dir (:857, not normalised),
// origin = inv(E)[:3,3] (:860).  The two products are written out as the fused multiply-add chains of a row-times-column dot
// product (what the sgemm behind torch.matmul does for a K=3 contraction up to its internal order).
__global__ void gen_rays_yolo_kernel(const float* __restrict__ inv_intr, const float* __restrict__ inv_extr, float* __restrict__ rays,
                                     int N, int H, int W, float z_near, float z_far) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const int n = (int)(i / ((long long)W * H));
  const float px = (float)x + 0.49f, py = (float)y + 0.49f;
  float dc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) dc[r] = fmaf(inv_intr[r * 3 + 2], 1.0f, fmaf(inv_intr[r * 3 + 1], py, inv_intr[r * 3 + 0] * px));
  const float* E = inv_extr + (size_t)n * 16;
  float* o = rays + i * 8;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[r] = E[r * 4 + 3];
    o[3 + r] = fmaf(E[r * 4 + 2], dc[2], fmaf(E[r * 4 + 1], dc[1], E[r * 4 + 0] * dc[0]));
  }
  o[6] = z_near;
  o[7] = z_far;
}

}  // namespace pnr

using namespace pnr;

extern "C" int pnr_pyramid_pack(const float* const* levels, const int32_t* level_C, const int32_t* level_H,
                                const int32_t* level_W, int n_levels, int N, void* dst, int to_fp32, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(levels && level_C && level_H && level_W && dst, PNR_ERR_ARG, "pnr_pyramid_pack: null pointer");
  PNR_REQUIRE(n_levels >= 1 && n_levels <= 8 && N >= 1 && N <= 65535, PNR_ERR_ARG, "pnr_pyramid_pack: bad level/map count");
  PyramidLevels lv = {};
  lv.n_levels = n_levels;
  int ctot = 0;
  long long blocks = 0;
  const int Ho = level_H[0], Wo = level_W[0];
  PNR_REQUIRE(Ho >= 1 && Wo >= 1 && Ho <= 65535, PNR_ERR_ARG, "pnr_pyramid_pack: bad output size");
  for (int l = 0; l < n_levels; ++l) {
    PNR_REQUIRE(levels[l] && level_C[l] > 0 && level_H[l] > 0 && level_W[l] > 0, PNR_ERR_ARG, "pnr_pyramid_pack: bad level %d", l);
    lv.src[l] = levels[l]; lv.C[l] = level_C[l]; lv.H[l] = level_H[l]; lv.W[l] = level_W[l]; lv.c0[l] = ctot;
    ctot += level_C[l];
    blocks += (long long)((level_C[l] + 31) / 32) * ((Wo + 31) / 32);
  }
  PNR_REQUIRE(blocks < (1LL << 31), PNR_ERR_ARG, "pnr_pyramid_pack: too many blocks");
  dim3 grid((unsigned)blocks, (unsigned)Ho, (unsigned)N), block(32, 8);
  if (to_fp32) pyramid_pack_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(lv, (float*)dst, Ho, Wo, ctot);
  else pyramid_pack_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(lv, (__nv_bfloat16*)dst, Ho, Wo, ctot);
  PNR_CHECK_LAUNCH("pyramid_pack_kernel");
  return PNR_OK;
}

extern "C" int pnr_gen_rays(const float* poses, const long long* pix_inds, float* rays, long long n_out, int N, int H,
                            int W, float fx, float fy, float cx, float cy, float z_near, float z_far, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(poses && rays, PNR_ERR_ARG, "pnr_gen_rays: null pointer");
  PNR_REQUIRE(N >= 1 && H >= 1 && W >= 1 && n_out >= 0, PNR_ERR_ARG, "pnr_gen_rays: bad shape");
  PNR_REQUIRE(pix_inds || n_out == (long long)N * H * W, PNR_ERR_ARG, "pnr_gen_rays: n_out must be N*H*W without pix_inds");
  PNR_REQUIRE(fx != 0.f && fy != 0.f, PNR_ERR_ARG, "pnr_gen_rays: zero focal length");
  if (n_out == 0) return PNR_OK;
  gen_rays_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, (cudaStream_t)stream>>>(poses, pix_inds, rays, n_out, N, H, W, fx,
                                                                                    fy, cx, cy, z_near, z_far);
  PNR_CHECK_LAUNCH("gen_rays_kernel");
  return PNR_OK;
}

extern "C" int pnr_gen_rays_yolo(const float* inv_intr, const float* inv_extr, float* rays, int N, int H, int W, float z_near,
                                 float z_far, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(inv_intr && inv_extr && rays, PNR_ERR_ARG, "pnr_gen_rays_yolo: null pointer");
  PNR_REQUIRE(N >= 1 && H >= 1 && W >= 1, PNR_ERR_ARG, "pnr_gen_rays_yolo: bad shape");
  const long long n = (long long)N * H * W;
  gen_rays_yolo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(inv_intr, inv_extr, rays, N, H, W, z_near, z_far);
  PNR_CHECK_LAUNCH("gen_rays_yolo_kernel");
  return PNR_OK;
}

extern "C" int pnr_image_output(const float* rgb, const float* depth, uint8_t* rgb_u8, float* depth_norm, long long B,
                                float z_near, float z_far, void* stream) {
  reset_launch_count();
  PNR_REQUIRE((rgb_u8 == nullptr || rgb) && (depth_norm == nullptr || depth) && (rgb_u8 || depth_norm), PNR_ERR_ARG,
              "pnr_image_output: null pointer");
  PNR_REQUIRE(B >= 0, PNR_ERR_ARG, "pnr_image_output: bad shape");
  if (B == 0) return PNR_OK;
  image_output_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rgb, depth, rgb_u8, depth_norm, B, z_near, z_far);
  PNR_CHECK_LAUNCH("image_output_kernel");
  return PNR_OK;
}

extern "C" int pnr_rgb_loss(const float* rgb, const float* gt, float* loss, float* d_rgb, long long n, int use_l1,
                            void* stream) {
  reset_launch_count();
  PNR_REQUIRE(rgb && gt && loss, PNR_ERR_ARG, "pnr_rgb_loss: null pointer");
  PNR_REQUIRE(n > 0, PNR_ERR_ARG, "pnr_rgb_loss: empty input (torch's mean loss of nothing is NaN; refuse instead)");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  rgb_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rgb, gt, loss, d_rgb, n, use_l1, 1.0f / (float)n);
  PNR_CHECK_LAUNCH("rgb_loss_kernel");
  return PNR_OK;
}
