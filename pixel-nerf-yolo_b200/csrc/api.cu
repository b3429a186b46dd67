// Library-level entry points: version, thread-local error string, launch counter, field dispatch.
#include "pnr_common.cuh"
#include <string.h>

namespace pnr {
static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

int validate_scene_points(const pnr_scene* sc, const pnr_points* q, const char* who);
size_t field_workspace_fp32(const pnr_scene* sc, const pnr_points* q, int d_in, int H);
int field_forward_fp32(const pnr_scene* sc, const pnr_points* q, const pnr_mlp_params* mp, float* out, void* ws,
                       size_t ws_bytes, int num_freqs, float freq_factor, int raw, cudaStream_t st);
size_t field_workspace_umma();
int field_forward_umma(const pnr_scene* sc, const pnr_points* q, const pnr_mlp_params* mp, const void* packed,
                       float* out, void* ws, size_t ws_bytes, int num_freqs, float freq_factor, int raw,
                       cudaStream_t st);
}  // namespace pnr

using namespace pnr;

extern "C" int pnr_version(void) { return PNR_ABI_VERSION; }
extern "C" const char* pnr_last_error(void) { return g_err; }
extern "C" int pnr_last_launch_count(void) { return g_launches; }

extern "C" int pnr_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return 0; }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) { set_error("device is sm_%d%d; this library only has sm_100a code", major, minor); return 0; }
  return 1;
}

extern "C" size_t pnr_field_workspace_bytes(const pnr_scene* scene, const pnr_points* pts, int precision) {
  if (!scene || !pts) return 0;
  if (precision == PNR_PREC_FP32) return field_workspace_fp32(scene, pts, 64, kHidden);
  if (precision == PNR_PREC_BF16) return field_workspace_umma();
  return 0;
}

extern "C" int pnr_field_forward(const pnr_scene* scene, const pnr_points* pts, const pnr_mlp_params* params,
                                 const void* packed, float* out, void* workspace, size_t workspace_bytes,
                                 int precision, int num_freqs, float freq_factor, void* stream) {
  reset_launch_count();
  int rc = validate_scene_points(scene, pts, "pnr_field_forward");
  if (rc) return rc;
  PNR_REQUIRE(params && out, PNR_ERR_ARG, "pnr_field_forward: null params/out");
  cudaStream_t st = (cudaStream_t)stream;
  const int raw = (scene->flags & PNR_SCENE_RAW_OUTPUT) ? 1 : 0;
  if (precision == PNR_PREC_FP32) {
    PNR_REQUIRE(params->d_hidden == kHidden, PNR_ERR_UNSUPPORTED, "pnr_field_forward: d_hidden=%d", params->d_hidden);
    return field_forward_fp32(scene, pts, params, out, workspace, workspace_bytes, num_freqs, freq_factor, raw, st);
  }
  if (precision == PNR_PREC_BF16)
    return field_forward_umma(scene, pts, params, packed, out, workspace, workspace_bytes, num_freqs, freq_factor, raw, st);
  PNR_REQUIRE(false, PNR_ERR_ARG, "pnr_field_forward: unknown precision %d", precision);
}

// ---- NeRFRenderer.forward as ONE call (src/render/nerf.py:257-309): sample_coarse -> field -> composite -> sample_fine +
// sample_fine_depth + sort -> field -> composite, all enqueued on `stream`; every buffer is the caller's.
namespace {
struct RenderLayout { size_t z_c, out_c, w_c, z_f, out_f, field_ws, total; };
RenderLayout render_layout(const pnr_render_args* a) {
  const size_t Bt = (size_t)a->scene->SB * a->B;
  const size_t Kc = a->n_coarse, Kf = a->n_fine > 0 ? (size_t)a->n_coarse + a->n_fine : 0;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  RenderLayout L;
  size_t off = 0;
  L.z_c = off; off = al(off + Bt * Kc * sizeof(float));
  L.out_c = off; off = al(off + Bt * Kc * 4 * sizeof(float));
  L.w_c = off; off = al(off + Bt * Kc * sizeof(float));
  L.z_f = off; off = al(off + Bt * Kf * sizeof(float));
  L.out_f = off; off = al(off + Bt * Kf * 4 * sizeof(float));
  L.field_ws = off; off = al(off + pnr::field_workspace_umma());
  L.total = off;
  return L;
}
}  // namespace

extern "C" size_t pnr_render_workspace_bytes(const pnr_render_args* a) {
  if (!a || !a->scene || a->B < 0 || a->n_coarse <= 0 || a->n_fine < 0) return 0;
  return render_layout(a).total;
}

extern "C" int pnr_render_forward(const pnr_render_args* a, void* stream) {
  reset_launch_count();
  PNR_REQUIRE(a && a->scene && a->rays && a->steps && a->noise_coarse && a->mlp_coarse && a->packed_coarse, PNR_ERR_ARG,
              "pnr_render_forward: null pointer");
  PNR_REQUIRE(a->rgb_coarse && a->depth_coarse, PNR_ERR_ARG, "pnr_render_forward: coarse outputs missing");
  PNR_REQUIRE(a->precision == PNR_PREC_BF16, PNR_ERR_UNSUPPORTED, "pnr_render_forward: the single-call render is the bf16 tensor-core path");
  PNR_REQUIRE(a->B >= 0 && a->n_coarse > 0 && a->n_fine >= 0 && a->n_fine_depth >= 0 && a->n_fine_depth <= a->n_fine, PNR_ERR_ARG,
              "pnr_render_forward: bad sample counts");
  const int kf = a->n_fine - a->n_fine_depth, kfd = a->n_fine_depth;
  const bool fine = a->n_fine > 0;
  if (fine) {
    PNR_REQUIRE(a->rgb_fine && a->depth_fine, PNR_ERR_ARG, "pnr_render_forward: fine outputs missing");
    PNR_REQUIRE((kf == 0 || (a->noise_u && a->noise_jitter)) && (kfd == 0 || a->noise_gauss), PNR_ERR_ARG,
                "pnr_render_forward: fine-pass noise missing");
  }
  const RenderLayout L = render_layout(a);
  PNR_REQUIRE(a->workspace && a->workspace_bytes >= L.total && ((uintptr_t)a->workspace & 255) == 0, PNR_ERR_ARG,
              "pnr_render_forward: workspace of pnr_render_workspace_bytes() = %zu bytes (256-byte aligned) required", L.total);
  const long long Bt_ll = (long long)a->scene->SB * a->B;
  PNR_REQUIRE(Bt_ll < (1LL << 31), PNR_ERR_ARG, "pnr_render_forward: too many rays");
  const int Bt = (int)Bt_ll;
  if (Bt == 0) return PNR_OK;
  uint8_t* ws = (uint8_t*)a->workspace;
  float* z_c = (float*)(ws + L.z_c);
  float* out_c = (float*)(ws + L.out_c);
  float* w_c = a->weights_coarse ? a->weights_coarse : (float*)(ws + L.w_c);
  int launches = 0, rc;
#define RSTEP(call) do { rc = (call); if (rc) return rc; launches += pnr_last_launch_count(); } while (0)
  RSTEP(pnr_sample_coarse(a->rays, a->steps, a->noise_coarse, z_c, Bt, a->n_coarse, a->lindisp, stream));
  pnr_points pts = {};
  pts.rays = a->rays; pts.z = z_c; pts.mode = 1; pts.K = a->n_coarse; pts.P = a->B * a->n_coarse;
  auto mark = [&](int i) { if (a->field_events[i]) cudaEventRecord((cudaEvent_t)a->field_events[i], (cudaStream_t)stream); };
  mark(0);
  RSTEP(pnr_field_forward(a->scene, &pts, a->mlp_coarse, a->packed_coarse, out_c, ws + L.field_ws, L.total - L.field_ws,
                          PNR_PREC_BF16, a->num_freqs, a->freq_factor, stream));
  mark(1);
  RSTEP(pnr_composite(out_c, z_c, a->rays, w_c, a->rgb_coarse, a->depth_coarse, Bt, a->n_coarse, a->white_bkgd, stream));
  if (fine) {
    const int K = a->n_coarse + a->n_fine;
    float* z_f = (float*)(ws + L.z_f);
    float* out_f = (float*)(ws + L.out_f);
    RSTEP(pnr_sample_fine(w_c, a->depth_coarse, a->rays, z_c, a->noise_u, a->noise_jitter, a->noise_gauss, z_f, nullptr, nullptr,
                          nullptr, Bt, a->n_coarse, kf, kfd, a->depth_std, a->lindisp, stream));
    pts.z = z_f; pts.K = K; pts.P = a->B * K;
    const pnr_mlp_params* mf = a->mlp_fine ? a->mlp_fine : a->mlp_coarse;          // models.py:291: no fine network -> coarse
    const void* pf = a->mlp_fine ? a->packed_fine : a->packed_coarse;
    PNR_REQUIRE(pf, PNR_ERR_ARG, "pnr_render_forward: packed fine weights missing");
    mark(2);
    RSTEP(pnr_field_forward(a->scene, &pts, mf, pf, out_f, ws + L.field_ws, L.total - L.field_ws, PNR_PREC_BF16, a->num_freqs,
                            a->freq_factor, stream));
    mark(3);
    RSTEP(pnr_composite(out_f, z_f, a->rays, a->weights_fine, a->rgb_fine, a->depth_fine, Bt, K, a->white_bkgd, stream));
  }
#undef RSTEP
  reset_launch_count();
  count_launch(launches);
  return PNR_OK;
}
