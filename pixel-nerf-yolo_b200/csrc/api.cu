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
