// lib.cu -- library plumbing: error strings, device info, twiddle tables.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace ds {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// one table set per (device, n_fft); n_fft = 64 << idx
static const int kMinLog = 6, kMaxLog = 11, kNumSizes = kMaxLog - kMinLog + 1, kMaxDev = 16;
static TwiddleSet g_tw[kMaxDev][kNumSizes];
static bool g_tw_ok[kMaxDev][kNumSizes];
static std::mutex g_mu;

static int build_tables(int dev, int idx) {
  const int N = 64 << idx, H = N / 2;
  std::vector<float2> h32(H), n32(H / 2 + 1);
  std::vector<double2> h64(H), n64(H / 2 + 1);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int i = 0; i < H; ++i) {
    long double a = -two_pi * (long double)i / (long double)H;
    long double c = cosl(a), s = sinl(a);
    // exact values on the axes / diagonals
    if ((4 * i) % H == 0) { int q = (4 * i) / H; c = (q == 0) ? 1 : (q == 2 ? -1 : 0); s = (q == 1) ? -1 : (q == 3 ? 1 : 0); }
    h64[i] = make_double2((double)c, (double)s);
    h32[i] = make_float2((float)c, (float)s);
  }
  for (int k = 0; k <= H / 2; ++k) {
    long double a = -two_pi * (long double)k / (long double)N;
    long double c = cosl(a), s = sinl(a);
    if (k == 0) { c = 1; s = 0; }
    if (2 * k == H) { c = 0; s = -1; }
    n64[k] = make_double2((double)c, (double)s);
    n32[k] = make_float2((float)c, (float)s);
  }
  float2 *dh32, *dn32; double2 *dh64, *dn64;
  DS_CUDA(cudaMalloc(&dh32, sizeof(float2) * H));
  DS_CUDA(cudaMalloc(&dn32, sizeof(float2) * (H / 2 + 1)));
  DS_CUDA(cudaMalloc(&dh64, sizeof(double2) * H));
  DS_CUDA(cudaMalloc(&dn64, sizeof(double2) * (H / 2 + 1)));
  DS_CUDA(cudaMemcpy(dh32, h32.data(), sizeof(float2) * H, cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dn32, n32.data(), sizeof(float2) * (H / 2 + 1), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dh64, h64.data(), sizeof(double2) * H, cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dn64, n64.data(), sizeof(double2) * (H / 2 + 1), cudaMemcpyHostToDevice));
  g_tw[dev][idx].h32 = dh32; g_tw[dev][idx].n32 = dn32;
  g_tw[dev][idx].h64 = dh64; g_tw[dev][idx].n64 = dn64;
  g_tw_ok[dev][idx] = true;
  return DS_OK;
}

int get_twiddles(int n_fft, TwiddleSet *out) {
  if (!is_pow2(n_fft) || n_fft < (1 << kMinLog) || n_fft > (1 << kMaxLog)) {
    set_error("unsupported n_fft %d (power of two in [64, 2048] required)", n_fft);
    return DS_EUNSUPPORTED;
  }
  int idx = 0;
  while ((64 << idx) != n_fft) ++idx;
  int dev = 0;
  DS_CUDA(cudaGetDevice(&dev));
  if (dev >= kMaxDev) { set_error("device index %d too large", dev); return DS_EUNSUPPORTED; }
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_tw_ok[dev][idx]) {
    int rc = build_tables(dev, idx);
    if (rc != DS_OK) return rc;
  }
  *out = g_tw[dev][idx];
  return DS_OK;
}

}  // namespace ds

extern "C" {

int ds_version(void) { return DS_VERSION; }

const char *ds_last_error(void) { return ds::g_err; }

int ds_init(void) {
  for (int lg = ds::kMinLog; lg <= ds::kMaxLog; ++lg) {
    ds::TwiddleSet t;
    int rc = ds::get_twiddles(1 << lg, &t);
    if (rc != DS_OK) return rc;
  }
  return DS_OK;
}

int ds_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  DS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DS_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return DS_OK;
}

}  // extern "C"
