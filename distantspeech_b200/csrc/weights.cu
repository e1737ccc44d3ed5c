// weights.cu -- small per-bin helpers exposed for API parity with the module
// functions of beamformer/beamformer.py: compute_mvdr_weight (:133-155),
// compute_pmwf_weight (:100-130) and the weight apply
// Y = einsum('ij,ij->i', W.conj(), X)  (fixedbeamformer.py:163).
// Arrays use the reference's own row-major layouts ([bins, M], [bins, M, M]).
#include "common.cuh"
#include "perbin.cuh"

namespace ds {

constexpr int WMAX = 16;

// w = R a / (a^H R a)
__global__ void mvdr_weight_kernel(int B, int M, const double2 *__restrict__ steer, const double2 *__restrict__ Rinv,
                                   double2 *__restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double2 *a = steer + (long long)b * M;
  const double2 *R = Rinv + (long long)b * M * M;
  double2 num[WMAX];
  double dr = 0.0, di = 0.0;
  for (int i = 0; i < M; ++i) {
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < M; ++j) {
      const double2 r = R[i * M + j], v = a[j];
      sr += r.x * v.x - r.y * v.y;
      si += r.x * v.y + r.y * v.x;
    }
    num[i] = make_double2(sr, si);
    // conj(a_i) * num_i
    dr += a[i].x * sr + a[i].y * si;
    di += a[i].x * si - a[i].y * sr;
  }
  const double dn = 1.0 / (dr * dr + di * di);
  for (int i = 0; i < M; ++i) {
    // num / den = num * conj(den) / |den|^2
    out[(long long)b * M + i] = make_double2((num[i].x * dr + num[i].y * di) * dn, (num[i].y * dr - num[i].x * di) * dn);
  }
}

// w = (Rvv_inv @ Rxx)[:, 0] / (beta + xi)
__global__ void pmwf_weight_kernel(int B, int M, const double *__restrict__ xi, const double2 *__restrict__ Rxx,
                                   const double2 *__restrict__ Rvv_inv, double beta, double2 *__restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double2 *X = Rxx + (long long)b * M * M;
  const double2 *R = Rvv_inv + (long long)b * M * M;
  const double dn = 1.0 / (beta + xi[b]);
  for (int i = 0; i < M; ++i) {
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < M; ++j) {
      const double2 r = R[i * M + j], v = X[j * M + 0];
      sr += r.x * v.x - r.y * v.y;
      si += r.x * v.y + r.y * v.x;
    }
    out[(long long)b * M + i] = make_double2(sr * dn, si * dn);
  }
}

// Y[s,t,k] = sum_m conj(W[k,m]) X[s,t,m,k]
template <typename XT>
__global__ void apply_weights_kernel(long long ST, int M, int K, const XT *__restrict__ X, const double2 *__restrict__ W,
                                     double2 *__restrict__ Y) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ST * K) return;
  const long long st = g / K;
  const int k = (int)(g % K);
  double yr = 0.0, yi = 0.0;
  for (int m = 0; m < M; ++m) {
    const XT x = X[(st * M + m) * K + k];
    const double2 w = W[(long long)k * M + m];
    yr += w.x * (double)x.x + w.y * (double)x.y;
    yi += w.x * (double)x.y - w.y * (double)x.x;
  }
  Y[g] = make_double2(yr, yi);
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_mvdr_weight_run(int n_bins, int n_mics, const void *steer, const void *Rvv_inv, void *w_out, void *stream) {
  DS_CHECK_ARG(steer && Rvv_inv && w_out, "ds_mvdr_weight_run: null argument");
  DS_CHECK_ARG(n_bins >= 1 && n_mics >= 1 && n_mics <= WMAX, "ds_mvdr_weight_run: n_mics must be 1..%d", WMAX);
  mvdr_weight_kernel<<<(n_bins + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_bins, n_mics, (const double2 *)steer,
                                                                               (const double2 *)Rvv_inv, (double2 *)w_out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int ds_pmwf_weight_run(int n_bins, int n_mics, const double *xi, const void *Rxx, const void *Rvv_inv, double beta,
                       void *w_out, void *stream) {
  DS_CHECK_ARG(xi && Rxx && Rvv_inv && w_out, "ds_pmwf_weight_run: null argument");
  DS_CHECK_ARG(n_bins >= 1 && n_mics >= 1, "ds_pmwf_weight_run: bad shape");
  pmwf_weight_kernel<<<(n_bins + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_bins, n_mics, xi, (const double2 *)Rxx,
                                                                               (const double2 *)Rvv_inv, beta, (double2 *)w_out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int ds_apply_weights_run(int n_streams, int n_frames, int n_mics, int n_bins, const void *X, int x_is_c128, const void *W,
                         void *Y, void *stream) {
  DS_CHECK_ARG(X && W && Y, "ds_apply_weights_run: null argument");
  DS_CHECK_ARG(n_streams >= 1 && n_frames >= 1 && n_mics >= 1 && n_bins >= 1, "ds_apply_weights_run: bad shape");
  const long long ST = (long long)n_streams * n_frames;
  const unsigned blocks = (unsigned)((ST * n_bins + 255) / 256);
  if (x_is_c128)
    apply_weights_kernel<double2><<<blocks, 256, 0, (cudaStream_t)stream>>>(ST, n_mics, n_bins, (const double2 *)X, (const double2 *)W, (double2 *)Y);
  else
    apply_weights_kernel<float2><<<blocks, 256, 0, (cudaStream_t)stream>>>(ST, n_mics, n_bins, (const float2 *)X, (const double2 *)W, (double2 *)Y);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"

// ---- OMLSA gain (mcspp_base.py:140-155) -----------------------------------
namespace ds {
__global__ void omlsa_gain_kernel(long long n, int K, const double *__restrict__ xi, const double *__restrict__ p, double Gmin,
                                  double *__restrict__ G, double *__restrict__ GH1) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const double h = xi[g] / (1.0 + xi[g]);
  double v = pow(h, p[g]) * pow(Gmin, 1.0 - p[g]);
  v = fmax(fmin(v, 1.0), Gmin);
  if ((int)(g % K) < 2) v = 0.0;
  G[g] = v;
  if (GH1) GH1[g] = h;
}
}  // namespace ds

extern "C" int ds_omlsa_gain_run(int n_rows, int n_bins, const double *xi, const double *p, double Gmin, double *G,
                                 double *G_H1, void *stream) {
  DS_CHECK_ARG(xi && p && G, "ds_omlsa_gain_run: null argument");
  DS_CHECK_ARG(n_rows >= 1 && n_bins >= 1, "ds_omlsa_gain_run: bad shape");
  const long long n = (long long)n_rows * n_bins;
  ds::omlsa_gain_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, n_bins, xi, p, Gmin, G, G_H1);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

// ---- spectral power / gain (FDGSC.py:288-294) -------------------------------------------
namespace ds {
template <typename C2>
__global__ void power_kernel(long long n, const C2 *__restrict__ X, int via_abs, double *__restrict__ out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const C2 v = X[g];
  if (via_abs) {                       // np.abs(X) ** 2: hypot, then squared (mcra.py:29-30)
    const double h = hypot((double)v.x, (double)v.y);
    out[g] = __dmul_rn(h, h);
  } else {                             // np.real(X * np.conj(X))
    out[g] = power_c((double)v.x, (double)v.y);
  }
}
template <typename C2>
__global__ void spectral_gain_kernel(long long n, const C2 *__restrict__ Yin, const double *__restrict__ G, int take_sqrt,
                                     double2 *__restrict__ Yout) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const C2 v = Yin[g];
  const double w = take_sqrt ? sqrt(G[g]) : G[g];
  Yout[g] = make_double2((double)v.x * w, (double)v.y * w);
}
}  // namespace ds

extern "C" int ds_power_run(long long n, const void *X, int x_is_c128, int via_abs, double *out, void *stream) {
  DS_CHECK_ARG(X && out && n >= 1, "ds_power_run: bad argument");
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (x_is_c128) ds::power_kernel<double2><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, (const double2 *)X, via_abs, out);
  else ds::power_kernel<float2><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, (const float2 *)X, via_abs, out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_spectral_gain_run(long long n, const void *Yin, int y_is_c128, const double *G, int take_sqrt, void *Yout,
                                    void *stream) {
  DS_CHECK_ARG(Yin && G && Yout && n >= 1, "ds_spectral_gain_run: bad argument");
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (y_is_c128)
    ds::spectral_gain_kernel<double2><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, (const double2 *)Yin, G, take_sqrt, (double2 *)Yout);
  else
    ds::spectral_gain_kernel<float2><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, (const float2 *)Yin, G, take_sqrt, (double2 *)Yout);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

// ---- PCM ingest / egress (beamformer/utils.py:182-196) --------------------------------
namespace ds {
// load_audio: int16 -> float32 / float(iinfo(int16).max)   (:184-185, note 32767 not 32768)
__global__ void pcm16_to_float_kernel(const short *__restrict__ in, float *__restrict__ out, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n && ((reinterpret_cast<size_t>(in + i) & 15) == 0) && ((reinterpret_cast<size_t>(out + i) & 15) == 0)) {
    const int4 v = *reinterpret_cast<const int4 *>(in + i);
    const short *s = reinterpret_cast<const short *>(&v);
    float4 a, b;
    a.x = __fdiv_rn((float)s[0], 32767.0f); a.y = __fdiv_rn((float)s[1], 32767.0f);
    a.z = __fdiv_rn((float)s[2], 32767.0f); a.w = __fdiv_rn((float)s[3], 32767.0f);
    b.x = __fdiv_rn((float)s[4], 32767.0f); b.y = __fdiv_rn((float)s[5], 32767.0f);
    b.z = __fdiv_rn((float)s[6], 32767.0f); b.w = __fdiv_rn((float)s[7], 32767.0f);
    *reinterpret_cast<float4 *>(out + i) = a;
    *reinterpret_cast<float4 *>(out + i + 4) = b;
  } else {
    for (size_t j = i; j < n && j < i + 8; ++j) out[j] = __fdiv_rn((float)in[j], 32767.0f);
  }
}
// save_audio: (audio * 32767).astype(int16)   (:193) -- C cast truncates toward zero like numpy
// float32 input: NumPy keeps the product in float32 (array * Python int); float64 input: product in double
template <typename T>
__global__ void float_to_pcm16_kernel(const T *__restrict__ in, short *__restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (short)(int)(in[i] * (T)32767);
}
}  // namespace ds

extern "C" int ds_pcm16_to_float_run(size_t n, const void *pcm, float *out, void *stream) {
  DS_CHECK_ARG(pcm && out, "ds_pcm16_to_float_run: null argument");
  if (n == 0) return DS_OK;
  const size_t threads = (n + 7) / 8;
  ds::pcm16_to_float_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const short *)pcm, out, n);
  DS_LAUNCH_CHECK();
  return DS_OK;
}
extern "C" int ds_float_to_pcm16_run(size_t n, const float *in, void *pcm, void *stream) {
  DS_CHECK_ARG(pcm && in, "ds_float_to_pcm16_run: null argument");
  if (n == 0) return DS_OK;
  ds::float_to_pcm16_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (short *)pcm, n);
  DS_LAUNCH_CHECK();
  return DS_OK;
}
extern "C" int ds_double_to_pcm16_run(size_t n, const double *in, void *pcm, void *stream) {
  DS_CHECK_ARG(pcm && in, "ds_double_to_pcm16_run: null argument");
  if (n == 0) return DS_OK;
  ds::float_to_pcm16_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (short *)pcm, n);
  DS_LAUNCH_CHECK();
  return DS_OK;
}
