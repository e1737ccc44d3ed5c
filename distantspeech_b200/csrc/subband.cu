// subband.cu -- STFT-domain ("subband") NLMS filters and the small time-domain helpers SubbandGSC needs.
//
//   SubbandLMS.update    adaptivefilter/SubbandLMS.py:28-84      (one input channel)
//   SubbandLmsMc.update  adaptivefilter/SubbandLmsMc.py:144-191  (C input channels)
//   SubbandAF            adaptivefilter/SubbandAF.py:41-49 (tap shift), :84-87 (weight update)
//
// Per bin and frame:  buf <- shift in X;  out = sum conj(W) buf;  err = D - out p;
//                     P = alpha P + (1 - alpha) sum |buf|^2 / C;  W += 2 mu p buf conj(err) / (P + eps)
// One thread per (stream, filter, bin), frames sequential, state float64 / complex128 like the reference.
#include "common.cuh"
#include "perbin.cuh"

namespace ds {

struct NlmsArgs {
  double *state;            // [S][F][NE][K]
  const float2 *X;          // [S][T][Cx][K]  (Cx = C, shared by the F filters of a stream)
  const float2 *D;          // [S][T][F][K]
  const double *p;          // [S][T][K] or null (p = 1)
  double2 *Err;             // [S][T][F][K]
  int S, F, K, T, C, L, one_minus_p;
  double mu, alpha, eps;
};
// state per (stream, filter, bin): W re/im [L*C], buf re/im [L*C] (newest tap first), P
__host__ __device__ inline int nlms_state_elems(int L, int C) { return 4 * L * C + 1; }

template <int L, int C>
__global__ void subband_nlms_kernel(NlmsArgs a) {
  constexpr int LC = L * C;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * a.F * a.K) return;
  const int k = (int)(g % a.K), f = (int)((g / a.K) % a.F), s = (int)(g / ((long long)a.K * a.F));
  const int K = a.K;
  constexpr int NE = 4 * LC + 1;
  double *blob = a.state + (((long long)s * a.F + f) * NE) * K + k;
  double wr[LC], wi[LC], br[LC], bi[LC];
#pragma unroll
  for (int e = 0; e < LC; ++e) {
    wr[e] = blob[(long long)e * K]; wi[e] = blob[(long long)(LC + e) * K];
    br[e] = blob[(long long)(2 * LC + e) * K]; bi[e] = blob[(long long)(3 * LC + e) * K];
  }
  double P = blob[(long long)(4 * LC) * K];
  const double om_alpha = 1.0 - a.alpha;
  for (int t = 0; t < a.T; ++t) {
    // tap shift: element index e = l * C + c, newest frame at l = 0            SubbandAF.py:47-49
#pragma unroll
    for (int e = LC - 1; e >= 0; --e) {
      if (e >= C) { br[e] = br[e - C]; bi[e] = bi[e - C]; }
    }
    const float2 *Xt = a.X + (((long long)s * a.T + t) * C) * K + k;
#pragma unroll
    for (int c = 0; c < C; ++c) { const float2 v = Xt[(long long)c * K]; br[c] = (double)v.x; bi[c] = (double)v.y; }
    const long long o = (((long long)s * a.T + t) * a.F + f) * K + k;
    const float2 dv = a.D[o];
    double pk = a.p ? a.p[((long long)s * a.T + t) * K + k] : 1.0;
    if (a.one_minus_p) pk = 1.0 - pk;
    double outr = 0.0, outi = 0.0, pw = 0.0;
#pragma unroll
    for (int e = 0; e < LC; ++e) {                    // conj(W) buf
      outr = fma(wr[e], br[e], fma(wi[e], bi[e], outr));
      outi = fma(wr[e], bi[e], fma(-wi[e], br[e], outi));
      pw = fma(br[e], br[e], fma(bi[e], bi[e], pw));
    }
    const double er = (double)dv.x - outr * pk, ei = (double)dv.y - outi * pk;
    P = a.alpha * P + om_alpha * pw / (double)C;
    const double step = 2.0 * a.mu * pk / (P + a.eps);
#pragma unroll
    for (int e = 0; e < LC; ++e) {                    // W += step * buf * conj(err)
      wr[e] = fma(step, fma(br[e], er, bi[e] * ei), wr[e]);
      wi[e] = fma(step, fma(bi[e], er, -br[e] * ei), wi[e]);
    }
    a.Err[o] = make_double2(er, ei);
  }
#pragma unroll
  for (int e = 0; e < LC; ++e) {
    blob[(long long)e * K] = wr[e]; blob[(long long)(LC + e) * K] = wi[e];
    blob[(long long)(2 * LC + e) * K] = br[e]; blob[(long long)(3 * LC + e) * K] = bi[e];
  }
  blob[(long long)(4 * LC) * K] = P;
}

// FilterDcNotch16.filter_dc_notch16 (adaptivefilter/feature.py:37-49), in place, one thread per (stream, channel);
// same operation order as dcnotch_kernel in fdgsc.cu, memories in a caller-owned [S][C][2] float64 array.
__global__ void dcnotch_generic_kernel(float *x, double *mem_all, int SC, int Ns, double radius, double den2) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= SC) return;
  float *xs = x + (size_t)g * Ns;
  double m0 = mem_all[2 * g], m1 = mem_all[2 * g + 1];
  for (int i = 0; i < Ns; ++i) {
    const double vin = (double)xs[i];
    const double vout = __dadd_rn(m0, vin);
    m0 = __dadd_rn(m1, __dmul_rn(2.0, __dadd_rn(-vin, __dmul_rn(radius, vout))));
    m1 = __dsub_rn(vin, __dmul_rn(den2, vout));
    xs[i] = (float)__dmul_rn(radius, vout);
  }
  mem_all[2 * g] = m0; mem_all[2 * g + 1] = m1;
}

// np.mean(x, axis=channels) of float64 [S][C][N] -> [S][N]: sequential sum in channel order, then / C
__global__ void channel_mean_kernel(const double *__restrict__ x, double *__restrict__ out, int S, int C, long long N) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * N) return;
  const long long s = g / N, n = g % N;
  const double *xs = x + (s * C) * N + n;
  double acc = xs[0];
  for (int c = 1; c < C; ++c) acc = __dadd_rn(acc, xs[(long long)c * N]);
  out[g] = acc / (double)C;
}

}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_subband_nlms_state_bytes(const ds_subband_nlms_params *p) {
  if (!p) return 0;
  return (size_t)p->n_streams * p->n_filters * nlms_state_elems(p->filter_len, p->n_ch) * p->n_bins * sizeof(double);
}

int ds_subband_nlms_run(const ds_subband_nlms_params *p, void *state, const void *X, const void *D, const double *prob,
                        void *Err, void *stream) {
  DS_CHECK_ARG(p && state && X && D && Err, "ds_subband_nlms_run: null argument");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_filters >= 1 && p->n_bins >= 1 && p->n_frames >= 1 && p->n_ch >= 1 && p->filter_len >= 1,
               "ds_subband_nlms_run: bad shape");
  NlmsArgs a;
  a.state = (double *)state; a.X = (const float2 *)X; a.D = (const float2 *)D; a.p = prob; a.Err = (double2 *)Err;
  a.S = p->n_streams; a.F = p->n_filters; a.K = p->n_bins; a.T = p->n_frames; a.C = p->n_ch; a.L = p->filter_len;
  a.one_minus_p = p->one_minus_p; a.mu = p->mu; a.alpha = p->alpha; a.eps = p->eps;
  const long long items = (long long)a.S * a.F * a.K;
  const unsigned blocks = (unsigned)((items + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
#define NLMS_CASE(LL, CC) \
  if (a.L == LL && a.C == CC) { subband_nlms_kernel<LL, CC><<<blocks, 128, 0, st>>>(a); launched = true; }
  bool launched = false;
  NLMS_CASE(1, 1) NLMS_CASE(2, 1) NLMS_CASE(3, 1) NLMS_CASE(4, 1) NLMS_CASE(1, 2) NLMS_CASE(2, 2) NLMS_CASE(1, 3) NLMS_CASE(2, 3)
  NLMS_CASE(1, 4) NLMS_CASE(2, 4) NLMS_CASE(3, 4) NLMS_CASE(4, 4) NLMS_CASE(1, 6) NLMS_CASE(2, 6) NLMS_CASE(1, 8) NLMS_CASE(2, 8)
#undef NLMS_CASE
  if (!launched) {
    set_error("ds_subband_nlms_run: filter_len %d x n_ch %d is not in the compiled set (taps 1..4 x {1,2,3,4} channels, 1..2 x {6,8})",
              a.L, a.C);
    return DS_EUNSUPPORTED;
  }
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int ds_dcnotch_run(int n_streams, int n_ch, int n_samples, double radius, double *mem, float *x, void *stream) {
  DS_CHECK_ARG(mem && x && n_streams >= 1 && n_ch >= 1 && n_samples >= 1, "ds_dcnotch_run: bad argument");
  const double den2 = radius * radius + 0.7 * (1 - radius) * (1 - radius);
  const int items = n_streams * n_ch;
  dcnotch_generic_kernel<<<(items + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, mem, items, n_samples, radius, den2);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int ds_channel_mean_run(int n_streams, int n_ch, long long n_samples, const double *x, double *out, void *stream) {
  DS_CHECK_ARG(x && out && n_streams >= 1 && n_ch >= 1 && n_samples >= 1, "ds_channel_mean_run: bad argument");
  const long long items = (long long)n_streams * n_samples;
  channel_mean_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, n_streams, n_ch, n_samples);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
