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
  int S, F, K, T, C, L, one_minus_p, plain_lms;
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
    double step = 2.0 * a.mu * pk;
    if (!a.plain_lms) {                               // normalization=True (default)          SubbandLMS.py:69-75
      P = a.alpha * P + om_alpha * pw / (double)C;
      step /= (P + a.eps);
    }
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

// SubbandRLS.update (adaptivefilter/SubbandRLS.py:44-71): per-bin RLS, L frame taps, one input channel.
//   err = D - conj(W) buf;  k = P buf / (lambda + buf^H P buf);  P = (P - k buf^H P) / lambda;  W += 2 mu conj(err) k
// state per (stream, bin): W re/im [L], buf re/im [L], P re/im [L*L] (row major); P(0) = I / 1e-3 is written by the
// host when the state is created.
struct RlsArgs {
  double *state;            // [S][NE][K]
  const float2 *X;          // [S][T][K]
  const float2 *D;          // [S][T][K]
  double2 *Err;             // [S][T][K]
  int S, K, T;
  double mu, lambda;
};

template <int L>
__global__ void subband_rls_kernel(RlsArgs a) {
  constexpr int NE = 4 * L + 2 * L * L;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * a.K) return;
  const int k = (int)(g % a.K), s = (int)(g / a.K), K = a.K;
  double *blob = a.state + ((long long)s * NE) * K + k;
  double wr[L], wi[L], br[L], bi[L], Pr[L][L], Pi[L][L];
#pragma unroll
  for (int e = 0; e < L; ++e) {
    wr[e] = blob[(long long)e * K]; wi[e] = blob[(long long)(L + e) * K];
    br[e] = blob[(long long)(2 * L + e) * K]; bi[e] = blob[(long long)(3 * L + e) * K];
  }
#pragma unroll
  for (int i = 0; i < L; ++i)
#pragma unroll
    for (int j = 0; j < L; ++j) {
      Pr[i][j] = blob[(long long)(4 * L + i * L + j) * K];
      Pi[i][j] = blob[(long long)(4 * L + L * L + i * L + j) * K];
    }
  const double inv_lambda = 1.0 / a.lambda;
  for (int t = 0; t < a.T; ++t) {
#pragma unroll
    for (int e = L - 1; e >= 1; --e) { br[e] = br[e - 1]; bi[e] = bi[e - 1]; }
    const long long o = ((long long)s * a.T + t) * K + k;
    { const float2 v = a.X[o]; br[0] = (double)v.x; bi[0] = (double)v.y; }
    const float2 dv = a.D[o];
    double outr = 0.0, outi = 0.0;
#pragma unroll
    for (int e = 0; e < L; ++e) {
      outr = fma(wr[e], br[e], fma(wi[e], bi[e], outr));
      outi = fma(wr[e], bi[e], fma(-wi[e], br[e], outi));
    }
    const double er = (double)dv.x - outr, ei = (double)dv.y - outi;
    // num = P buf;  den = lambda + buf^H num (complex);  k = num / den
    double nr[L], ni[L], denr = a.lambda, deni = 0.0;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        sr = fma(Pr[i][j], br[j], fma(-Pi[i][j], bi[j], sr));
        si = fma(Pr[i][j], bi[j], fma(Pi[i][j], br[j], si));
      }
      nr[i] = sr; ni[i] = si;
      denr = fma(br[i], sr, fma(bi[i], si, denr));        // conj(buf_i) num_i
      deni = fma(br[i], si, fma(-bi[i], sr, deni));
    }
    const double d2 = 1.0 / fma(denr, denr, deni * deni);
    double kr[L], ki[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {                          // num / den
      kr[i] = fma(nr[i], denr, ni[i] * deni) * d2;
      ki[i] = fma(ni[i], denr, -nr[i] * deni) * d2;
    }
    // v = buf^H P (row vector);  P = (P - k v) / lambda
    double vr[L], vi[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int i = 0; i < L; ++i) {
        sr = fma(br[i], Pr[i][j], fma(bi[i], Pi[i][j], sr));
        si = fma(br[i], Pi[i][j], fma(-bi[i], Pr[i][j], si));
      }
      vr[j] = sr; vi[j] = si;
    }
#pragma unroll
    for (int i = 0; i < L; ++i)
#pragma unroll
      for (int j = 0; j < L; ++j) {
        Pr[i][j] = (Pr[i][j] - fma(kr[i], vr[j], -ki[i] * vi[j])) * inv_lambda;
        Pi[i][j] = (Pi[i][j] - fma(kr[i], vi[j], ki[i] * vr[j])) * inv_lambda;
      }
    const double two_mu = 2.0 * a.mu;
#pragma unroll
    for (int e = 0; e < L; ++e) {                          // W += 2 mu conj(err) k
      wr[e] = fma(two_mu, fma(er, kr[e], ei * ki[e]), wr[e]);
      wi[e] = fma(two_mu, fma(er, ki[e], -ei * kr[e]), wi[e]);
    }
    a.Err[o] = make_double2(er, ei);
  }
#pragma unroll
  for (int e = 0; e < L; ++e) {
    blob[(long long)e * K] = wr[e]; blob[(long long)(L + e) * K] = wi[e];
    blob[(long long)(2 * L + e) * K] = br[e]; blob[(long long)(3 * L + e) * K] = bi[e];
  }
#pragma unroll
  for (int i = 0; i < L; ++i)
#pragma unroll
    for (int j = 0; j < L; ++j) {
      blob[(long long)(4 * L + i * L + j) * K] = Pr[i][j];
      blob[(long long)(4 * L + L * L + i * L + j) * K] = Pi[i][j];
    }
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
  a.one_minus_p = p->one_minus_p; a.plain_lms = p->plain_lms; a.mu = p->mu; a.alpha = p->alpha; a.eps = p->eps;
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

size_t ds_subband_rls_state_bytes(int n_streams, int n_bins, int filter_len) {
  return (size_t)n_streams * (4 * filter_len + 2 * filter_len * filter_len) * n_bins * sizeof(double);
}

int ds_subband_rls_run(int n_streams, int n_bins, int n_frames, int filter_len, double mu, double forgetting_factor, void *state,
                       const void *X, const void *D, void *Err, void *stream) {
  DS_CHECK_ARG(state && X && D && Err, "ds_subband_rls_run: null argument");
  DS_CHECK_ARG(n_streams >= 1 && n_bins >= 1 && n_frames >= 1 && forgetting_factor > 0.0, "ds_subband_rls_run: bad shape");
  RlsArgs a;
  a.state = (double *)state; a.X = (const float2 *)X; a.D = (const float2 *)D; a.Err = (double2 *)Err;
  a.S = n_streams; a.K = n_bins; a.T = n_frames; a.mu = mu; a.lambda = forgetting_factor;
  const long long items = (long long)a.S * a.K;
  const unsigned blocks = (unsigned)((items + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  switch (filter_len) {
    case 1: subband_rls_kernel<1><<<blocks, 128, 0, st>>>(a); break;
    case 2: subband_rls_kernel<2><<<blocks, 128, 0, st>>>(a); break;
    case 3: subband_rls_kernel<3><<<blocks, 128, 0, st>>>(a); break;
    case 4: subband_rls_kernel<4><<<blocks, 128, 0, st>>>(a); break;
    default: set_error("ds_subband_rls_run: filter_len %d outside the compiled range 1..4", filter_len); return DS_EUNSUPPORTED;
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
