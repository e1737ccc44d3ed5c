// idoa.cu -- spatial speech-presence probability from the instantaneous-DOA feature:
// Idoa.estimate / Idoa.process, doa/idoa.py:92-209.
//
// Two kernels, because the reference's frame loop separates the same way:
//   idoa_rtf_kernel   thread per (stream, bin), frames sequential: the recursive RTF estimate
//                     B_hat = smooth(X_i conj(X_0)) / smooth(|X_0|^2) (:125-127) and its norm -- shared by every
//                     direction, so it is computed once and staged in HBM ([S][T][2(M-1)+1][K] float64);
//   idoa_spp_kernel   CTA per (stream, direction tile), thread per bin, frames sequential: similarity Delta to
//                     the free-field RTF of each direction (eq. 8, :130-136), its running statistics under H0 / Hd
//                     (:138-148), beta_n from the mean of mu_Delta over bins 72..127 -- the only coupling between
//                     bins, one shared-memory exchange + barrier per frame -- and the presence probability p
//                     (eq. 9-13, :150-163).  Psi, the statistics and p live in registers for the whole call.
//                     Directions never interact, so any subset of the grid can be run; Idoa.process needs one.
#include "common.cuh"
#include "perbin.cuh"

namespace ds {

constexpr int IDOA_TH = 2;          // directions per thread
constexpr int IDOA_MAXM = 8;
constexpr int IDOA_LO = 72, IDOA_HI = 128;      // mean(mu_Delta[72:128])          idoa.py:148
constexpr int IDOA_GLO = 64, IDOA_GHI = 128;    // mean(p[64:128, :, direction])   idoa.py:199

// state: [S][1 + 2(M-1)][K] float64 -- Y_smooth, then Re / Im of Y_xcorr_smooth per channel pair (zero = reset)
template <typename CT>
__global__ void idoa_rtf_kernel(int S, int T, int M, int K, double alpha, const CT *__restrict__ X, double *state,
                                double *__restrict__ Bout) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  const int D = M - 1, NE = 1 + 2 * D;
  double *st = state + (long long)s * NE * K + k;
  double ys = st[0], xr[IDOA_MAXM - 1], xi[IDOA_MAXM - 1];
  for (int i = 0; i < D; ++i) { xr[i] = st[(long long)(1 + 2 * i) * K]; xi[i] = st[(long long)(2 + 2 * i) * K]; }
  const double oma = 1.0 - alpha;
  for (int t = 0; t < T; ++t) {
    const CT *x = X + ((long long)s * T + t) * M * K + k;
    const double r0 = (double)x[0].x, i0 = (double)x[0].y;
    ys = __dadd_rn(__dmul_rn(oma, ys), __dmul_rn(alpha, power_c(r0, i0)));                    // :125
    const double scl = 1.0 / ys;              // numpy divides complex by real through the reciprocal
    double *b = Bout + ((long long)s * T + t) * NE * K + k;
    double n2 = 0.0;
    for (int i = 0; i < D; ++i) {
      const double ar = (double)x[(long long)(i + 1) * K].x, ai = (double)x[(long long)(i + 1) * K].y;
      const double cr = __dadd_rn(__dmul_rn(ar, r0), __dmul_rn(ai, i0));                      // X_i conj(X_0)
      const double ci = __dsub_rn(__dmul_rn(ai, r0), __dmul_rn(ar, i0));
      xr[i] = __dadd_rn(__dmul_rn(oma, xr[i]), __dmul_rn(alpha, cr));                          // :126
      xi[i] = __dadd_rn(__dmul_rn(oma, xi[i]), __dmul_rn(alpha, ci));
      const double br = xr[i] * scl, bi = xi[i] * scl;                                         // :127
      b[(long long)(2 * i) * K] = br; b[(long long)(2 * i + 1) * K] = bi;
      n2 += br * br + bi * bi;
    }
    b[(long long)(2 * D) * K] = sqrt(n2);
  }
  st[0] = ys;
  for (int i = 0; i < D; ++i) { st[(long long)(1 + 2 * i) * K] = xr[i]; st[(long long)(2 + 2 * i) * K] = xi[i]; }
}

struct IdoaSppArgs {
  int S, T, M, K, n_slots, only_theta, want_gain;
  const int *theta;          // [n_slots] direction index of every state slot
  const double2 *Psi;        // [n_theta][M-1][K] free-field RTFs
  const double *B;           // [S][T][2(M-1)+1][K] from idoa_rtf_kernel
  double *state;             // [S][n_slots][4][K]: mu_Delta, mu_Delta_h0, var_Delta_h0 - 0.1, p   (zero = reset)
  double *p_out;             // [S][T][n_slots][K] or null
  const void *X;             // [S][T][M][K] (gain mode: Y = g X_0)
  int x_is_c128;
  double2 *Yout;             // [S][T][K] or null
};

// grid: (ceil(n_slots / IDOA_TH), S); block: K rounded up to a warp multiple.  M is a template parameter so that
// Psi, B and the statistics are indexed with compile-time constants (registers, not local memory).
template <int M>
__global__ void __launch_bounds__(288) idoa_spp_kernel(IdoaSppArgs a) {
  __shared__ double sm_mu[2][IDOA_TH][IDOA_HI - IDOA_LO];
  __shared__ double sm_p[2][IDOA_GHI - IDOA_GLO];
  constexpr int D = M - 1, NE = 1 + 2 * D;
  const int k = threadIdx.x, s = blockIdx.y, K = a.K;
  const bool live = k < K;
  const int slot0 = blockIdx.x * IDOA_TH;
  const int nth = min(IDOA_TH, a.n_slots - slot0);
  double pr[IDOA_TH][D], pi[IDOA_TH][D], npsi[IDOA_TH];
  double mu[IDOA_TH], mu0[IDOA_TH], var0[IDOA_TH], p[IDOA_TH];
  bool zero_delta[IDOA_TH];
#pragma unroll
  for (int j = 0; j < IDOA_TH; ++j) {
    mu[j] = mu0[j] = p[j] = 0.0; var0[j] = 0.1; npsi[j] = 0.0; zero_delta[j] = true;
    if (j < nth && live) {
      const int th = a.theta[slot0 + j];
      zero_delta[j] = a.only_theta >= 0 && th != a.only_theta;                 // estimate(X, theta=int): the other columns see Delta = 0
      double n2 = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double2 v = a.Psi[((long long)th * D + i) * K + k];
        pr[j][i] = v.x; pi[j][i] = v.y;
        n2 += v.x * v.x + v.y * v.y;
      }
      npsi[j] = sqrt(n2);                                                     // np.linalg.norm(Psi[:, :, theta], axis=-1)
      const double *st = a.state + (((long long)s * a.n_slots + slot0 + j) * 4) * K + k;
      mu[j] = st[0]; mu0[j] = st[K]; var0[j] = st[2 * (long long)K] + 0.1; p[j] = st[3 * (long long)K];
    }
  }
  for (int t = 0; t < a.T; ++t) {
    const int buf = t & 1;
    double delta[IDOA_TH];
    if (live) {
      const double *b = a.B + ((long long)s * a.T + t) * NE * K + k;
      double br[D], bi[D];
#pragma unroll
      for (int i = 0; i < D; ++i) { br[i] = b[(long long)(2 * i) * K]; bi[i] = b[(long long)(2 * i + 1) * K]; }
      const double nB = b[(long long)(2 * D) * K];
#pragma unroll
      for (int j = 0; j < IDOA_TH; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) acc += pr[j][i] * br[i] + pi[j][i] * bi[i];               // Re(conj(Psi) B)
        delta[j] = zero_delta[j] ? 0.0 : acc / (npsi[j] * nB + 1e-6);                          // eq. 8
        double avg = (1.0 - p[j]) * 0.98;
        mu[j] = avg * mu[j] + (1.0 - avg) * delta[j];                                           // eq. 15
        avg = 0.998 + (1.0 - 0.998) * p[j];
        mu0[j] = avg * mu0[j] + (1.0 - avg) * delta[j];
        const double dd = delta[j] - mu0[j];
        var0[j] = fmax((1.0 - avg) * var0[j] + avg * (dd * dd), 0.01);
        if (k >= IDOA_LO && k < IDOA_HI) sm_mu[buf][j][k - IDOA_LO] = mu[j];
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int j = 0; j < IDOA_TH; ++j) {
        double sum = 0.0;                      // np.mean over axis 0: rows added in order
        for (int q = 0; q < IDOA_HI - IDOA_LO; ++q) sum += sm_mu[buf][j][q];
        const double beta_n = 1.0 / (1.0 - sum / (double)(IDOA_HI - IDOA_LO));
        const double dd = delta[j] - mu0[j];
        const double p_h0 = exp(-(dd * dd) / (2.0 * 0.5 * 0.5));                               // eq. 11
        const double p_hd = beta_n * exp(7.6 * (delta[j] - 1.0));                              // eq. 9
        const double lam = p_hd / (p_h0 + 1e-6);                                               // eq. 13
        p[j] = lam / (1.0 + lam);                                                              // eq. 12
        if (a.p_out && j < nth) a.p_out[(((long long)s * a.T + t) * a.n_slots + slot0 + j) * K + k] = p[j];
      }
    }
    if (a.want_gain) {                         // Idoa.process: Y = max(mean(p[64:128]), 0.01) X_0   (:199)
      if (live && k >= IDOA_GLO && k < IDOA_GHI) sm_p[buf][k - IDOA_GLO] = p[0];
      __syncthreads();
      if (live) {
        double sum = 0.0;
        for (int q = 0; q < IDOA_GHI - IDOA_GLO; ++q) sum += sm_p[buf][q];
        const double gain = fmax(sum / (double)(IDOA_GHI - IDOA_GLO), 0.01);
        const long long xo = ((long long)s * a.T + t) * M * K + k;
        double xr, xi;
        if (a.x_is_c128) { const double2 v = ((const double2 *)a.X)[xo]; xr = v.x; xi = v.y; }
        else { const float2 v = ((const float2 *)a.X)[xo]; xr = (double)v.x; xi = (double)v.y; }
        a.Yout[((long long)s * a.T + t) * K + k] = make_double2(gain * xr, gain * xi);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int j = 0; j < IDOA_TH; ++j)
      if (j < nth) {
        double *st = a.state + (((long long)s * a.n_slots + slot0 + j) * 4) * K + k;
        st[0] = mu[j]; st[K] = mu0[j]; st[2 * (long long)K] = var0[j] - 0.1; st[3 * (long long)K] = p[j];
      }
  }
}

}  // namespace ds

using namespace ds;

extern "C" size_t ds_idoa_rtf_state_bytes(int n_streams, int n_mics, int n_bins) {
  return (size_t)n_streams * (1 + 2 * (n_mics - 1)) * n_bins * sizeof(double);
}
extern "C" size_t ds_idoa_spp_state_bytes(int n_streams, int n_slots, int n_bins) {
  return (size_t)n_streams * n_slots * 4 * n_bins * sizeof(double);
}

extern "C" int ds_idoa_rtf_run(int n_streams, int n_frames, int n_mics, int n_bins, double alpha, const void *X,
                               int x_is_c128, void *state, double *B, void *stream) {
  DS_CHECK_ARG(X && state && B, "idoa_rtf: null pointer");
  DS_CHECK_ARG(n_streams > 0 && n_frames > 0 && n_bins > 0, "idoa_rtf: empty shape");
  if (n_mics < 2 || n_mics > IDOA_MAXM) { set_error("idoa: n_mics %d outside 2..%d", n_mics, IDOA_MAXM); return DS_EUNSUPPORTED; }
  const long long n = (long long)n_streams * n_bins;
  const unsigned grid = (unsigned)((n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_is_c128) idoa_rtf_kernel<double2><<<grid, 128, 0, st>>>(n_streams, n_frames, n_mics, n_bins, alpha, (const double2 *)X, (double *)state, B);
  else idoa_rtf_kernel<float2><<<grid, 128, 0, st>>>(n_streams, n_frames, n_mics, n_bins, alpha, (const float2 *)X, (double *)state, B);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_idoa_spp_run(int n_streams, int n_frames, int n_mics, int n_bins, int n_slots, const int *theta,
                               int only_theta, const void *Psi, const double *B, void *state, double *p_out,
                               const void *X, int x_is_c128, void *Yout, void *stream) {
  DS_CHECK_ARG(theta && Psi && B && state, "idoa_spp: null pointer");
  DS_CHECK_ARG(n_streams > 0 && n_frames > 0 && n_slots > 0, "idoa_spp: empty shape");
  DS_CHECK_ARG(n_bins >= IDOA_HI && n_bins <= 288, "idoa_spp: n_bins %d outside %d..288 (the reference averages bins 72..127)", n_bins, IDOA_HI);
  DS_CHECK_ARG(!Yout || (X && n_slots == 1), "idoa_spp: the gain output needs X and exactly one direction");
  if (n_mics < 2 || n_mics > IDOA_MAXM) { set_error("idoa: n_mics %d outside 2..%d", n_mics, IDOA_MAXM); return DS_EUNSUPPORTED; }
  IdoaSppArgs a;
  a.S = n_streams; a.T = n_frames; a.M = n_mics; a.K = n_bins; a.n_slots = n_slots; a.only_theta = only_theta;
  a.want_gain = Yout != nullptr; a.theta = theta; a.Psi = (const double2 *)Psi; a.B = B; a.state = (double *)state;
  a.p_out = p_out; a.X = X; a.x_is_c128 = x_is_c128; a.Yout = (double2 *)Yout;
  const dim3 grid((n_slots + IDOA_TH - 1) / IDOA_TH, n_streams);
  const int nt = ((n_bins + 31) / 32) * 32;
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_mics) {
    case 2: idoa_spp_kernel<2><<<grid, nt, 0, st>>>(a); break;
    case 3: idoa_spp_kernel<3><<<grid, nt, 0, st>>>(a); break;
    case 4: idoa_spp_kernel<4><<<grid, nt, 0, st>>>(a); break;
    case 5: idoa_spp_kernel<5><<<grid, nt, 0, st>>>(a); break;
    case 6: idoa_spp_kernel<6><<<grid, nt, 0, st>>>(a); break;
    case 7: idoa_spp_kernel<7><<<grid, nt, 0, st>>>(a); break;
    default: idoa_spp_kernel<8><<<grid, nt, 0, st>>>(a); break;
  }
  DS_LAUNCH_CHECK();
  return DS_OK;
}
