// gsc.cu -- McMcra (noise_estimation/mc_mcra.py:179-221) and the frequency-domain GSC that uses it as
// presence detector and postfilter (beamformer/GSC.py:174-294).  One thread per (stream, bin), frames
// sequential, all recursive state float64 like the reference.
//
// Per frame and bin (file:line of the reference):
//   McMcra.estimation                                                         mc_mcra.py
//     Phi_yy = .92 Phi_yy + .08 Re(conj(y)^T y)                               :182-184   (real parts only)
//     first 5 frames: Phi_vv = Phi_yy                                         :186-187
//     A = inv(Phi_vv + 1e-6 I);  psi~ = tr(A Phi_yy);  xi = clip(psi~ - M)    :191-194
//     gamma = clip(Re(conj(y) A (Phi_yy - Phi_vv) A y^T))                     :196-199
//     psi = Re(y A conj(y)^T);  q from the thresholds of compute_q_local     :89-103   (psi_0 = psi~_0 = 100)
//     p = clip(1 / (1 + q/(1-q) (1+xi) exp(-gamma/(1+xi))), .01, .99)         :142-150
//     Phi_vv = Re(at Phi_vv + (1 - at) conj(y)^T y), at = .95 + .05 p         :207-219
//     G = clip((xi/(1+xi))^p Gmin^(1-p), Gmin, 1), G[:2] = 0                  :152-156
//   GSC.process (method != 0)                                                 GSC.py
//     U_i = conj(a_0) Z_0 - conj(a_{i+1}) Z_{i+1}      (BM^H Z)                :220-225, :255
//     Yfbf = sum_m conj(W_m) Z_m,  W = a / (a^H a)                             :219, :257
//     Y = Yfbf - sum_i conj(G_i) U_i;  G_i += mu (1 - p) U_i conj(Y)           :260-271   (mu = .01, Pest = 1)
//     output spectrum Y * G_postfilter                                         :286
#include "chain_step.cuh"

namespace ds {

struct GscArgs {
  double *state;            // [S][NE][K]
  const double2 *a;         // [M][K] propagation vectors exp(-j w_k tau_m), or null: McMcra only
  const void *X;            // [S][T][M][K] c64 / c128
  int x_c128;
  float2 *Yout;             // [S][T][K] or null
  double *tp, *tG, *txi, *tgamma, *tq;      // taps [S][T][K] or null
  int S, K, T, frm_cnt, method, init_frames;
  double alpha, alpha_d, eps, psi0, q_min, q_max, p_min, p_max, snr_min, snr_max, Gmin, logGmin, mu;
};

// state: Phi_yy[NP] Phi_vv[NP] (packed real upper triangles), Gw_re[M-1] Gw_im[M-1], then the scalars of the
// last frame: p q xi gamma G
template <int M> __host__ __device__ constexpr int gsc_state_elems() { return M * (M + 1) + 2 * (M - 1) + 5; }

template <int M, int NT>
__global__ void __launch_bounds__(NT) gsc_kernel(GscArgs a) {
  constexpr int NP = M * (M + 1) / 2, NE = gsc_state_elems<M>();
  constexpr int OFF_G = 2 * NP, OFF_S = 2 * NP + 2 * (M - 1);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x, K = a.K;
  const long long g = (long long)blockIdx.x * NT + tid;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *blob = a.state + (long long)s * NE * K + k;
  double *smy = sm + tid, *smv = sm + NP * NT + tid;
#pragma unroll
  for (int e = 0; e < NP; ++e) { smy[e * NT] = blob[(long long)e * K]; smv[e * NT] = blob[(long long)(NP + e) * K]; }
  double gr[M - 1], gi[M - 1];
#pragma unroll
  for (int i = 0; i < M - 1; ++i) { gr[i] = blob[(long long)(OFF_G + i) * K]; gi[i] = blob[(long long)(OFF_G + M - 1 + i) * K]; }
  double p = 0.0, q = 0.0, xi = 0.0, gamma = 0.0, G = 0.0;
  int frm = a.frm_cnt;
  const double one_m_alpha = 1.0 - a.alpha;

  startup_dephase(8000);
  for (int t = 0; t < a.T; ++t, ++frm) {
    double yr[M], yi[M];
    {
      const long long base = ((long long)s * a.T + t) * M * K + k;
      if (a.x_c128) {
        const double2 *X = reinterpret_cast<const double2 *>(a.X) + base;
#pragma unroll
        for (int m = 0; m < M; ++m) { const double2 v = X[(long long)m * K]; yr[m] = v.x; yi[m] = v.y; }
      } else {
        const float2 *X = reinterpret_cast<const float2 *>(a.X) + base;
#pragma unroll
        for (int m = 0; m < M; ++m) { const float2 v = X[(long long)m * K]; yr[m] = (double)v.x; yi[m] = (double)v.y; }
      }
    }
    // ---- McMcra ----------------------------------------------------------------------------
    // Phi_yy update; the first frames copy it into Phi_vv
    const bool init = frm < a.init_frames;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double pyy = fma(one_m_alpha, fma(yr[i], yr[j], yi[i] * yi[j]), a.alpha * smy[e * NT]);
        smy[e * NT] = pyy;
        if (init) smv[e * NT] = pyy;
      }
    double A[NP];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) A[pidx<M>(i, j)] = smv[pidx<M>(i, j) * NT] + ((i == j) ? a.eps : 0.0);
    spd_inverse_packed<M>(A);
#define AS(i, j) (((i) <= (j)) ? A[pidx<M>(i, j)] : A[pidx<M>(j, i)])
    double ur[M], ui[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) { sr = fma(AS(i, j), yr[j], sr); si = fma(AS(i, j), yi[j], si); }
      ur[i] = sr; ui[i] = si;
    }
#undef AS
    double psi = 0.0;                       // Re(y A conj(y)^T)
#pragma unroll
    for (int i = 0; i < M; ++i) psi = fma(yr[i], ur[i], fma(yi[i], ui[i], psi));
    double psit = 0.0, gam = 0.0;           // tr(A Phi_yy), Re(u^H (Phi_yy - Phi_vv) u)
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double w = (i == j) ? 1.0 : 2.0;
        const double pyy = smy[e * NT];
        const double x = pyy - smv[e * NT];
        psit = fma(w * A[e], pyy, psit);
        gam = fma(w * x, fma(ui[i], ui[j], ur[i] * ur[j]), gam);
      }
    xi = fmin(fmax(psit - (double)M, a.snr_min), a.snr_max);
    gamma = fmin(fmax(gam, a.snr_min), a.snr_max);
    if (psi >= a.psi0 || psit > a.psi0) q = a.q_min;                 // compute_q_local, mc_mcra.py:95-103
    else if (psit < (double)M) q = a.q_max;
    else q = fmin(fmax((a.psi0 - psit) / (a.psi0 - (double)M), a.q_min), a.q_max);
    p = 1.0 / (1.0 + q / (1.0 - q) * (1.0 + xi) * exp(-1.0 * (gamma / (1.0 + xi))));
    p = fmin(fmax(p, a.p_min), a.p_max);
    const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
    const double om_at = 1.0 * (1.0 - at);
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        smv[e * NT] = fma(om_at, fma(yr[i], yr[j], yi[i] * yi[j]), at * smv[e * NT]);
      }
    G = pow(xi / (1.0 + xi), p) * exp((1.0 - p) * a.logGmin);
    G = fmax(fmin(G, 1.0), a.Gmin);
    if (k < 2) G = 0.0;
    const long long o = ((long long)s * a.T + t) * K + k;
    if (a.tp) a.tp[o] = p;
    if (a.tG) a.tG[o] = G;
    if (a.txi) a.txi[o] = xi;
    if (a.tgamma) a.tgamma[o] = gamma;
    if (a.tq) a.tq[o] = q;
    // ---- GSC ---------------------------------------------------------------------------------
    if (a.a && a.Yout) {
      double Yr, Yi;
      if (a.method == 0) {
        Yr = yr[0]; Yi = yi[0];                                       // output channel_1 (GSC.py:242)
      } else {
        double ar[M], ai[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { const double2 v = a.a[(long long)m * K + k]; ar[m] = v.x; ai[m] = v.y; }
        double nrm = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) nrm = fma(ar[m], ar[m], fma(ai[m], ai[m], nrm));
        // fixed beam: sum conj(a_m) Z_m / (a^H a)
        double fr = 0.0, fi = 0.0;
        double cr[M], ci[M];                                          // conj(a_m) Z_m
#pragma unroll
        for (int m = 0; m < M; ++m) {
          cr[m] = fma(ar[m], yr[m], ai[m] * yi[m]);
          ci[m] = fma(ar[m], yi[m], -ai[m] * yr[m]);
          fr += cr[m]; fi += ci[m];
        }
        Yr = fr / nrm; Yi = fi / nrm;
        double Ur[M - 1], Ui[M - 1];
#pragma unroll
        for (int i = 0; i < M - 1; ++i) {
          Ur[i] = cr[0] - cr[i + 1]; Ui[i] = ci[0] - ci[i + 1];      // BM^H Z
          Yr -= fma(gr[i], Ur[i], gi[i] * Ui[i]);                     // conj(G_i) U_i
          Yi -= fma(gr[i], Ui[i], -gi[i] * Ur[i]);
        }
        const double step = a.mu * (1.0 - p);
#pragma unroll
        for (int i = 0; i < M - 1; ++i) {                             // G_i += mu (1-p) U_i conj(Y)
          gr[i] = fma(step, fma(Ur[i], Yr, Ui[i] * Yi), gr[i]);
          gi[i] = fma(step, fma(Ui[i], Yr, -Ur[i] * Yi), gi[i]);
        }
        Yr *= G; Yi *= G;                                             // postfilter (:286)
      }
      a.Yout[o] = make_float2((float)Yr, (float)Yi);
    }
  }
#pragma unroll
  for (int e = 0; e < NP; ++e) { blob[(long long)e * K] = smy[e * NT]; blob[(long long)(NP + e) * K] = smv[e * NT]; }
#pragma unroll
  for (int i = 0; i < M - 1; ++i) { blob[(long long)(OFF_G + i) * K] = gr[i]; blob[(long long)(OFF_G + M - 1 + i) * K] = gi[i]; }
  blob[(long long)(OFF_S + 0) * K] = p; blob[(long long)(OFF_S + 1) * K] = q; blob[(long long)(OFF_S + 2) * K] = xi;
  blob[(long long)(OFF_S + 3) * K] = gamma; blob[(long long)(OFF_S + 4) * K] = G;
}

template <int M> static int launch_gsc(const GscArgs &a, cudaStream_t st) {
  constexpr int NT = 64, NP = M * (M + 1) / 2;
  const size_t smem = (size_t)2 * NP * NT * sizeof(double);
  auto kern = gsc_kernel<M, NT>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)a.S * a.K;
  kern<<<(unsigned)((items + NT - 1) / NT), NT, smem, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // namespace ds

using namespace ds;

extern "C" {

void ds_gsc_default_params(ds_gsc_params *p, int n_fft, int n_streams, int n_mics, int n_frames) {
  if (!p) return;
  p->n_fft = n_fft; p->n_streams = n_streams; p->n_mics = n_mics; p->n_frames = n_frames;
  p->frm_cnt = 0; p->method = 2; p->init_frames = 5; p->reserved = 0;
  p->alpha = 0.92; p->alpha_d = 0.95; p->diag_eps = 1e-6; p->psi_0 = 100.0; p->q_min = 0.01; p->q_max = 0.99;
  p->p_min = 0.01; p->p_max = 0.99; p->snr_min = 1e-6; p->snr_max = 1e6; p->Gmin = 0.0631; p->mu = 0.01;
}

size_t ds_gsc_state_bytes(const ds_gsc_params *p) {
  if (!p || p->n_mics < 2) return 0;
  const int M = p->n_mics;
  return (size_t)p->n_streams * (M * (M + 1) + 2 * (M - 1) + 5) * (p->n_fft / 2 + 1) * sizeof(double);
}

int ds_gsc_run(const ds_gsc_params *p, void *state, const void *a, const void *X, int x_is_c128, void *Yout,
               const ds_gsc_taps *taps, void *stream) {
  DS_CHECK_ARG(p && state && X, "ds_gsc_run: null argument");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_frames >= 1 && p->n_fft >= 64, "ds_gsc_run: bad shape");
  DS_CHECK_ARG(!Yout || a, "ds_gsc_run: the GSC output needs the propagation vectors");
  GscArgs g;
  g.state = (double *)state; g.a = (const double2 *)a; g.X = X; g.x_c128 = x_is_c128; g.Yout = (float2 *)Yout;
  g.tp = taps ? taps->p : nullptr; g.tG = taps ? taps->G : nullptr; g.txi = taps ? taps->xi : nullptr;
  g.tgamma = taps ? taps->gamma : nullptr; g.tq = taps ? taps->q : nullptr;
  g.S = p->n_streams; g.K = p->n_fft / 2 + 1; g.T = p->n_frames; g.frm_cnt = p->frm_cnt; g.method = p->method;
  g.init_frames = p->init_frames;
  g.alpha = p->alpha; g.alpha_d = p->alpha_d; g.eps = p->diag_eps; g.psi0 = p->psi_0; g.q_min = p->q_min; g.q_max = p->q_max;
  g.p_min = p->p_min; g.p_max = p->p_max; g.snr_min = p->snr_min; g.snr_max = p->snr_max; g.Gmin = p->Gmin;
  g.logGmin = log(p->Gmin); g.mu = p->mu;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->n_mics) {
    case 2: return launch_gsc<2>(g, st);
    case 3: return launch_gsc<3>(g, st);
    case 4: return launch_gsc<4>(g, st);
    case 5: return launch_gsc<5>(g, st);
    case 6: return launch_gsc<6>(g, st);
    case 7: return launch_gsc<7>(g, st);
    case 8: return launch_gsc<8>(g, st);
  }
  set_error("ds_gsc_run: n_mics %d outside the compiled range 2..8", p->n_mics);
  return DS_EUNSUPPORTED;
}

}  // extern "C"
