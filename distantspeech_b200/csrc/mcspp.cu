// mcspp.cu -- MCRA and McSppBase estimators, MVDR weights, OMLSA gain and the
// weight/gain apply, one thread per (stream, frequency bin), frames sequential.
//
// Reference behaviour restated (file:line relative to the reference tree):
//   NoiseEstimationMCRA.estimation   noise_estimation/mcra.py:27-77
//   McSppBase.estimation             noise_estimation/mcspp_base.py:262-297
//   McSppBase.compute_omlsa_weight   mcspp_base.py:140-155
//   McSppBase.compute_pmwf_weight    mcspp_base.py:220-240
//   compute_mvdr_weight              beamformer/beamformer.py:133-155
//
// All recursive state is float64 like the reference (fp32 state collapses the
// SPP feedback loop to ~21 dB).  During a launch the covariance
// state of a thread lives in shared memory ([element][thread], conflict free)
// so HBM traffic is one read of X and one write of Y per frame; the blob in
// HBM is only touched at the start and the end of the launch.
#include "common.cuh"
#include "perbin.cuh"
#include "mcspp_args.cuh"

namespace ds {

// ===========================================================================
// MCRA standalone
// ===========================================================================
struct McraArgs {
  double *state; const double *Y; double *lam_out; double *p_out;
  int S, K, T, frm_cnt, ell;
  McraConst c;
};

__global__ void mcra_kernel(McraArgs a) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * a.K) return;
  const int s = (int)(g / a.K), k = (int)(g % a.K);
  double *st = a.state + (long long)s * 5 * a.K + k;
  double S = st[0], Smin = st[(long long)a.K], Stmp = st[2LL * a.K], p = st[3LL * a.K], lam = st[4LL * a.K];
  int frm = a.frm_cnt, ell = a.ell % a.c.L;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
  for (int t = 0; t < a.T; ++t) {
    const double *y = a.Y + ((long long)s * a.T + t) * a.K;
    const double Y0 = y[k];
    const double Ym1 = (k > 0) ? y[k - 1] : 0.0;
    const double Yp1 = (k < a.K - 1) ? y[k + 1] : 0.0;
    const bool reset = (frm > 0) && (ell == 0);
    mcra_step(S, Smin, Stmp, p, lam, Ym1, Y0, Yp1, k, a.K, frm, reset, a.c);
    if (reset) ell = 0;
    ++ell; ++frm;
    if (ell == a.c.L) ell = 0;
    const long long o = ((long long)s * a.T + t) * a.K + k;
    if (a.lam_out) a.lam_out[o] = lam;
    if (a.p_out) a.p_out[o] = p;
  }
  st[0] = S; st[(long long)a.K] = Smin; st[2LL * a.K] = Stmp; st[3LL * a.K] = p; st[4LL * a.K] = lam;
}

// ===========================================================================
// McSppBase + MVDR + OMLSA
// ===========================================================================
// McsppArgs: see mcspp_args.cuh


template <typename XT> struct XLoad;
template <> struct XLoad<float2> {
  __device__ static __forceinline__ void ld(const void *p, long long i, double &re, double &im) {
    float2 v = reinterpret_cast<const float2 *>(p)[i]; re = (double)v.x; im = (double)v.y;
  }
};
template <> struct XLoad<double2> {
  __device__ static __forceinline__ void ld(const void *p, long long i, double &re, double &im) {
    double2 v = reinterpret_cast<const double2 *>(p)[i]; re = v.x; im = v.y;
  }
};

template <int M, bool FULL, typename XT, int NT>
__global__ void __launch_bounds__(NT) mcspp_kernel(McsppArgs a) {
  constexpr int NP = M * (M + 1) / 2, NQ = M * (M - 1) / 2;
  constexpr int NE = mcspp_state_elems<M>();
  // blob element order: PyyR[NP] PvvR[NP] mcra[5] PyyI[NQ] PvvI[NQ]
  constexpr int OFF_YR = 0, OFF_VR = NP, OFF_MC = 2 * NP, OFF_YI = 2 * NP + 5, OFF_VI = 2 * NP + 5 + NQ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  const int Kp = a.K - a.k_first;
  const long long g = (long long)blockIdx.x * NT + tid;
  if (g >= (long long)a.S * Kp) return;     // no block-wide sync below: threads are independent
  const int s = (int)(g / Kp), k = a.k_first + (int)(g % Kp);
  const int K = a.K;
  double *blob = a.state + (long long)s * NE * K + k;
#define SM_YR(e) sm[(e) * NT + tid]
#define SM_VR(e) sm[(NP + (e)) * NT + tid]
#define SM_YI(e) sm[(2 * NP + (e)) * NT + tid]
#define SM_VI(e) sm[(2 * NP + NQ + (e)) * NT + tid]
#pragma unroll
  for (int e = 0; e < NP; ++e) { SM_YR(e) = blob[(long long)(OFF_YR + e) * K]; SM_VR(e) = blob[(long long)(OFF_VR + e) * K]; }
  if (FULL) {
#pragma unroll
    for (int e = 0; e < NQ; ++e) { SM_YI(e) = blob[(long long)(OFF_YI + e) * K]; SM_VI(e) = blob[(long long)(OFF_VI + e) * K]; }
  }
  double mS = blob[(long long)(OFF_MC + 0) * K], mSmin = blob[(long long)(OFF_MC + 1) * K], mStmp = blob[(long long)(OFF_MC + 2) * K],
         mp = blob[(long long)(OFF_MC + 3) * K], mlam = blob[(long long)(OFF_MC + 4) * K];

  double ar[M], ai[M];
  if (a.a0) {
#pragma unroll
    for (int m = 0; m < M; ++m) { double2 v = a.a0[(long long)m * K + k]; ar[m] = v.x; ai[m] = v.y; }
  }
  int frm = a.frm_cnt, ell = a.ell % a.mc.L;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
  const double one_m_alpha = 1.0 - a.alpha;

  startup_dephase(8000);
  for (int t = 0; t < a.T; ++t) {
    const long long xb = ((long long)s * a.T + t) * M * K;
    double yr[M], yi[M];
#pragma unroll
    for (int m = 0; m < M; ++m) XLoad<XT>::ld(a.X, xb + (long long)m * K + k, yr[m], yi[m]);
    double Ym1 = 0.0, Yp1 = 0.0;
    if (k > 0) { double r, i; XLoad<XT>::ld(a.X, xb + k - 1, r, i); Ym1 = power_c(r, i); }
    if (k < K - 1) { double r, i; XLoad<XT>::ld(a.X, xb + k + 1, r, i); Yp1 = power_c(r, i); }
    const double Y0 = power_c(yr[0], yi[0]);

    // ---- A = inv(Re Phi_vv + eps I)                      mcspp_base.py:278
    double A[M][M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) A[i][j] = SM_VR(pidx<M>(i, j)) + ((i == j) ? a.eps : 0.0);
    spd_inverse_upper<M>(A);
#define AS(i, j) (((i) <= (j)) ? A[i][j] : A[j][i])

    if (a.tAinv && t == a.T - 1) {
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) a.tAinv[(((long long)s * K + k) * M + i) * M + j] = AS(i, j);
    }

    // ---- u = A y ; MVDR numerator b = A a, denominator a^H b   beamformer.py:152-153
    double ur[M], ui[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) { sr = fma(AS(i, j), yr[j], sr); si = fma(AS(i, j), yi[j], si); }
      ur[i] = sr; ui[i] = si;
    }
    double Yr = 0.0, Yi = 0.0, den_r = 1.0, den_i = 0.0;
    double br[M], bi[M];
    if (a.a0) {
      den_r = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) { sr = fma(AS(i, j), ar[j], sr); si = fma(AS(i, j), ai[j], si); }
        br[i] = sr; bi[i] = si;
        // a^H b : conj(a_i) * b_i
        den_r = fma(ar[i], sr, fma(ai[i], si, den_r));
        den_i = fma(ar[i], si, fma(-ai[i], sr, den_i));
        // b^H y : conj(b_i) * y_i
        Yr = fma(sr, yr[i], fma(si, yi[i], Yr));
        Yi = fma(sr, yi[i], fma(-si, yr[i], Yi));
      }
    }

    // ---- Phi_yy update, Phi_xx = Phi_yy - Phi_vv, xi = tr(A Xr), gamma = u^H Xr u   :84-90,274-284
    double tr = 0.0;
    double xur[M], xui[M];
#pragma unroll
    for (int i = 0; i < M; ++i) { xur[i] = 0.0; xui[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double psd = fma(yi[i], yi[j], yr[i] * yr[j]);
        const double pyy = fma(a.alpha, SM_YR(e), one_m_alpha * psd);
        SM_YR(e) = pyy;
        const double x = pyy - SM_VR(e);
        if (i == j) {
          tr = fma(A[i][j], x, tr);
          xur[i] = fma(x, ur[i], xur[i]); xui[i] = fma(x, ui[i], xui[i]);
        } else {
          tr = fma(2.0 * A[i][j], x, tr);
          xur[i] = fma(x, ur[j], xur[i]); xui[i] = fma(x, ui[j], xui[i]);
          xur[j] = fma(x, ur[i], xur[j]); xui[j] = fma(x, ui[i], xui[j]);
        }
      }
    }
    double gam = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) gam = fma(ur[i], xur[i], fma(ui[i], xui[i], gam));
    double xi = fmin(fmax(tr, a.snr_min), a.snr_max);                     // :286-287
    gam = fmin(fmax(gam, a.snr_min), a.snr_max);

    // ---- prior from MCRA on channel 0                                :98-122
    const bool reset = (frm > 0) && (ell == 0);
    mcra_step(mS, mSmin, mStmp, mp, mlam, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
    if (reset) ell = 0;
    ++ell; ++frm;
    if (ell == a.mc.L) ell = 0;
    double q = sqrt(1.0 - mp);
    q = fmin(fmax(q, a.q_min), a.q_max);
    // ---- posterior SPP                                               :124-138
    double p = 1.0 / (1.0 + q / (1.0 - q) * (1.0 + xi) * exp(-1.0 * (gam / (1.0 + xi))));
    p = fmin(fmax(p, a.p_min), a.p_max);

    // ---- PMWF weights (needs Phi_xx[:,0] before Phi_vv moves)        :220-240
    if (FULL) {
      // Phi_xx[j][0] = conj(Phi_xx[0][j]) = XR_0j - i XI_0j
      double cr[M], ci[M];
      cr[0] = SM_YR(pidx<M>(0, 0)) - SM_VR(pidx<M>(0, 0)); ci[0] = 0.0;
#pragma unroll
      for (int j = 1; j < M; ++j) {
        cr[j] = SM_YR(pidx<M>(0, j)) - SM_VR(pidx<M>(0, j));
        // imaginary part of Phi_yy[0][j] after this frame's update
        const double psdi = fma(yi[0], yr[j], -yr[0] * yi[j]);
        const double pyi = fma(a.alpha, SM_YI(qidx<M>(0, j)), one_m_alpha * psdi);
        ci[j] = -(pyi - SM_VI(qidx<M>(0, j)));
      }
      if (a.tw_pmwf) {
        const double dn = 1.0 / (1.0 + xi);
#pragma unroll
        for (int i = 0; i < M; ++i) {
          double sr = 0.0, si = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j) { sr = fma(AS(i, j), cr[j], sr); si = fma(AS(i, j), ci[j], si); }
          a.tw_pmwf[(((long long)s * a.T + t) * M + i) * K + k] = make_double2(sr * dn, si * dn);
        }
      }
    }

    // ---- noise PSD update                                            :299-319
    const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
    const double one_m_at = 1.0 * (1.0 - at);
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double psd = fma(yi[i], yi[j], yr[i] * yr[j]);
        SM_VR(e) = fma(at, SM_VR(e), one_m_at * psd);
        if (FULL && i < j) {
          const int f = qidx<M>(i, j);
          const double psdi = fma(yi[i], yr[j], -yr[i] * yi[j]);
          SM_YI(f) = fma(a.alpha, SM_YI(f), one_m_alpha * psdi);
          SM_VI(f) = fma(at, SM_VI(f), one_m_at * psdi);
        }
      }
    }

    // ---- OMLSA gain                                                  :140-155
    const double GH1 = xi / (1.0 + xi);
    double G = exp(p * log(GH1) + (1.0 - p) * a.logGmin);   // GH1^p * Gmin^(1-p)
    G = fmax(fmin(G, 1.0), a.Gmin);
    if (k < 2) G = 0.0;

    const long long o = ((long long)s * a.T + t) * K + k;
    if (a.tp) a.tp[o] = p;
    if (a.txi) a.txi[o] = xi;
    if (a.tgamma) a.tgamma[o] = gam;
    if (a.tq) a.tq[o] = q;
    if (a.tG) a.tG[o] = G;
    if (a.a0) {
      // w = b / den ;  Y = sum conj(w) y = (b^H y) / conj(den)
      const double dn2 = 1.0 / (den_r * den_r + den_i * den_i);
      if (a.tw_mvdr) {
#pragma unroll
        for (int i = 0; i < M; ++i) {
          const double wr = (br[i] * den_r + bi[i] * den_i) * dn2;
          const double wi = (bi[i] * den_r - br[i] * den_i) * dn2;
          a.tw_mvdr[(((long long)s * a.T + t) * M + i) * K + k] = make_double2(wr, wi);
        }
      }
      if (a.Yout) {
        // (Yr + i Yi) / (den_r - i den_i) = (Yr + i Yi)(den_r + i den_i) / |den|^2
        double or_ = (Yr * den_r - Yi * den_i) * dn2;
        double oi_ = (Yr * den_i + Yi * den_r) * dn2;
        if (a.apply_gain) { or_ *= G; oi_ *= G; }
        a.Yout[o] = make_float2((float)or_, (float)oi_);
        if (a.k_first == 2 && k == 2) {        // bins 0,1 are not tracked in output-only mode: G[:2] = 0
          a.Yout[o - 1] = make_float2(0.f, 0.f);
          a.Yout[o - 2] = make_float2(0.f, 0.f);
        }
      }
    }
#undef AS
  }

#pragma unroll
  for (int e = 0; e < NP; ++e) { blob[(long long)(OFF_YR + e) * K] = SM_YR(e); blob[(long long)(OFF_VR + e) * K] = SM_VR(e); }
  if (FULL) {
#pragma unroll
    for (int e = 0; e < NQ; ++e) { blob[(long long)(OFF_YI + e) * K] = SM_YI(e); blob[(long long)(OFF_VI + e) * K] = SM_VI(e); }
  }
  blob[(long long)(OFF_MC + 0) * K] = mS; blob[(long long)(OFF_MC + 1) * K] = mSmin; blob[(long long)(OFF_MC + 2) * K] = mStmp;
  blob[(long long)(OFF_MC + 3) * K] = mp; blob[(long long)(OFF_MC + 4) * K] = mlam;
#undef SM_YR
#undef SM_VR
#undef SM_YI
#undef SM_VI
}

template <int M, bool FULL, typename XT>
static int launch_mcspp_t(const McsppArgs &a, cudaStream_t st) {
  constexpr int NT = FULL ? 64 : 128;
  constexpr int NP = M * (M + 1) / 2, NQ = M * (M - 1) / 2;
  const size_t smem = (size_t)(FULL ? (2 * NP + 2 * NQ) : (2 * NP)) * NT * sizeof(double);
  auto kern = mcspp_kernel<M, FULL, XT, NT>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)a.S * (a.K - a.k_first);
  const unsigned blocks = (unsigned)((items + NT - 1) / NT);
  kern<<<blocks, NT, smem, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

template <int M>
static int launch_mcspp_m(const McsppArgs &a, bool full, bool c128, cudaStream_t st) {
  if (full) return c128 ? launch_mcspp_t<M, true, double2>(a, st) : launch_mcspp_t<M, true, float2>(a, st);
  return c128 ? launch_mcspp_t<M, false, double2>(a, st) : launch_mcspp_t<M, false, float2>(a, st);
}

int launch_mcspp(int M, const McsppArgs &a, bool full, bool c128, cudaStream_t st) {
  switch (M) {
    case 2: return launch_mcspp_m<2>(a, full, c128, st);
    case 3: return launch_mcspp_m<3>(a, full, c128, st);
    case 4: return launch_mcspp_m<4>(a, full, c128, st);
    case 5: return launch_mcspp_m<5>(a, full, c128, st);
    case 6: return launch_mcspp_m<6>(a, full, c128, st);
    case 7: return launch_mcspp_m<7>(a, full, c128, st);
    case 8: return launch_mcspp_m<8>(a, full, c128, st);
  }
  set_error("mcspp: n_mics %d outside the compiled range 2..8", M);
  return DS_EUNSUPPORTED;
}

// export Phi_yy / Phi_vv as dense complex [S][K][M][M]
__global__ void mcspp_export_kernel(const double *state, double2 *out, int S, int K, int M, int which) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * K * M * M;
  if (g >= total) return;
  const int j = (int)(g % M), i = (int)((g / M) % M), k = (int)((g / ((long long)M * M)) % K), s = (int)(g / ((long long)M * M * K));
  const int NP = M * (M + 1) / 2, NQ = M * (M - 1) / 2, NE = 2 * M * M + 5;
  const int offR = which ? NP : 0, offI = 2 * NP + 5 + (which ? NQ : 0);
  const int lo = min(i, j), hi = max(i, j);
  const double *b = state + (long long)s * NE * K + k;
  double re = b[(long long)(offR + lo * M - (lo * (lo - 1)) / 2 + (hi - lo)) * K];
  double im = 0.0;
  if (i != j) {
    im = b[(long long)(offI + lo * (M - 1) - (lo * (lo - 1)) / 2 + (hi - lo - 1)) * K];
    if (i > j) im = -im;
  }
  out[g] = make_double2(re, im);
}

static void fill_args(const ds_mcspp_params *p, McsppArgs &a) {
  a.S = p->n_streams; a.K = p->n_fft / 2 + 1; a.T = p->n_frames; a.frm_cnt = p->frm_cnt; a.ell = p->ell;
  a.alpha = p->alpha; a.alpha_d = p->alpha_d; a.eps = p->diag_eps; a.q_min = p->q_min; a.q_max = p->q_max;
  a.p_min = p->p_min; a.p_max = p->p_max; a.snr_min = p->snr_min; a.snr_max = p->snr_max; a.Gmin = p->Gmin;
  a.logGmin = log(p->Gmin);
  a.mc.alpha_d = p->mcra_alpha_d; a.mc.alpha_s = p->mcra_alpha_s; a.mc.delta_s = p->mcra_delta_s;
  a.mc.alpha_p = p->mcra_alpha_p; a.mc.p_min = p->mcra_p_min; a.mc.p_max = p->mcra_p_max; a.mc.L = p->mcra_L;
}

int mcspp_run_impl(const ds_mcspp_params *p, void *state, const void *a0, const void *X, int x_is_c128, void *Yout,
                   int apply_gain, const ds_mcspp_taps *taps, cudaStream_t st) {
  McsppArgs a;
  fill_args(p, a);
  a.state = (double *)state; a.a0 = (const double2 *)a0; a.X = X; a.Yout = (float2 *)Yout; a.apply_gain = apply_gain;
  a.tp = taps ? taps->p : nullptr; a.txi = taps ? taps->xi : nullptr; a.tgamma = taps ? taps->gamma : nullptr;
  a.tq = taps ? taps->q : nullptr; a.tG = taps ? taps->G : nullptr;
  a.tw_mvdr = taps ? (double2 *)taps->w_mvdr : nullptr; a.tw_pmwf = taps ? (double2 *)taps->w_pmwf : nullptr;
  a.tAinv = taps ? taps->Phi_vv_inv_last : nullptr;
  const bool any_tap = a.tp || a.txi || a.tgamma || a.tq || a.tG || a.tw_mvdr || a.tw_pmwf || a.tAinv;
  // output-only mode may skip bins 0 and 1 (their gain is identically 0)
  const bool full = p->full_state != 0;
  a.k_first = (!full && apply_gain && a0 && Yout && !any_tap) ? 2 : 0;
  // the output-only kernel also serves the p tap alone (it then has to visit every bin)
  const bool only_p = a.tp && !(a.txi || a.tgamma || a.tq || a.tG || a.tw_mvdr || a.tw_pmwf || a.tAinv);
  if (!full && a0 && Yout && (!any_tap || only_p) && !x_is_c128) return launch_mcspp_fast(p->n_mics, a, st);
  return launch_mcspp(p->n_mics, a, full, x_is_c128 != 0, st);
}

}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_mcra_state_bytes(const ds_mcra_params *p) {
  if (!p) return 0;
  return (size_t)p->n_streams * 5 * p->n_bins * sizeof(double);
}

void ds_mcra_advance(int32_t L, int32_t n_frames, int32_t *frm_cnt, int32_t *ell) {
  int f = *frm_cnt, e = *ell;
  for (int t = 0; t < n_frames; ++t) {
    if (f > 0 && L > 0 && (e % L == 0)) e = 0;
    ++e; ++f;
  }
  *frm_cnt = f; *ell = e;
}

int ds_mcra_run(const ds_mcra_params *p, void *state, const double *Ypow, double *lambda_out, double *p_out, void *stream) {
  DS_CHECK_ARG(p && state && Ypow, "ds_mcra_run: null argument");
  DS_CHECK_ARG(p->n_bins >= 1 && p->n_streams >= 1 && p->n_frames >= 1 && p->L >= 1, "ds_mcra_run: bad shape");
  McraArgs a;
  a.state = (double *)state; a.Y = Ypow; a.lam_out = lambda_out; a.p_out = p_out;
  a.S = p->n_streams; a.K = p->n_bins; a.T = p->n_frames; a.frm_cnt = p->frm_cnt; a.ell = p->ell;
  a.c.alpha_d = p->alpha_d; a.c.alpha_s = p->alpha_s; a.c.delta_s = p->delta_s; a.c.alpha_p = p->alpha_p;
  a.c.p_min = p->p_min; a.c.p_max = p->p_max; a.c.L = p->L;
  const long long items = (long long)a.S * a.K;
  mcra_kernel<<<(unsigned)((items + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

void ds_mcspp_default_params(ds_mcspp_params *p, int n_fft, int n_streams, int n_mics, int n_frames) {
  if (!p) return;
  p->n_fft = n_fft; p->n_streams = n_streams; p->n_mics = n_mics; p->n_frames = n_frames;
  p->frm_cnt = 0; p->ell = 1; p->mcra_L = 15; p->full_state = 0;
  p->alpha = 0.92; p->alpha_d = 0.92; p->diag_eps = 1e-6;
  p->q_min = 0.01; p->q_max = 0.99; p->p_min = 0.01; p->p_max = 0.99;
  p->snr_min = 1e-6; p->snr_max = 1e6; p->Gmin = 0.0631;
  p->mcra_alpha_d = 0.95; p->mcra_alpha_s = 0.8; p->mcra_delta_s = 5.0; p->mcra_alpha_p = 0.2;
  p->mcra_p_min = 1e-3; p->mcra_p_max = 0.999;
}

size_t ds_mcspp_state_bytes(const ds_mcspp_params *p) {
  if (!p) return 0;
  const int M = p->n_mics, K = p->n_fft / 2 + 1;
  return (size_t)p->n_streams * (2 * M * M + 5) * K * sizeof(double);
}

int ds_mcspp_run(const ds_mcspp_params *p, void *state, const void *a0, const void *X, int x_is_c128, void *Yout,
                 int apply_gain, const ds_mcspp_taps *taps, void *stream) {
  DS_CHECK_ARG(p && state && X, "ds_mcspp_run: null argument");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_frames >= 1 && p->n_fft >= 4 && p->mcra_L >= 1, "ds_mcspp_run: bad shape");
  DS_CHECK_ARG(!(taps && taps->w_pmwf) || p->full_state, "ds_mcspp_run: w_pmwf tap needs full_state");
  DS_CHECK_ARG(!(taps && taps->w_mvdr) || a0, "ds_mcspp_run: w_mvdr tap needs a0");
  return mcspp_run_impl(p, state, a0, X, x_is_c128, Yout, apply_gain, taps, (cudaStream_t)stream);
}

int ds_mcspp_export(const ds_mcspp_params *p, const void *state, int field, void *out, void *stream) {
  DS_CHECK_ARG(p && state && out, "ds_mcspp_export: null argument");
  const int M = p->n_mics, K = p->n_fft / 2 + 1, S = p->n_streams;
  cudaStream_t st = (cudaStream_t)stream;
  if (field == 0 || field == 1) {
    const long long total = (long long)S * K * M * M;
    mcspp_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const double *)state, (double2 *)out, S, K, M, field);
    DS_LAUNCH_CHECK();
    return DS_OK;
  }
  if (field == 2) {
    const int NE = 2 * M * M + 5, NP = M * (M + 1) / 2;
    DS_CUDA(cudaMemcpy2DAsync(out, (size_t)5 * K * sizeof(double), (const double *)state + (size_t)2 * NP * K,
                              (size_t)NE * K * sizeof(double), (size_t)5 * K * sizeof(double), S, cudaMemcpyDeviceToDevice, st));
    return DS_OK;
  }
  set_error("ds_mcspp_export: unknown field %d", field);
  return DS_EINVAL;
}

}  // extern "C"
