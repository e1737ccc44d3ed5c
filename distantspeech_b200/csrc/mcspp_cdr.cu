// mcspp_cdr.cu -- McSpp (noise_estimation/mcspp.py:46-305) with its McCDR prior
// (noise_estimation/mccdr.py:25-192, coherence/BinauralEnhancement.py:24-59), 4 microphones.
//
// Per frame the reference does (file:line):
//   q = 1 - McCDR.estimation(y)                               mcspp.py:117-118
//        Pxii/Pxij recursions (alpha .9)                       BinauralEnhancement.py:44-59, mccdr.py:126-127
//        Fx = Pxij(1,2) / sqrt(Pxii_1 Pxii_2)                  BinauralEnhancement.py:24-29
//        unbiased CDR of pair (1,2), squared, clipped          mccdr.py:134-159
//        sqrt(CDR * p_mcra),  MCRA with L = 65 on channel 0    mccdr.py:172-175
//   loading = q_avg 1e-1 + (1 - q_avg) 1e-4, q_avg = mean(q[500 Hz .. 2 kHz])     mcspp.py:262-269
//   Phi_yy = .92 Phi_yy + .08 y y^H                            :271-273
//   first 10 frames: Phi_vv = Phi_yy, q = .99                  :276-278
//   Phi_vv_inv = inv(herm(Phi_vv) + loading I); xi = tr Re(Phi_vv_inv Phi_yy) - M; where xi < 0 the
//   inverse falls back to inv(Phi_yy [+ loading I while frm_cnt < 5]); gamma; p         :199-243
//   Phi_vv = at Phi_vv + (1 - at) y y^H, at = .92 + .08 p      mcspp_base.py:312-319
//   w = Phi_vv_inv Phi_xx e_0 / (10 + xi)                      mcspp_base.py:238-240, mcspp.py:286
//
// The only coupling between bins is q_avg, so the work is cut into three launches with no
// per-frame synchronisation: (1) the CDR prior for every (stream, frame, bin), one thread per
// (stream, bin) walking the frames; (2) q_avg per (stream, frame) in NumPy's pairwise order;
// (3) the covariance recursions, again one thread per (stream, bin), Hermitian matrices packed
// in registers.  All recursive state is float64 like the reference.
#include "common.cuh"
#include "herm.cuh"

namespace ds {

constexpr int CDR_M = 4, CDR_NQ = 6;
// state blob [S][CDR_NE][K] float64
enum {
  CS_PYY = 0,      // d[4] ur[6] ui[6]
  CS_PVV = 16,
  CS_PXII = 32,    // [4]
  CS_PXIJ_R = 36,  // [6] pairs in (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) order
  CS_PXIJ_I = 42,
  CS_MCRA = 48,    // S Smin Stmp p lambda
  CS_P = 53,
  CS_AINV = 54,    // outputs of the last frame from here on
  CS_PXX = 70,
  CS_W_R = 86, CS_W_I = 90,
  CS_XI = 94, CS_GAMMA = 95, CS_Q = 96, CS_CDR = 97,
  CDR_NE = 98
};

struct CdrArgs {
  double *state;
  const void *X;            // [S][T][4][K]
  int x_c128;
  const double *Fn;         // [K] diffuse coherence of pair (1,2)
  double *q;                // [S][T][K] workspace: prior speech absence probability
  double *qavg;             // [S][T]    workspace
  double *tp, *txi, *tgamma, *tq, *tcdr;     // taps [S][T][K] or null
  double2 *tw;              // [S][T][4][K] or null
  float2 *Yout;             // [S][T][K] or null
  int S, K, T, frm_cnt, ell, lo, hi, init_frames, fallback_frames;
  double alpha, alpha_d, alpha_cdr, load_min, load_max, snr_min, snr_max, beta, q_init;
  McraConst mc;
};

__device__ __forceinline__ void load_y4(const CdrArgs &a, long long base, int K, double (&yr)[4], double (&yi)[4]) {
  if (a.x_c128) {
    const double2 *X = reinterpret_cast<const double2 *>(a.X) + base;
#pragma unroll
    for (int m = 0; m < 4; ++m) { const double2 v = X[(long long)m * K]; yr[m] = v.x; yi[m] = v.y; }
  } else {
    const float2 *X = reinterpret_cast<const float2 *>(a.X) + base;
#pragma unroll
    for (int m = 0; m < 4; ++m) { const float2 v = X[(long long)m * K]; yr[m] = (double)v.x; yi[m] = (double)v.y; }
  }
}
__device__ __forceinline__ double load_pow0(const CdrArgs &a, long long idx) {
  double re, im;
  if (a.x_c128) { const double2 v = reinterpret_cast<const double2 *>(a.X)[idx]; re = v.x; im = v.y; }
  else { const float2 v = reinterpret_cast<const float2 *>(a.X)[idx]; re = (double)v.x; im = (double)v.y; }
  const double h = hypot(re, im);        // mcra.py:29-30: np.abs(Y) ** 2 for complex input
  return __dmul_rn(h, h);
}

// ---- (1) CDR prior ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cdr_prior_kernel(CdrArgs a) {
  const int K = a.K;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *blob = a.state + (long long)s * CDR_NE * K + k;
  double pii[4], pr[CDR_NQ], pi[CDR_NQ];
#pragma unroll
  for (int i = 0; i < 4; ++i) pii[i] = blob[(long long)(CS_PXII + i) * K];
#pragma unroll
  for (int e = 0; e < CDR_NQ; ++e) { pr[e] = blob[(long long)(CS_PXIJ_R + e) * K]; pi[e] = blob[(long long)(CS_PXIJ_I + e) * K]; }
  double mS = blob[(long long)(CS_MCRA + 0) * K], mSmin = blob[(long long)(CS_MCRA + 1) * K], mStmp = blob[(long long)(CS_MCRA + 2) * K],
         mp = blob[(long long)(CS_MCRA + 3) * K], mlam = blob[(long long)(CS_MCRA + 4) * K];
  const double Fn = a.Fn[k], Fn2 = __dmul_rn(Fn, Fn);
  const double al = a.alpha_cdr, om = __dsub_rn(1.0, al);
  int frm = a.frm_cnt, ell = a.ell % a.mc.L;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
  double cdr = 0.0;
  for (int t = 0; t < a.T; ++t) {
    const long long base = ((long long)s * a.T + t) * 4 * K + k;
    double yr[4], yi[4];
    load_y4(a, base, K, yr, yi);
    // auto / cross spectra, operation order of the reference (no contraction)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      pii[i] = __dadd_rn(__dmul_rn(al, pii[i]), __dmul_rn(om, __dadd_rn(__dmul_rn(yr[i], yr[i]), __dmul_rn(yi[i], yi[i]))));
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        const int e = qidx<4>(i, j);
        const double cr = __dadd_rn(__dmul_rn(yr[i], yr[j]), __dmul_rn(yi[i], yi[j]));      // y_i conj(y_j)
        const double ci = __dsub_rn(__dmul_rn(yi[i], yr[j]), __dmul_rn(yr[i], yi[j]));
        pr[e] = __dadd_rn(__dmul_rn(al, pr[e]), __dmul_rn(om, cr));
        pi[e] = __dadd_rn(__dmul_rn(al, pi[e]), __dmul_rn(om, ci));
      }
    // coherence of pair (1,2): complex / real the way NumPy divides (multiply by 1/d)
    const int e12 = qidx<4>(1, 2);
    const double scl = 1.0 / sqrt(__dmul_rn(pii[1], pii[2]));
    const double fr = __dmul_rn(pr[e12], scl), fi = __dmul_rn(pi[e12], scl);
    const double fa = hypot(fr, fi);
    const double Fx2 = __dmul_rn(fa, fa);
    // mccdr.py:141-145, left to right
    double rad = __dsub_rn(__dmul_rn(Fn2, __dmul_rn(fr, fr)), __dmul_rn(Fn2, Fx2));
    rad = __dadd_rn(rad, Fn2);
    rad = __dsub_rn(rad, __dmul_rn(__dmul_rn(2.0, Fn), fr));
    rad = __dadd_rn(rad, Fx2);
    const double num = __dsub_rn(__dsub_rn(__dmul_rn(Fn, fr), Fx2), sqrt(rad));
    const double den = fmin(__dsub_rn(Fx2, 1.0), -1e-3);
    double G = num / den;
    G = __dmul_rn(G, G);
    if (G > 1.0) G = 1.0;                           // NaN falls through both tests, like NumPy's masks
    if (G < 0.0) G = 1e-3;
    cdr = G;
    // MCRA (L = 65) on channel 0
    const long long i0 = ((long long)s * a.T + t) * 4 * K;
    const double Y0 = load_pow0(a, i0 + k);
    const double Ym1 = (k > 0) ? load_pow0(a, i0 + k - 1) : 0.0;
    const double Yp1 = (k < K - 1) ? load_pow0(a, i0 + k + 1) : 0.0;
    const bool reset = (frm > 0) && (ell == 0);
    mcra_step(mS, mSmin, mStmp, mp, mlam, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
    if (reset) ell = 0;
    ++ell; ++frm;
    if (ell == a.mc.L) ell = 0;
    const double gam = sqrt(__dmul_rn(G, mp));      // McCDR.estimation return value
    const long long o = ((long long)s * a.T + t) * K + k;
    a.q[o] = __dsub_rn(1.0, gam);
    if (a.tcdr) a.tcdr[o] = gam;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) blob[(long long)(CS_PXII + i) * K] = pii[i];
#pragma unroll
  for (int e = 0; e < CDR_NQ; ++e) { blob[(long long)(CS_PXIJ_R + e) * K] = pr[e]; blob[(long long)(CS_PXIJ_I + e) * K] = pi[e]; }
  blob[(long long)(CS_MCRA + 0) * K] = mS; blob[(long long)(CS_MCRA + 1) * K] = mSmin; blob[(long long)(CS_MCRA + 2) * K] = mStmp;
  blob[(long long)(CS_MCRA + 3) * K] = mp; blob[(long long)(CS_MCRA + 4) * K] = mlam;
  blob[(long long)CS_CDR * K] = cdr;
}

// ---- (2) q_avg = np.mean(q[lo:hi]) with NumPy's pairwise summation order ----------------------------
__device__ double np_sum_block(const double *v, int n) {        // n <= 128
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, v[i]);
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], v[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, v[i]);
  return res;
}
__device__ double np_sum(const double *v, int n) {
  if (n <= 128) return np_sum_block(v, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(np_sum(v, n2), np_sum(v + n2, n - n2));
}
__global__ void cdr_qavg_kernel(CdrArgs a) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * a.T) return;
  const int n = a.hi - a.lo;
  a.qavg[g] = np_sum(a.q + g * a.K + a.lo, n) / (double)n;
}

// ---- (3) covariance recursions, SPP, PMWF weights -----------------------------------------------------
__device__ __forceinline__ void herm_load(Herm<4> &h, const double *blob, int off, int K) {
#pragma unroll
  for (int i = 0; i < 4; ++i) h.d[i] = blob[(long long)(off + i) * K];
#pragma unroll
  for (int e = 0; e < CDR_NQ; ++e) { h.ur[e] = blob[(long long)(off + 4 + e) * K]; h.ui[e] = blob[(long long)(off + 10 + e) * K]; }
}
__device__ __forceinline__ void herm_store(const Herm<4> &h, double *blob, int off, int K) {
#pragma unroll
  for (int i = 0; i < 4; ++i) blob[(long long)(off + i) * K] = h.d[i];
#pragma unroll
  for (int e = 0; e < CDR_NQ; ++e) { blob[(long long)(off + 4 + e) * K] = h.ur[e]; blob[(long long)(off + 10 + e) * K] = h.ui[e]; }
}
// Re tr(A B) for Hermitian A, B
__device__ __forceinline__ double herm_trace_prod(const Herm<4> &A, const Herm<4> &B) {
  double d = 0.0, o = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) d = fma(A.d[i], B.d[i], d);
#pragma unroll
  for (int e = 0; e < CDR_NQ; ++e) o = fma(A.ur[e], B.ur[e], fma(A.ui[e], B.ui[e], o));
  return fma(2.0, o, d);
}
// u = A y
__device__ __forceinline__ void herm_matvec(const Herm<4> &A, const double (&yr)[4], const double (&yi)[4], double (&ur)[4], double (&ui)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double sr = A.d[i] * yr[i], si = A.d[i] * yi[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j == i) continue;
      const double ar = (i < j) ? A.ur[qidx<4>(i, j)] : A.ur[qidx<4>(j, i)];
      const double ai = (i < j) ? A.ui[qidx<4>(i, j)] : -A.ui[qidx<4>(j, i)];
      sr = fma(ar, yr[j], fma(-ai, yi[j], sr));
      si = fma(ar, yi[j], fma(ai, yr[j], si));
    }
    ur[i] = sr; ui[i] = si;
  }
}
// Re(u^H B u)
__device__ __forceinline__ double herm_quad(const Herm<4> &B, const double (&ur)[4], const double (&ui)[4]) {
  double d = 0.0, o = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) d = fma(B.d[i], fma(ur[i], ur[i], ui[i] * ui[i]), d);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      // conj(u_i) B_ij u_j + c.c. = 2 Re(conj(u_i) u_j B_ij)
      const int e = qidx<4>(i, j);
      const double cr = fma(ur[i], ur[j], ui[i] * ui[j]);      // conj(u_i) u_j
      const double ci = fma(ur[i], ui[j], -ui[i] * ur[j]);
      o = fma(cr, B.ur[e], fma(-ci, B.ui[e], o));
    }
  return fma(2.0, o, d);
}

__global__ void __launch_bounds__(64) mcspp_cdr_kernel(CdrArgs a) {
  const int K = a.K;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *blob = a.state + (long long)s * CDR_NE * K + k;
  Herm<4> Pyy, Pvv, A, Pxx;
  herm_load(Pyy, blob, CS_PYY, K);
  herm_load(Pvv, blob, CS_PVV, K);
  double p_prev = blob[(long long)CS_P * K];
  int frm = a.frm_cnt;
  double xi = 0.0, gamma = 0.0, q = 0.0, wr[4] = {0, 0, 0, 0}, wi[4] = {0, 0, 0, 0};
  const double om_alpha = 1.0 - a.alpha;
  for (int t = 0; t < a.T; ++t, ++frm) {
    const long long base = ((long long)s * a.T + t) * 4 * K + k;
    const long long o = ((long long)s * a.T + t) * K + k;
    double yr[4], yi[4];
    load_y4(a, base, K, yr, yi);
    q = a.q[o];
    const double qa = a.qavg[(long long)s * a.T + t];
    const double load = qa * a.load_max + (1.0 - qa) * a.load_min;                 // mcspp.py:269
    // Phi_yy                                                                        :271-273
#pragma unroll
    for (int i = 0; i < 4; ++i) Pyy.d[i] = a.alpha * Pyy.d[i] + om_alpha * fma(yr[i], yr[i], yi[i] * yi[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        const int e = qidx<4>(i, j);
        Pyy.ur[e] = a.alpha * Pyy.ur[e] + om_alpha * fma(yr[i], yr[j], yi[i] * yi[j]);
        Pyy.ui[e] = a.alpha * Pyy.ui[e] + om_alpha * fma(yi[i], yr[j], -yr[i] * yi[j]);
      }
    if (frm < a.init_frames) { Pvv = Pyy; q = a.q_init; }                           // :276-278
    // estimation_core                                                               :199-243
#pragma unroll
    for (int i = 0; i < 4; ++i) Pxx.d[i] = Pyy.d[i] - Pvv.d[i];
#pragma unroll
    for (int e = 0; e < CDR_NQ; ++e) { Pxx.ur[e] = Pyy.ur[e] - Pvv.ur[e]; Pxx.ui[e] = Pyy.ui[e] - Pvv.ui[e]; }
    A = Pvv;
#pragma unroll
    for (int i = 0; i < 4; ++i) A.d[i] += load;
    herm_inverse<4, true>(A);
    xi = herm_trace_prod(A, Pyy) - 4.0;                                              // :217
    if (xi < 0.0) {                                                                  // :220-228
      A = Pyy;
      if (frm < a.fallback_frames) {
#pragma unroll
        for (int i = 0; i < 4; ++i) A.d[i] += load;
      }
      herm_inverse<4, true>(A);
      xi = herm_trace_prod(A, Pyy) - 4.0;
    }
    xi = (xi != xi) ? xi : fmin(fmax(xi, a.snr_min), a.snr_max);                     // :229 (NaN propagates like np.minimum)
    double ur[4], ui[4];
    herm_matvec(A, yr, yi, ur, ui);
    double yAy = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) yAy = fma(yr[i], ur[i], fma(yi[i], ui[i], yAy));
    gamma = herm_quad(Pyy, ur, ui) - yAy;                                            // :231-235
    gamma = (gamma != gamma) ? gamma : fmin(fmax(gamma, a.snr_min), a.snr_max);
    // compute_p(alpha_p = 0)                                                        :76-91
    double p = 1.0 / (1.0 + q / (1.0 - q) * (1.0 + xi) * exp(-1.0 * (gamma / (1.0 + xi))));
    p = 0.0 * p_prev + p;
    p = (p != p) ? p : fmin(fmax(p, 0.0), 1.0);
    p_prev = p;
    // PMWF weights from the matrices estimation_core used (Phi_xx before the noise update)   :286
    {
      // column 0 of Phi_xx: (Pxx_00, conj(Pxx_01), conj(Pxx_02), conj(Pxx_03))
      const double cr[4] = {Pxx.d[0], Pxx.ur[qidx<4>(0, 1)], Pxx.ur[qidx<4>(0, 2)], Pxx.ur[qidx<4>(0, 3)]};
      const double ci[4] = {0.0, -Pxx.ui[qidx<4>(0, 1)], -Pxx.ui[qidx<4>(0, 2)], -Pxx.ui[qidx<4>(0, 3)]};
      herm_matvec(A, cr, ci, wr, wi);
      const double sc = 1.0 / (a.beta + xi);
#pragma unroll
      for (int i = 0; i < 4; ++i) { wr[i] *= sc; wi[i] *= sc; }
    }
    // update_noise_psd(beta = 1)                                                    mcspp_base.py:312-319
    const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
    const double om_at = 1.0 * (1.0 - at);
#pragma unroll
    for (int i = 0; i < 4; ++i) Pvv.d[i] = at * Pvv.d[i] + om_at * fma(yr[i], yr[i], yi[i] * yi[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        const int e = qidx<4>(i, j);
        Pvv.ur[e] = at * Pvv.ur[e] + om_at * fma(yr[i], yr[j], yi[i] * yi[j]);
        Pvv.ui[e] = at * Pvv.ui[e] + om_at * fma(yi[i], yr[j], -yr[i] * yi[j]);
      }
    if (a.tp) a.tp[o] = p;
    if (a.txi) a.txi[o] = xi;
    if (a.tgamma) a.tgamma[o] = gamma;
    if (a.tq) a.tq[o] = q;
    if (a.tw) {
#pragma unroll
      for (int i = 0; i < 4; ++i) a.tw[base + (long long)i * K] = make_double2(wr[i], wi[i]);
    }
    if (a.Yout) {                               // Y = w^H y
      double Yr = 0.0, Yi = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        Yr = fma(wr[i], yr[i], fma(wi[i], yi[i], Yr));
        Yi = fma(wr[i], yi[i], fma(-wi[i], yr[i], Yi));
      }
      a.Yout[o] = make_float2((float)Yr, (float)Yi);
    }
  }
  herm_store(Pyy, blob, CS_PYY, K);
  herm_store(Pvv, blob, CS_PVV, K);
  herm_store(A, blob, CS_AINV, K);
  herm_store(Pxx, blob, CS_PXX, K);
  blob[(long long)CS_P * K] = p_prev;
#pragma unroll
  for (int i = 0; i < 4; ++i) { blob[(long long)(CS_W_R + i) * K] = wr[i]; blob[(long long)(CS_W_I + i) * K] = wi[i]; }
  blob[(long long)CS_XI * K] = xi; blob[(long long)CS_GAMMA * K] = gamma; blob[(long long)CS_Q * K] = q;
}

// ---- export ---------------------------------------------------------------------------------------------
__global__ void cdr_export_herm_kernel(const double *state, int off, int S, int K, double2 *out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * K * 16) return;
  const int ij = (int)(g % 16), k = (int)((g / 16) % K), s = (int)(g / (16LL * K));
  const int i = ij / 4, j = ij % 4;
  const double *b = state + (long long)s * CDR_NE * K + k;
  double re, im = 0.0;
  if (i == j) re = b[(long long)(off + i) * K];
  else {
    const int lo = min(i, j), hi = max(i, j), e = lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo - 1);
    re = b[(long long)(off + 4 + e) * K];
    im = b[(long long)(off + 10 + e) * K];
    if (i > j) im = -im;
  }
  out[g] = make_double2(re, im);
}
__global__ void cdr_export_rows_kernel(const double *state, int off, int n, int S, int K, double *out) {   // -> [S][n][K]
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * n * K) return;
  const int k = (int)(g % K), e = (int)((g / K) % n), s = (int)(g / ((long long)n * K));
  out[g] = state[((long long)s * CDR_NE + off + e) * K + k];
}

}  // namespace ds

using namespace ds;

extern "C" {

void ds_mcspp_cdr_default_params(ds_mcspp_cdr_params *p, int n_fft, int n_streams, int n_mics, int n_frames) {
  if (!p) return;
  p->n_fft = n_fft; p->n_streams = n_streams; p->n_mics = n_mics; p->n_frames = n_frames;
  p->frm_cnt = 0; p->ell = 1; p->mcra_L = 65; p->cdr_only = 0;
  p->band_lo_bin = (int)(500.0 * n_fft / 16000.0); p->band_hi_bin = (int)(2000.0 * n_fft / 16000.0);
  p->init_frames = 10; p->fallback_loaded_frames = 5;
  p->alpha = 0.92; p->alpha_d = 0.92; p->alpha_cdr = 0.9;
  p->load_min = 1e-4; p->load_max = 1e-1; p->snr_min = 1e-6; p->snr_max = 1e8; p->pmwf_beta = 10.0; p->q_init = 0.99;
  p->mcra_alpha_d = 0.95; p->mcra_alpha_s = 0.8; p->mcra_delta_s = 5.0; p->mcra_alpha_p = 0.2; p->mcra_p_min = 1e-3; p->mcra_p_max = 0.999;
}

size_t ds_mcspp_cdr_state_bytes(const ds_mcspp_cdr_params *p) {
  return p ? (size_t)p->n_streams * CDR_NE * (p->n_fft / 2 + 1) * sizeof(double) : 0;
}
size_t ds_mcspp_cdr_workspace_bytes(const ds_mcspp_cdr_params *p) {
  if (!p) return 0;
  const size_t K = p->n_fft / 2 + 1, ST = (size_t)p->n_streams * p->n_frames;
  return (ST * K + ST) * sizeof(double);
}

int ds_mcspp_cdr_run(const ds_mcspp_cdr_params *p, void *state, void *workspace, const double *Fn, const void *X, int x_is_c128,
                     void *Yout, const ds_mcspp_cdr_taps *taps, void *stream) {
  DS_CHECK_ARG(p && state && workspace && Fn && X, "ds_mcspp_cdr_run: null argument");
  if (p->n_mics != 4) {
    set_error("ds_mcspp_cdr_run: n_mics must be 4 (the reference builds McCDR with 4 channels, mcspp.py:54, and crashes above)");
    return DS_EUNSUPPORTED;
  }
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_frames >= 1 && p->n_fft >= 64 && p->mcra_L >= 1, "ds_mcspp_cdr_run: bad shape");
  const int K = p->n_fft / 2 + 1;
  DS_CHECK_ARG(p->band_lo_bin >= 0 && p->band_hi_bin > p->band_lo_bin && p->band_hi_bin <= K, "ds_mcspp_cdr_run: bad prior band");
  CdrArgs a;
  a.state = (double *)state; a.X = X; a.x_c128 = x_is_c128; a.Fn = Fn;
  a.q = (double *)workspace; a.qavg = a.q + (size_t)p->n_streams * p->n_frames * K;
  a.tp = taps ? taps->p : nullptr; a.txi = taps ? taps->xi : nullptr; a.tgamma = taps ? taps->gamma : nullptr;
  a.tq = taps ? taps->q : nullptr; a.tcdr = taps ? taps->cdr : nullptr; a.tw = taps ? (double2 *)taps->w : nullptr;
  a.Yout = (float2 *)Yout;
  a.S = p->n_streams; a.K = K; a.T = p->n_frames; a.frm_cnt = p->frm_cnt; a.ell = p->ell; a.lo = p->band_lo_bin; a.hi = p->band_hi_bin;
  a.init_frames = p->init_frames; a.fallback_frames = p->fallback_loaded_frames;
  a.alpha = p->alpha; a.alpha_d = p->alpha_d; a.alpha_cdr = p->alpha_cdr; a.load_min = p->load_min; a.load_max = p->load_max;
  a.snr_min = p->snr_min; a.snr_max = p->snr_max; a.beta = p->pmwf_beta; a.q_init = p->q_init;
  a.mc.alpha_d = p->mcra_alpha_d; a.mc.alpha_s = p->mcra_alpha_s; a.mc.delta_s = p->mcra_delta_s;
  a.mc.alpha_p = p->mcra_alpha_p; a.mc.p_min = p->mcra_p_min; a.mc.p_max = p->mcra_p_max; a.mc.L = p->mcra_L;
  cudaStream_t st = (cudaStream_t)stream;
  const long long items = (long long)a.S * K;
  cdr_prior_kernel<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  if (p->cdr_only) return DS_OK;
  const long long frames = (long long)a.S * a.T;
  cdr_qavg_kernel<<<(unsigned)((frames + 127) / 128), 128, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  mcspp_cdr_kernel<<<(unsigned)((items + 63) / 64), 64, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int ds_mcspp_cdr_export(const ds_mcspp_cdr_params *p, const void *state, int field, void *out, void *stream) {
  DS_CHECK_ARG(p && state && out, "ds_mcspp_cdr_export: null argument");
  const int K = p->n_fft / 2 + 1, S = p->n_streams;
  cudaStream_t st = (cudaStream_t)stream;
  const double *sd = (const double *)state;
  if (field >= 0 && field <= 3) {
    const int off = field == 0 ? CS_PYY : field == 1 ? CS_PVV : field == 2 ? CS_AINV : CS_PXX;
    const long long n = (long long)S * K * 16;
    cdr_export_herm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sd, off, S, K, (double2 *)out);
  } else {
    int off, cnt;
    switch (field) {
      case 4: off = CS_W_R; cnt = 8; break;       // [S][8][K]: re[4], im[4]
      case 5: off = CS_XI; cnt = 4; break;        // xi gamma q cdr
      case 6: off = CS_PXII; cnt = 16; break;     // Pxii[4] Pxij_re[6] Pxij_im[6]
      case 7: off = CS_MCRA; cnt = 6; break;      // S Smin Stmp p lambda | p (posterior)
      default: set_error("ds_mcspp_cdr_export: unknown field %d", field); return DS_EINVAL;
    }
    const long long n = (long long)S * cnt * K;
    cdr_export_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sd, off, cnt, S, K, (double *)out);
  }
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
