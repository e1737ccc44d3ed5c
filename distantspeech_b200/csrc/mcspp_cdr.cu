// mcspp_cdr.cu -- McSpp (noise_estimation/mcspp.py:46-305) with its McCDR prior
// (noise_estimation/mccdr.py:25-192, coherence/BinauralEnhancement.py:24-59), M = 4 ... 8 microphones (M = 4 is the
// only count the reference runs as shipped: mcspp.py:54 builds McCDR with its default channels = 4; above 4 the pin
// is the reference with McCDR(nfft, channels = M) handed in, oracle/ref_harness.make_mcspp).
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

// state blob [S][NE][K] float64; element offsets for M microphones (NQ = M (M - 1) / 2 pairs, row-major i < j)
template <int M> struct CdrLayout {
  static constexpr int NQ = M * (M - 1) / 2, MM = M * M;
  static constexpr int PYY = 0;                 // d[M] ur[NQ] ui[NQ]
  static constexpr int PVV = MM;
  static constexpr int PXII = 2 * MM;           // [M]
  static constexpr int PXIJ_R = PXII + M;       // [NQ]
  static constexpr int PXIJ_I = PXIJ_R + NQ;
  static constexpr int MCRA = PXIJ_I + NQ;      // S Smin Stmp p lambda
  static constexpr int P = MCRA + 5;
  static constexpr int AINV = P + 1;            // outputs of the last frame from here on
  static constexpr int PXX = AINV + MM;
  static constexpr int W_R = PXX + MM, W_I = W_R + M;
  static constexpr int XI = W_I + M, GAMMA = XI + 1, Q = XI + 2, CDR = XI + 3;
  static constexpr int NE = XI + 4;
};
struct CdrOffsets { int M, NQ, PYY, PVV, PXII, MCRA, AINV, PXX, W_R, XI, NE; };
static CdrOffsets cdr_offsets(int M) {
  CdrOffsets o;
  const int NQ = M * (M - 1) / 2, MM = M * M;
  o.M = M; o.NQ = NQ; o.PYY = 0; o.PVV = MM; o.PXII = 2 * MM; o.MCRA = o.PXII + M + 2 * NQ; o.AINV = o.MCRA + 6;
  o.PXX = o.AINV + MM; o.W_R = o.PXX + MM; o.XI = o.W_R + 2 * M; o.NE = o.XI + 4;
  return o;
}

struct CdrArgs {
  double *state;
  const void *X;            // [S][T][M][K]
  int x_c128;
  const double *Fn;         // [K] diffuse coherence of pair (1,2)
  double *q;                // [S][T][K] workspace: prior speech absence probability
  double *qavg;             // [S][T]    workspace
  double *tp, *txi, *tgamma, *tq, *tcdr;     // taps [S][T][K] or null
  double2 *tw;              // [S][T][M][K] or null
  float2 *Yout;             // [S][T][K] or null
  int S, K, T, frm_cnt, ell, lo, hi, init_frames, fallback_frames, repeat;
  double alpha, alpha_d, alpha_cdr, load_min, load_max, snr_min, snr_max, beta, q_init;
  McraConst mc;
};

template <int M>
__device__ __forceinline__ void load_y4(const CdrArgs &a, long long base, int K, double (&yr)[M], double (&yi)[M]) {
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
__device__ __forceinline__ double load_pow0(const CdrArgs &a, long long idx) {
  double re, im;
  if (a.x_c128) { const double2 v = reinterpret_cast<const double2 *>(a.X)[idx]; re = v.x; im = v.y; }
  else { const float2 v = reinterpret_cast<const float2 *>(a.X)[idx]; re = (double)v.x; im = (double)v.y; }
  const double h = hypot(re, im);        // mcra.py:29-30: np.abs(Y) ** 2 for complex input
  return __dmul_rn(h, h);
}

// ---- (1) CDR prior ---------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(128) cdr_prior_kernel(CdrArgs a) {
  typedef CdrLayout<M> LO;
  constexpr int NQ = LO::NQ;
  const int K = a.K;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *blob = a.state + (long long)s * LO::NE * K + k;
  double pii[M], pr[NQ], pi[NQ];
#pragma unroll
  for (int i = 0; i < M; ++i) pii[i] = blob[(long long)(LO::PXII + i) * K];
#pragma unroll
  for (int e = 0; e < NQ; ++e) { pr[e] = blob[(long long)(LO::PXIJ_R + e) * K]; pi[e] = blob[(long long)(LO::PXIJ_I + e) * K]; }
  double mS = blob[(long long)(LO::MCRA + 0) * K], mSmin = blob[(long long)(LO::MCRA + 1) * K], mStmp = blob[(long long)(LO::MCRA + 2) * K],
         mp = blob[(long long)(LO::MCRA + 3) * K], mlam = blob[(long long)(LO::MCRA + 4) * K];
  const double Fn = a.Fn[k], Fn2 = __dmul_rn(Fn, Fn);
  const double al = a.alpha_cdr, om = __dsub_rn(1.0, al);
  int frm = a.frm_cnt, ell = a.ell % a.mc.L;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
  double cdr = 0.0;
  for (int t = 0; t < a.T; ++t) {
    const long long base = ((long long)s * a.T + t) * M * K + k;
    double yr[M], yi[M];
    load_y4<M>(a, base, K, yr, yi);
    // auto / cross spectra, operation order of the reference (no contraction)
#pragma unroll
    for (int i = 0; i < M; ++i)
      pii[i] = __dadd_rn(__dmul_rn(al, pii[i]), __dmul_rn(om, __dadd_rn(__dmul_rn(yr[i], yr[i]), __dmul_rn(yi[i], yi[i]))));
#pragma unroll
    for (int i = 0; i < M - 1; ++i)
#pragma unroll
      for (int j = i + 1; j < M; ++j) {
        const int e = qidx<M>(i, j);
        const double cr = __dadd_rn(__dmul_rn(yr[i], yr[j]), __dmul_rn(yi[i], yi[j]));      // y_i conj(y_j)
        const double ci = __dsub_rn(__dmul_rn(yi[i], yr[j]), __dmul_rn(yr[i], yi[j]));
        pr[e] = __dadd_rn(__dmul_rn(al, pr[e]), __dmul_rn(om, cr));
        pi[e] = __dadd_rn(__dmul_rn(al, pi[e]), __dmul_rn(om, ci));
      }
    // coherence of pair (1,2): complex / real the way NumPy divides (multiply by 1/d)
    constexpr int e12 = qidx<M>(1, 2);
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
    const long long i0 = ((long long)s * a.T + t) * M * K;
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
  for (int i = 0; i < M; ++i) blob[(long long)(LO::PXII + i) * K] = pii[i];
#pragma unroll
  for (int e = 0; e < NQ; ++e) { blob[(long long)(LO::PXIJ_R + e) * K] = pr[e]; blob[(long long)(LO::PXIJ_I + e) * K] = pi[e]; }
  blob[(long long)(LO::MCRA + 0) * K] = mS; blob[(long long)(LO::MCRA + 1) * K] = mSmin; blob[(long long)(LO::MCRA + 2) * K] = mStmp;
  blob[(long long)(LO::MCRA + 3) * K] = mp; blob[(long long)(LO::MCRA + 4) * K] = mlam;
  blob[(long long)LO::CDR * K] = cdr;
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
template <int M> __device__ __forceinline__ void herm_load(Herm<M> &h, const double *blob, int off, int K) {
  constexpr int NQ = Herm<M>::NQ;
#pragma unroll
  for (int i = 0; i < M; ++i) h.d[i] = blob[(long long)(off + i) * K];
#pragma unroll
  for (int e = 0; e < NQ; ++e) { h.ur[e] = blob[(long long)(off + M + e) * K]; h.ui[e] = blob[(long long)(off + M + NQ + e) * K]; }
}
template <int M> __device__ __forceinline__ void herm_store(const Herm<M> &h, double *blob, int off, int K) {
  constexpr int NQ = Herm<M>::NQ;
#pragma unroll
  for (int i = 0; i < M; ++i) blob[(long long)(off + i) * K] = h.d[i];
#pragma unroll
  for (int e = 0; e < NQ; ++e) { blob[(long long)(off + M + e) * K] = h.ur[e]; blob[(long long)(off + M + NQ + e) * K] = h.ui[e]; }
}
// Re tr(A B) for Hermitian A, B
template <int M> __device__ __forceinline__ double herm_trace_prod(const Herm<M> &A, const Herm<M> &B) {
  double d = 0.0, o = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) d = fma(A.d[i], B.d[i], d);
#pragma unroll
  for (int e = 0; e < Herm<M>::NQ; ++e) o = fma(A.ur[e], B.ur[e], fma(A.ui[e], B.ui[e], o));
  return fma(2.0, o, d);
}
// u = A y
template <int M>
__device__ __forceinline__ void herm_matvec(const Herm<M> &A, const double (&yr)[M], const double (&yi)[M], double (&ur)[M], double (&ui)[M]) {
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double sr = A.d[i] * yr[i], si = A.d[i] * yi[i];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (j == i) continue;
      const double ar = (i < j) ? A.ur[qidx<M>(i, j)] : A.ur[qidx<M>(j, i)];
      const double ai = (i < j) ? A.ui[qidx<M>(i, j)] : -A.ui[qidx<M>(j, i)];
      sr = fma(ar, yr[j], fma(-ai, yi[j], sr));
      si = fma(ar, yi[j], fma(ai, yr[j], si));
    }
    ur[i] = sr; ui[i] = si;
  }
}
// Re(u^H B u)
template <int M> __device__ __forceinline__ double herm_quad(const Herm<M> &B, const double (&ur)[M], const double (&ui)[M]) {
  double d = 0.0, o = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) d = fma(B.d[i], fma(ur[i], ur[i], ui[i] * ui[i]), d);
#pragma unroll
  for (int i = 0; i < M - 1; ++i)
#pragma unroll
    for (int j = i + 1; j < M; ++j) {
      // conj(u_i) B_ij u_j + c.c. = 2 Re(conj(u_i) u_j B_ij)
      const int e = qidx<M>(i, j);
      const double cr = fma(ur[i], ur[j], ui[i] * ui[j]);      // conj(u_i) u_j
      const double ci = fma(ur[i], ui[j], -ui[i] * ur[j]);
      o = fma(cr, B.ur[e], fma(-ci, B.ui[e], o));
    }
  return fma(2.0, o, d);
}

// estimation_core (mcspp.py:199-243): inverse with loading and the xi < 0 fallback, xi, gamma, posterior p
template <int M>
__device__ __forceinline__ void cdr_core(const CdrArgs &a, const Herm<M> &Pyy, const Herm<M> &Pvv, Herm<M> &A, double load, int frm,
                                         const double (&yr)[M], const double (&yi)[M], double q, double &p_prev, double &xi, double &gamma) {
  A = Pvv;
#pragma unroll
  for (int i = 0; i < M; ++i) A.d[i] += load;
  herm_inverse<M, true>(A);
  xi = herm_trace_prod<M>(A, Pyy) - (double)M;                                     // :217
  if (xi < 0.0) {                                                                  // :220-228
    A = Pyy;
    if (frm < a.fallback_frames) {
#pragma unroll
      for (int i = 0; i < M; ++i) A.d[i] += load;
    }
    herm_inverse<M, true>(A);
    xi = herm_trace_prod<M>(A, Pyy) - (double)M;
  }
  xi = (xi != xi) ? xi : fmin(fmax(xi, a.snr_min), a.snr_max);                     // :229 (NaN propagates like np.minimum)
  double ur[M], ui[M];
  herm_matvec<M>(A, yr, yi, ur, ui);
  double yAy = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) yAy = fma(yr[i], ur[i], fma(yi[i], ui[i], yAy));
  gamma = herm_quad<M>(Pyy, ur, ui) - yAy;                                         // :231-235
  gamma = (gamma != gamma) ? gamma : fmin(fmax(gamma, a.snr_min), a.snr_max);
  // compute_p(alpha_p = 0)                                                        :76-91
  double p = 1.0 / (1.0 + q / (1.0 - q) * (1.0 + xi) * exp(-1.0 * (gamma / (1.0 + xi))));
  p = 0.0 * p_prev + p;
  p = (p != p) ? p : fmin(fmax(p, 0.0), 1.0);
  p_prev = p;
}

// M > 4: the four packed matrices no longer fit the register file; the compiler keeps what does not fit in local
// memory (L1-resident).  M = 4 -- the reference's own case -- compiles without spills.
template <int M>
__global__ void __launch_bounds__(64) mcspp_cdr_kernel(CdrArgs a) {
  typedef CdrLayout<M> LO;
  constexpr int NQ = LO::NQ;
  const int K = a.K;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *blob = a.state + (long long)s * LO::NE * K + k;
  Herm<M> Pyy, Pvv, A, Pxx;
  herm_load<M>(Pyy, blob, LO::PYY, K);
  herm_load<M>(Pvv, blob, LO::PVV, K);
  double p_prev = blob[(long long)LO::P * K];
  int frm = a.frm_cnt;
  double xi = 0.0, gamma = 0.0, q = 0.0, wr[M], wi[M];
#pragma unroll
  for (int i = 0; i < M; ++i) { wr[i] = 0.0; wi[i] = 0.0; }
  const double om_alpha = 1.0 - a.alpha;
  for (int t = 0; t < a.T; ++t, ++frm) {
    const long long base = ((long long)s * a.T + t) * M * K + k;
    const long long o = ((long long)s * a.T + t) * K + k;
    double yr[M], yi[M];
    load_y4<M>(a, base, K, yr, yi);
    q = a.q[o];
    const double qa = a.qavg[(long long)s * a.T + t];
    const double load = qa * a.load_max + (1.0 - qa) * a.load_min;                 // mcspp.py:269
    // Phi_yy                                                                        :271-273
#pragma unroll
    for (int i = 0; i < M; ++i) Pyy.d[i] = a.alpha * Pyy.d[i] + om_alpha * fma(yr[i], yr[i], yi[i] * yi[i]);
#pragma unroll
    for (int i = 0; i < M - 1; ++i)
#pragma unroll
      for (int j = i + 1; j < M; ++j) {
        const int e = qidx<M>(i, j);
        Pyy.ur[e] = a.alpha * Pyy.ur[e] + om_alpha * fma(yr[i], yr[j], yi[i] * yi[j]);
        Pyy.ui[e] = a.alpha * Pyy.ui[e] + om_alpha * fma(yi[i], yr[j], -yr[i] * yi[j]);
      }
    if (frm < a.init_frames) { Pvv = Pyy; q = a.q_init; }                           // :276-278
    double p = 0.0;
    for (int pass = 0; pass < 1 + a.repeat; ++pass) {
      // Phi_xx = Phi_yy - Phi_vv of the matrices estimation_core sees                 :209
#pragma unroll
      for (int i = 0; i < M; ++i) Pxx.d[i] = Pyy.d[i] - Pvv.d[i];
#pragma unroll
      for (int e = 0; e < NQ; ++e) { Pxx.ur[e] = Pyy.ur[e] - Pvv.ur[e]; Pxx.ui[e] = Pyy.ui[e] - Pvv.ui[e]; }
      cdr_core<M>(a, Pyy, Pvv, A, load, frm, yr, yi, q, p_prev, xi, gamma);
      if (pass == 0) {
        p = p_prev;
        // update_noise_psd(beta = 1) with the first pass's p                           mcspp_base.py:312-319, mcspp.py:281
        const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
        const double om_at = 1.0 * (1.0 - at);
        if (a.repeat) {
          // repeat = True re-runs estimation_core on the UPDATED Phi_vv (:282-284): update first, then second pass
#pragma unroll
          for (int i = 0; i < M; ++i) Pvv.d[i] = at * Pvv.d[i] + om_at * fma(yr[i], yr[i], yi[i] * yi[i]);
#pragma unroll
          for (int i = 0; i < M - 1; ++i)
#pragma unroll
            for (int j = i + 1; j < M; ++j) {
              const int e = qidx<M>(i, j);
              Pvv.ur[e] = at * Pvv.ur[e] + om_at * fma(yr[i], yr[j], yi[i] * yi[j]);
              Pvv.ui[e] = at * Pvv.ui[e] + om_at * fma(yi[i], yr[j], -yr[i] * yi[j]);
            }
        }
      }
    }
    // PMWF weights from the matrices the last estimation_core pass used                :286
    {
      double cr[M], ci[M];      // column 0 of Phi_xx: (Pxx_00, conj(Pxx_01), conj(Pxx_02), ...)
      cr[0] = Pxx.d[0]; ci[0] = 0.0;
#pragma unroll
      for (int i = 1; i < M; ++i) { cr[i] = Pxx.ur[qidx<M>(0, i)]; ci[i] = -Pxx.ui[qidx<M>(0, i)]; }
      herm_matvec<M>(A, cr, ci, wr, wi);
      const double sc = 1.0 / (a.beta + xi);
#pragma unroll
      for (int i = 0; i < M; ++i) { wr[i] *= sc; wi[i] *= sc; }
    }
    if (!a.repeat) {
      // update_noise_psd(beta = 1)                                                    mcspp_base.py:312-319
      const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
      const double om_at = 1.0 * (1.0 - at);
#pragma unroll
      for (int i = 0; i < M; ++i) Pvv.d[i] = at * Pvv.d[i] + om_at * fma(yr[i], yr[i], yi[i] * yi[i]);
#pragma unroll
      for (int i = 0; i < M - 1; ++i)
#pragma unroll
        for (int j = i + 1; j < M; ++j) {
          const int e = qidx<M>(i, j);
          Pvv.ur[e] = at * Pvv.ur[e] + om_at * fma(yr[i], yr[j], yi[i] * yi[j]);
          Pvv.ui[e] = at * Pvv.ui[e] + om_at * fma(yi[i], yr[j], -yr[i] * yi[j]);
        }
    }
    if (a.tp) a.tp[o] = p_prev;
    if (a.txi) a.txi[o] = xi;
    if (a.tgamma) a.tgamma[o] = gamma;
    if (a.tq) a.tq[o] = q;
    if (a.tw) {
#pragma unroll
      for (int i = 0; i < M; ++i) a.tw[base + (long long)i * K] = make_double2(wr[i], wi[i]);
    }
    if (a.Yout) {                               // Y = w^H y
      double Yr = 0.0, Yi = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        Yr = fma(wr[i], yr[i], fma(wi[i], yi[i], Yr));
        Yi = fma(wr[i], yi[i], fma(-wi[i], yr[i], Yi));
      }
      a.Yout[o] = make_float2((float)Yr, (float)Yi);
    }
  }
  herm_store<M>(Pyy, blob, LO::PYY, K);
  herm_store<M>(Pvv, blob, LO::PVV, K);
  herm_store<M>(A, blob, LO::AINV, K);
  herm_store<M>(Pxx, blob, LO::PXX, K);
  blob[(long long)LO::P * K] = p_prev;
#pragma unroll
  for (int i = 0; i < M; ++i) { blob[(long long)(LO::W_R + i) * K] = wr[i]; blob[(long long)(LO::W_I + i) * K] = wi[i]; }
  blob[(long long)LO::XI * K] = xi; blob[(long long)LO::GAMMA * K] = gamma; blob[(long long)LO::Q * K] = q;
}

// ---- export ---------------------------------------------------------------------------------------------
__global__ void cdr_export_herm_kernel(const double *state, int off, int S, int K, int M, int NE, double2 *out) {
  const int MM = M * M, NQ = M * (M - 1) / 2;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * K * MM) return;
  const int ij = (int)(g % MM), k = (int)((g / MM) % K), s = (int)(g / ((long long)MM * K));
  const int i = ij / M, j = ij % M;
  const double *b = state + (long long)s * NE * K + k;
  double re, im = 0.0;
  if (i == j) re = b[(long long)(off + i) * K];
  else {
    const int lo = min(i, j), hi = max(i, j), e = lo * (M - 1) - (lo * (lo - 1)) / 2 + (hi - lo - 1);
    re = b[(long long)(off + M + e) * K];
    im = b[(long long)(off + M + NQ + e) * K];
    if (i > j) im = -im;
  }
  out[g] = make_double2(re, im);
}
__global__ void cdr_export_rows_kernel(const double *state, int off, int n, int S, int K, int NE, double *out) {   // -> [S][n][K]
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * n * K) return;
  const int k = (int)(g % K), e = (int)((g / K) % n), s = (int)(g / ((long long)n * K));
  out[g] = state[((long long)s * NE + off + e) * K + k];
}

template <int M>
static int launch_cdr(const CdrArgs &a, int cdr_only, cudaStream_t st) {
  const long long items = (long long)a.S * a.K;
  cdr_prior_kernel<M><<<(unsigned)((items + 127) / 128), 128, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  if (cdr_only) return DS_OK;
  const long long frames = (long long)a.S * a.T;
  cdr_qavg_kernel<<<(unsigned)((frames + 127) / 128), 128, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  mcspp_cdr_kernel<M><<<(unsigned)((items + 63) / 64), 64, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
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
  if (!p || p->n_mics < 4 || p->n_mics > 8) return 0;
  return (size_t)p->n_streams * cdr_offsets(p->n_mics).NE * (p->n_fft / 2 + 1) * sizeof(double);
}
size_t ds_mcspp_cdr_workspace_bytes(const ds_mcspp_cdr_params *p) {
  if (!p) return 0;
  const size_t K = p->n_fft / 2 + 1, ST = (size_t)p->n_streams * p->n_frames;
  return (ST * K + ST) * sizeof(double);
}

int ds_mcspp_cdr_run(const ds_mcspp_cdr_params *p, void *state, void *workspace, const double *Fn, const void *X, int x_is_c128,
                     void *Yout, const ds_mcspp_cdr_taps *taps, void *stream) {
  DS_CHECK_ARG(p && state && workspace && Fn && X, "ds_mcspp_cdr_run: null argument");
  if (p->n_mics < 4 || p->n_mics > 8) {
    set_error("ds_mcspp_cdr_run: n_mics must be 4..8 (pair (1, 2) of the CDR prior is undefined below 4 in the reference)");
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
  a.repeat = (p->cdr_only & 2) ? 1 : 0;          // bit 1 of cdr_only: McSpp.estimation(repeat=True), mcspp.py:282-284
  a.alpha = p->alpha; a.alpha_d = p->alpha_d; a.alpha_cdr = p->alpha_cdr; a.load_min = p->load_min; a.load_max = p->load_max;
  a.snr_min = p->snr_min; a.snr_max = p->snr_max; a.beta = p->pmwf_beta; a.q_init = p->q_init;
  a.mc.alpha_d = p->mcra_alpha_d; a.mc.alpha_s = p->mcra_alpha_s; a.mc.delta_s = p->mcra_delta_s;
  a.mc.alpha_p = p->mcra_alpha_p; a.mc.p_min = p->mcra_p_min; a.mc.p_max = p->mcra_p_max; a.mc.L = p->mcra_L;
  cudaStream_t st = (cudaStream_t)stream;
  const int cdr_only = p->cdr_only & 1;
  switch (p->n_mics) {
    case 4: return launch_cdr<4>(a, cdr_only, st);
    case 5: return launch_cdr<5>(a, cdr_only, st);
    case 6: return launch_cdr<6>(a, cdr_only, st);
    case 7: return launch_cdr<7>(a, cdr_only, st);
    case 8: return launch_cdr<8>(a, cdr_only, st);
  }
  return DS_EUNSUPPORTED;
}

int ds_mcspp_cdr_export(const ds_mcspp_cdr_params *p, const void *state, int field, void *out, void *stream) {
  DS_CHECK_ARG(p && state && out, "ds_mcspp_cdr_export: null argument");
  DS_CHECK_ARG(p->n_mics >= 4 && p->n_mics <= 8, "ds_mcspp_cdr_export: n_mics must be 4..8");
  const int K = p->n_fft / 2 + 1, S = p->n_streams, M = p->n_mics;
  const CdrOffsets lo = cdr_offsets(M);
  cudaStream_t st = (cudaStream_t)stream;
  const double *sd = (const double *)state;
  if (field >= 0 && field <= 3) {
    const int off = field == 0 ? lo.PYY : field == 1 ? lo.PVV : field == 2 ? lo.AINV : lo.PXX;
    const long long n = (long long)S * K * M * M;
    cdr_export_herm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sd, off, S, K, M, lo.NE, (double2 *)out);
  } else {
    int off, cnt;
    switch (field) {
      case 4: off = lo.W_R; cnt = 2 * M; break;            // [S][2M][K]: re[M], im[M]
      case 5: off = lo.XI; cnt = 4; break;                 // xi gamma q cdr
      case 6: off = lo.PXII; cnt = M + 2 * lo.NQ; break;   // Pxii[M] Pxij_re[NQ] Pxij_im[NQ]
      case 7: off = lo.MCRA; cnt = 6; break;               // S Smin Stmp p lambda | p (posterior)
      default: set_error("ds_mcspp_cdr_export: unknown field %d", field); return DS_EINVAL;
    }
    const long long n = (long long)S * cnt * K;
    cdr_export_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sd, off, cnt, S, K, lo.NE, (double *)out);
  }
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
