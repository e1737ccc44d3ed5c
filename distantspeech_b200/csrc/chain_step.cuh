// chain_step.cuh -- one frame of the McSppBase + MVDR + OMLSA chain for one frequency bin
// (output-only state: packed real parts of Phi_yy / Phi_vv in shared memory), used by the per-bin
// kernel of the headline path (mcspp_fast.cu); gsc.cu shares the sweep inverse and the compile-time
// loop helpers.  Reference citations: mcspp_base.py:262-297, :140-155, beamformer.py:133-155.
#pragma once
#include <type_traits>
#include "mcspp_args.cuh"

namespace ds {

// compile-time loop: f(integral_constant<int, I>) for I in [B, E) -- guarantees that every
// array index below is a constant, so the packed matrices stay in registers
template <int B, int E, typename F> __device__ __forceinline__ void sfor(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    sfor<B + 1, E>(f);
  }
}
#define SIDX(ic) (decltype(ic)::value)

// rendezvous of the two warps that share a scheduler (mcspp_fast.cu): predicated, so the frame body stays one basic block
__device__ __forceinline__ void pair_sync(int id, bool on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p bar.sync %1, 64;\n\t}" ::"r"((unsigned)on), "r"(id));
}

template <int M> __host__ __device__ constexpr int psym(int i, int j) { return i <= j ? pidx<M>(i, j) : pidx<M>(j, i); }

// In-place inverse of an SPD matrix in packed upper storage by symmetric Gauss-Jordan
// sweeps (sweep operator): after sweeping every pivot the array holds -A^-1, which is
// negated on the way out.  Compared with Cholesky (U, U^-1, U^-1 U^-T) the dependent
// chain per pivot is one reciprocal + two FMA levels, and the 28 rank-1 updates of a
// pivot are independent -- this is what the fp64 pipe needs at 2-3 warps per scheduler.
#ifndef DS_CHAIN_RCP1
#define DS_CHAIN_RCP1 0
#endif
#if DS_CHAIN_RCP1
#define CHAIN_RCP rcp_pos_1ulp
#else
#define CHAIN_RCP rcp_pos
#endif
template <int M> __device__ __forceinline__ void spd_inverse_packed(double (&a)[M * (M + 1) / 2]) {
  sfor<0, M>([&](auto kc) {
    constexpr int k = SIDX(kc);
    const double r = CHAIN_RCP(a[pidx<M>(k, k)]);
    double t[M];
    sfor<0, M>([&](auto ic) { constexpr int i = SIDX(ic); if constexpr (i != k) t[i] = a[psym<M>(i, k)] * r; });
    sfor<0, M>([&](auto ic) {
      constexpr int i = SIDX(ic);
      if constexpr (i != k) {
        sfor<i, M>([&](auto jc) {
          constexpr int j = SIDX(jc);
          if constexpr (j != k) a[pidx<M>(i, j)] = fma(-t[i], a[psym<M>(j, k)], a[pidx<M>(i, j)]);
        });
      }
    });
    sfor<0, M>([&](auto ic) { constexpr int i = SIDX(ic); if constexpr (i != k) a[psym<M>(i, k)] = t[i]; });
    a[pidx<M>(k, k)] = -r;
  });
  sfor<0, M * (M + 1) / 2>([&](auto ec) { constexpr int e = SIDX(ec); a[e] = -a[e]; });
}

// non-CSE-able read-only loads: the kernel re-reads small per-bin constants instead of
// keeping them live across phases (registers are the scarce resource here)
__device__ __forceinline__ float2 ld_f2_once(const float2 *p) {
  float2 v;
#ifdef DS_X_NOALLOC
  // the spectrum is streamed once: keep it out of L1 so that the per-bin constants read through ld_f64_once stay there
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
#else
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
#endif
  return v;
}
// one 16-byte load for a (re, im) pair: half the L2 sectors of two strided 8-byte loads
__device__ __forceinline__ double2 ld_f64x2_once(const double *p) {
  double2 v;
#ifdef DS_A0_EVICT_LAST
  asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#else
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#endif
  return v;
}
__device__ __forceinline__ double ld_f64_once(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

struct McraRegs { double S, Smin, Stmp, p, lam; };



// exp(x) for x <= 0 without the library's out-of-range branch: the argument is clamped at -700
// (exp(-700) ~ 1e-304 is already far below anything that can move the SPP), so 2^n stays a normal
// number and the scaling is a plain exponent-field add.  Cody-Waite reduction + degree-13 Taylor
// polynomial on |r| <= ln2/2 (truncation error 4e-18), evaluated as two interleaved Horner chains.
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, -700.0);
  const double magic = 6755399441055744.0;                  // 1.5 * 2^52: round-to-nearest integer in the low word
  const double t = fma(x, 1.4426950408889634, magic);
  const int ni = __double2loint(t);
  const double n = t - magic;
  double r = fma(n, -6.93147180369123816490e-01, x);
  r = fma(n, -1.90821492927058770002e-10, r);
  const double r2 = r * r;
  double pe = fma(1.0 / 479001600.0, r2, 1.0 / 3628800.0);  // even powers: 1/12!, 1/10!, ...
  double po = fma(1.0 / 6227020800.0, r2, 1.0 / 39916800.0); // odd powers: 1/13!, 1/11!, ...
  pe = fma(pe, r2, 1.0 / 40320.0);   po = fma(po, r2, 1.0 / 362880.0);
  pe = fma(pe, r2, 1.0 / 720.0);     po = fma(po, r2, 1.0 / 5040.0);
  pe = fma(pe, r2, 1.0 / 24.0);      po = fma(po, r2, 1.0 / 120.0);
  pe = fma(pe, r2, 0.5);             po = fma(po, r2, 1.0 / 6.0);
  pe = fma(pe, r2, 1.0);             po = fma(po, r2, 1.0);
  const double v = fma(po, r, pe);
  return __hiloint2double(__double2hiint(v) + (ni << 20), __double2loint(v));
}

// mcra_step (perbin.cuh) with every case distinction turned into a select: the same operations in
// the same order on the general path (bit-identical decisions), but one basic block, so that the
// serial S -> S/(Smin + 1e-6) -> p chain can be interleaved with the matrix work around it.
__device__ __forceinline__ void mcra_step_sel(McraRegs &m, double Ym1, double Y0, double Yp1, int k, int K, int frm_cnt,
                                              bool reset, const McraConst &c) {
  const bool tracked = k < K - 1, first = frm_cnt == 0;
  const bool general = tracked && !first && k > 0;
  const double Sf = __dadd_rn(__dadd_rn(__dmul_rn(Ym1, 0.25), __dmul_rn(Y0, 0.5)), __dmul_rn(Yp1, 0.25));   // mcra.py:46
  const double Sn = __dadd_rn(__dmul_rn(c.alpha_s, m.S), __dmul_rn(__dsub_rn(1.0, c.alpha_s), Sf));        // :47
  const double Stmp1 = fmin(m.Stmp, Sn);                                                                   // :49-50
  const double Smin_g = reset ? Stmp1 : fmin(m.Smin, Sn);                                                  // :52-56
  const double Stmp_g = reset ? Sn : Stmp1;
  const double Sr = div_rn_fast(Sn, __dadd_rn(Smin_g, 1e-6));                                              // :58
  const double I = (Sr > c.delta_s) ? 1.0 : 0.0;
  const double pg = __dadd_rn(__dmul_rn(c.alpha_p, m.p), __dmul_rn(__dsub_rn(1.0, c.alpha_p), I));         // :65-67
  const bool warm = frm_cnt < 2 * c.L;                                                                     // :68-69
  const bool p_zero = tracked && (first ? warm : (k == 0 || warm));
  const bool p_keep = !tracked || (first && !warm);
  double p = p_keep ? m.p : pg;
  p = p_zero ? 0.0 : p;
  m.S = general ? Sn : m.S;
  m.Smin = general ? Smin_g : ((tracked && first) ? Y0 : m.Smin);
  m.Stmp = general ? Stmp_g : ((tracked && first) ? Y0 : m.Stmp);
  double lam = (tracked && first) ? Y0 : m.lam;
  p = fmax(fmin(p, c.p_max), c.p_min);                                                                     // :70
  if (!tracked) lam = 1e-8;                                                                                // :73
  const double at = __dadd_rn(c.alpha_d, __dmul_rn(__dsub_rn(1.0, c.alpha_d), p));                         // Base :57
  m.lam = __dadd_rn(__dmul_rn(at, lam), __dmul_rn(__dmul_rn(1.0, __dsub_rn(1.0, at)), Y0));                // Base :60
  m.p = p;
}

// yf[m]: spectrum of this frame at bin k (complex64), ynb0/ynb1: channel-0 spectrum at k-1 / k+1.
// smy/smv/smc: this thread's Phi_yy / Phi_vv / C columns, element e at [e * NT].
// a0: steering vector of this bin in global memory ((re, im) pairs, mic stride 2K doubles).
// Returns the beamformed (and gained) output bin; p_post receives the posterior speech-presence probability.
//
// xi and gamma are evaluated through A (Phi_vv + eps I) = I, which takes Phi_vv out of both forms:
//   xi    = tr(A (Phi_yy' - Phi_vv))        = tr(A Phi_yy') - M + eps tr(A)
//   gamma = u^H (Phi_yy' - Phi_vv) u, u=A y = sum_ij Phi_yy'_ij Re(conj(u_i) u_j) - Re(y^H u) + eps |u|^2
// (measured against the direct forms over the synthetic streams and an ill-conditioned two-source
// mixture, cond 1e8: |dp| <= 2.2e-8).  The point is the schedule, not the flop count: pass X needs
// A but not u, pass Z needs u but not A, the updated Phi_yy' takes over A's registers element by
// element, and Phi_vv is read once per frame less -- the fused loop this replaces had ~200 live
// registers and ran its nine-operation chain per element almost serially (8-cycle DFMA latency).
// MIXED (A/B experiment, compiled in only with -DDS_CHAIN_MIXED): u = A y, tr(A Phi_yy') and the quadratic form of gamma on
// the otherwise idle fp32 pipe from float copies of A and Phi_yy' (288 fewer fp64 instructions per bin and frame); the sweep
// inverse, both covariance recursions, the MCRA / SPP chain and the MVDR numerator / denominator stay in fp64.  MEASURED
// (round 2, B200, 1024 streams x 10 s): 22.67 ms per step against 22.25 ms -- SLOWER: the 88 float<->double conversions
// it needs cost more than the 288 DFMA it saves -- and less accurate: 100 dB instead of 108 dB on the synthetic streams,
// 46 dB (below the 60 dB contract) when Phi_vv is close to singular, because xi and gamma are differences of nearly equal
// sums (tools/mixed_precision_study.py).  Kept only so the measurement can be repeated.
template <int M, int NT, bool USE_C, bool MIXED = false, bool PAIRED = false>
__device__ __forceinline__ float2 chain_bin_step(const float2 (&yf)[M], float2 ynb0, float2 ynb1, int k, int K, int frm,
                                                 bool reset, McraRegs &mc, double *smy, double *smv, const double *smc,
                                                 const double *a0, const McsppArgs &a, double &p_post, int mid_bar = 0) {
  constexpr int NP = M * (M + 1) / 2;
    // ---- P1: A = inv(Re Phi_vv + eps I)                                     mcspp_base.py:278
    double A[NP];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) A[pidx<M>(i, j)] = smv[pidx<M>(i, j) * NT] + ((i == j) ? a.eps : 0.0);
    spd_inverse_packed<M>(A);

    // ---- prior from MCRA on channel 0                                       :98-122
    double yr[M], yi[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { yr[m] = (double)yf[m].x; yi[m] = (double)yf[m].y; }
    double q;
    {
      const double Ym1 = (k > 0) ? power_c((double)ynb0.x, (double)ynb0.y) : 0.0;
      const double Yp1 = (k < K - 1) ? power_c((double)ynb1.x, (double)ynb1.y) : 0.0;
      const double Y0 = power_c(yr[0], yi[0]);
      mcra_step_sel(mc, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
      const double omp = fmax(1.0 - mc.p, 1e-300);                           // p <= p_max < 1: never taken
      const double rs = rsqrt_pos(omp);                                      // sqrt_pos without its x <= 0 branch
      double sq = omp * rs;
      sq = fma(fma(-sq, sq, omp), 0.5 * rs, sq);
      q = fmin(fmax(sq, a.q_min), a.q_max);
    }
#define AS(i, j) (((i) <= (j)) ? A[pidx<M>(i, j)] : A[pidx<M>(j, i)])

    // ---- MVDR denominator den = a^H A a = sum_{i<=j} A_ij C_ij with the per-bin constants
    //      C_ij = (2 - delta_ij) Re(conj(a_i) a_j) staged in shared memory      beamformer.py:152-153
    double den4[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (USE_C) {
#pragma unroll
      for (int e = 0; e < NP; ++e) den4[e & 3] = fma(A[e], smc[e * NT], den4[e & 3]);
    } else {
      // no room for the constants in shared memory: rebuild C_ij from the steering vector
      double ar[M], ai[M];
#pragma unroll
      for (int m = 0; m < M; ++m) { ar[m] = ld_f64_once(a0 + 2 * m * K); ai[m] = ld_f64_once(a0 + 2 * m * K + 1); }
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i; j < M; ++j) {
          const int e = pidx<M>(i, j);
          const double c = fma(ai[i], ai[j], ar[i] * ar[j]);
          den4[(i == j) ? (i & 1) : 2 + (e & 1)] = fma(A[e], c, den4[(i == j) ? (i & 1) : 2 + (e & 1)]);
        }
      den4[2] *= 2.0; den4[3] *= 2.0;
    }
    const double den = (den4[0] + den4[1]) + (den4[2] + den4[3]);
    double trA = 0.0, trA2 = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) { if (i & 1) trA2 += A[pidx<M>(i, i)]; else trA += A[pidx<M>(i, i)]; }
    trA += trA2;
    // rendezvous point of the second warp of a scheduler pair (mcspp_fast.cu; mid_bar = 0: none).  Measured positions
    // (per-bin ms, unpaired 16.87): in the middle of the sweep 16.82, after the sweep 16.87, here 16.49 (16.43 in the final form),
    // after u = A y 16.49, after pass X 17.46, after pass Z 17.50, after the SPP chain 17.99, before the noise update 17.67
    if constexpr (PAIRED) pair_sync(mid_bar, mid_bar != 0);

    // ---- u = A y, numerator a^H u, s_yu = Re(y^H u) = y^H A y, uu = |u|^2     :282-284, beamformer.py:152
    double ur[M], ui[M];
#ifdef DS_HOIST_A0
    double a0r[M], a0i[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { a0r[m] = ld_f64_once(a0 + 2 * m * K); a0i[m] = ld_f64_once(a0 + 2 * m * K + 1); }
#endif
    float Af[MIXED ? NP : 1], urf[MIXED ? M : 1], uif[MIXED ? M : 1];
    if constexpr (MIXED) {
#pragma unroll
      for (int e = 0; e < NP; ++e) Af[e] = (float)A[e];
#define ASF(i, j) (((i) <= (j)) ? Af[pidx<M>(i, j)] : Af[pidx<M>(j, i)])
#pragma unroll
      for (int i = 0; i < M; ++i) {
        float sr = 0.f, sr2 = 0.f, si = 0.f, si2 = 0.f;
#pragma unroll
        for (int j = 0; j < M; ++j) {
          if (j & 1) { sr2 = fmaf(ASF(i, j), yf[j].x, sr2); si2 = fmaf(ASF(i, j), yf[j].y, si2); }
          else { sr = fmaf(ASF(i, j), yf[j].x, sr); si = fmaf(ASF(i, j), yf[j].y, si); }
        }
        urf[i] = sr + sr2; uif[i] = si + si2;
        ur[i] = (double)urf[i]; ui[i] = (double)uif[i];
      }
#undef ASF
    } else {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double sr = 0.0, sr2 = 0.0, si = 0.0, si2 = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        if (j & 1) { sr2 = fma(AS(i, j), yr[j], sr2); si2 = fma(AS(i, j), yi[j], si2); }
        else { sr = fma(AS(i, j), yr[j], sr); si = fma(AS(i, j), yi[j], si); }
      }
      ur[i] = sr + sr2;
      ui[i] = si + si2;
    }
    }
    double Yr = 0.0, Yi = 0.0, Yr2 = 0.0, Yi2 = 0.0, syu = 0.0, syu2 = 0.0, uu = 0.0, uu2 = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) {
#ifdef DS_HOIST_A0
      const double ar = a0r[m], ai = a0i[m];
#else
      const double2 av = ld_f64x2_once(a0 + 2 * m * K);
      const double ar = av.x, ai = av.y;
#endif
      Yr = fma(ar, ur[m], Yr);   Yi = fma(-ai, ur[m], Yi);
      Yr2 = fma(ai, ui[m], Yr2); Yi2 = fma(ar, ui[m], Yi2);
      syu = fma(yr[m], ur[m], syu); syu2 = fma(yi[m], ui[m], syu2);
      uu = fma(ur[m], ur[m], uu);   uu2 = fma(ui[m], ui[m], uu2);
    }

    // ---- pass X: Phi_yy' = alpha Phi_yy + (1 - alpha) Re(y y^H), tr(A Phi_yy'); Phi_yy' replaces A   :84-90, :280
    const double alpha = a.alpha, one_m_alpha = 1.0 - a.alpha;
    double trd[2] = {0.0, 0.0}, tro[4] = {0.0, 0.0, 0.0, 0.0};
    float trdf[2] = {0.f, 0.f}, trof[4] = {0.f, 0.f, 0.f, 0.f}, gmdf[2] = {0.f, 0.f}, gmof[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const double tr_i = one_m_alpha * yr[i], ti_i = one_m_alpha * yi[i];
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double pyy = fma(tr_i, yr[j], fma(ti_i, yi[j], alpha * smy[e * NT]));
        smy[e * NT] = pyy;
        if constexpr (MIXED) {
          // trace term and (pass Z fused here) the quadratic-form term of gamma, both on the fp32 pipe
          const float pf32 = (float)pyy;
          const float z = fmaf(uif[i], uif[j], urf[i] * urf[j]);
          if (i == j) { trdf[i & 1] = fmaf(Af[e], pf32, trdf[i & 1]); gmdf[i & 1] = fmaf(pf32, z, gmdf[i & 1]); }
          else { trof[e & 3] = fmaf(Af[e], pf32, trof[e & 3]); gmof[e & 3] = fmaf(pf32, z, gmof[e & 3]); }
        } else {
          if (i == j) trd[i & 1] = fma(A[e], pyy, trd[i & 1]);
          else tro[e & 3] = fma(A[e], pyy, tro[e & 3]);
          A[e] = pyy;
        }
      }
    }
#undef AS
    // ---- pass Z: sum_ij Phi_yy'_ij Re(conj(u_i) u_j)
    double gmd[2] = {0.0, 0.0}, gmo[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (MIXED) {
      trd[0] = (double)(trdf[0] + trdf[1]); trd[1] = 0.0;
      tro[0] = (double)((trof[0] + trof[1]) + (trof[2] + trof[3])); tro[1] = tro[2] = tro[3] = 0.0;
      gmd[0] = (double)(gmdf[0] + gmdf[1]);
      gmo[0] = (double)((gmof[0] + gmof[1]) + (gmof[2] + gmof[3]));
    } else {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double z = fma(ui[i], ui[j], ur[i] * ur[j]);
        if (i == j) gmd[i & 1] = fma(A[e], z, gmd[i & 1]);
        else gmo[e & 3] = fma(A[e], z, gmo[e & 3]);
      }
    }
    double xi = fma(2.0, (tro[0] + tro[1]) + (tro[2] + tro[3]), trd[0] + trd[1]);
    xi = fma(a.eps, trA, xi - (double)M);
    double gam = fma(2.0, (gmo[0] + gmo[1]) + (gmo[2] + gmo[3]), gmd[0] + gmd[1]);
    gam = fma(a.eps, uu + uu2, gam - (syu + syu2));
    xi = fmin(fmax(xi, a.snr_min), a.snr_max);                               // :286-287
    gam = fmin(fmax(gam, a.snr_min), a.snr_max);

    // ---- P6: posterior SPP                                                   :124-138
    const double xi1 = 1.0 + xi;
    const double rxi1 = CHAIN_RCP(xi1);
    double p = CHAIN_RCP(1.0 + q * CHAIN_RCP(1.0 - q) * xi1 * exp_nonpos(-1.0 * (gam * rxi1)));
    p = fmin(fmax(p, a.p_min), a.p_max);
    p_post = p;

    // ---- OMLSA gain and output  Y = (w^H y) G,  w = A a / den                :140-155
    // (ahead of the noise-PSD update in program order: its serial fp32 log/exp chain then overlaps the update's
    //  shared-memory traffic instead of trailing it)
    double scale = CHAIN_RCP(den);
    if (a.apply_gain) {
      // the gain only scales the output (no feedback into the recursions): fp32 exp/log are enough
      const float pf = (float)p;
      double G = (double)expf(pf * logf((float)(xi * rxi1)) + (1.0f - pf) * (float)a.logGmin);
      G = fmax(fmin(G, 1.0), a.Gmin);
      if (k < 2) G = 0.0;
      scale *= G;
    }
    const float2 yout = make_float2((float)((Yr + Yr2) * scale), (float)((Yi + Yi2) * scale));

    // ---- noise PSD update                                                    :299-319
    const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
    const double one_m_at = 1.0 * (1.0 - at);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const double tr_i = one_m_at * yr[i], ti_i = one_m_at * yi[i];
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        smv[e * NT] = fma(tr_i, yr[j], fma(ti_i, yi[j], at * smv[e * NT]));
      }
    }

    return yout;
}

// ---- version 2 of the frame step: the same quantities, restructured for register pressure -------------------------------
// Phi_yy' = alpha Phi_yy + (1 - alpha) Re(y y^H) enters xi and gamma linearly, so both forms are evaluated against the OLD
// Phi_yy (read from shared memory) and the rank-two part in closed form:
//   tr(A Phi_yy')              = alpha tr(A Phi_yy) + (1 - alpha) Re(y^H A y)                       (Re(y^H A y) = s_yu, already there)
//   sum_ij Phi_yy'_ij z_ij     = alpha sum_ij Phi_yy_ij z_ij + (1 - alpha) (d1^2 + d2^2 + d3^2 + d4^2),  z_ij = Re(conj(u_i) u_j),
//                                d1 = yr.ur, d2 = yi.ui, d3 = yr.ui, d4 = yi.ur
// One pass over the matrix (4 fp64 operations per element) then needs A and u but NOT y, A dies right after it, and the
// Phi_yy recursion itself becomes a streaming pass (load, 3 operations, store) that depends on nothing adaptive -- the
// scheduler can sink it under the serial speech-presence chain.  u = A y is formed in two halves (real, imaginary part) so
// that y is held in double precision for one half at a time.  +24 fp64 instructions per bin and frame against version 1,
// in exchange for ~40 fewer live registers in the passes that used to hold A, y and u together.
template <int M, int NT, bool USE_C>
__device__ __forceinline__ float2 chain_bin_step_v2(const float2 (&yf)[M], float2 ynb0, float2 ynb1, int k, int K, int frm,
                                                    bool reset, McraRegs &mc, double *smy, double *smv, const double *smc,
                                                    const double *a0, const McsppArgs &a, double &p_post) {
  constexpr int NP = M * (M + 1) / 2;
    // ---- A = inv(Re Phi_vv + eps I)                                         mcspp_base.py:278
    double A[NP];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) A[pidx<M>(i, j)] = smv[pidx<M>(i, j) * NT] + ((i == j) ? a.eps : 0.0);
    spd_inverse_packed<M>(A);

    // ---- prior from MCRA on channel 0                                       :98-122
    double q;
    {
      const double Ym1 = (k > 0) ? power_c((double)ynb0.x, (double)ynb0.y) : 0.0;
      const double Yp1 = (k < K - 1) ? power_c((double)ynb1.x, (double)ynb1.y) : 0.0;
      const double Y0 = power_c((double)yf[0].x, (double)yf[0].y);
      mcra_step_sel(mc, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
      const double omp = fmax(1.0 - mc.p, 1e-300);                           // p <= p_max < 1: never taken
      const double rs = rsqrt_pos(omp);
      double sq = omp * rs;
      sq = fma(fma(-sq, sq, omp), 0.5 * rs, sq);
      q = fmin(fmax(sq, a.q_min), a.q_max);
    }
#define AS(i, j) (((i) <= (j)) ? A[pidx<M>(i, j)] : A[pidx<M>(j, i)])
    // ---- MVDR denominator den = a^H A a = sum_{i<=j} A_ij C_ij                beamformer.py:152-153
    double den4[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (USE_C) {
#pragma unroll
      for (int e = 0; e < NP; ++e) den4[e & 3] = fma(A[e], smc[e * NT], den4[e & 3]);
    } else {
      double ar[M], ai[M];
#pragma unroll
      for (int m = 0; m < M; ++m) { const double2 av = ld_f64x2_once(a0 + 2 * m * K); ar[m] = av.x; ai[m] = av.y; }
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = i; j < M; ++j) {
          const int e = pidx<M>(i, j);
          const double c = fma(ai[i], ai[j], ar[i] * ar[j]);
          den4[(i == j) ? (i & 1) : 2 + (e & 1)] = fma(A[e], c, den4[(i == j) ? (i & 1) : 2 + (e & 1)]);
        }
      den4[2] *= 2.0; den4[3] *= 2.0;
    }
    const double den = (den4[0] + den4[1]) + (den4[2] + den4[3]);
    double trA = 0.0, trA2 = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) { if (i & 1) trA2 += A[pidx<M>(i, i)]; else trA += A[pidx<M>(i, i)]; }
    trA += trA2;

    // ---- u = A y in two halves; numerator a^H u, s_yu = Re(y^H u), |u|^2, the four dot products      :282-284, beamformer.py:152
    double ur[M], ui[M];
    double Yr = 0.0, Yi = 0.0, Yr2 = 0.0, Yi2 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0, d4 = 0.0, uu = 0.0, uu2 = 0.0;
    {
      double yr[M];
#pragma unroll
      for (int m = 0; m < M; ++m) yr[m] = (double)yf[m].x;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double sr = 0.0, sr2 = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) { if (j & 1) sr2 = fma(AS(i, j), yr[j], sr2); else sr = fma(AS(i, j), yr[j], sr); }
        ur[i] = sr + sr2;
      }
      double yi[M];
#pragma unroll
      for (int m = 0; m < M; ++m) yi[m] = (double)yf[m].y;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double si = 0.0, si2 = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) { if (j & 1) si2 = fma(AS(i, j), yi[j], si2); else si = fma(AS(i, j), yi[j], si); }
        ui[i] = si + si2;
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double2 av = ld_f64x2_once(a0 + 2 * m * K);
        Yr = fma(av.x, ur[m], Yr);   Yi = fma(-av.y, ur[m], Yi);
        Yr2 = fma(av.y, ui[m], Yr2); Yi2 = fma(av.x, ui[m], Yi2);
        d1 = fma(yr[m], ur[m], d1);  d2 = fma(yi[m], ui[m], d2);
        d3 = fma(yr[m], ui[m], d3);  d4 = fma(yi[m], ur[m], d4);
        uu = fma(ur[m], ur[m], uu);  uu2 = fma(ui[m], ui[m], uu2);
      }
    }
    const double syu = d1 + d2;

    // ---- one pass over the OLD Phi_yy: tr(A Phi_yy) and sum_ij Phi_yy_ij Re(conj(u_i) u_j); A is dead afterwards   :280-284
    double trd[2] = {0.0, 0.0}, tro[4] = {0.0, 0.0, 0.0, 0.0}, gmd[2] = {0.0, 0.0}, gmo[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        const double ph = smy[e * NT];
        const double z = fma(ui[i], ui[j], ur[i] * ur[j]);
        if (i == j) { trd[i & 1] = fma(A[e], ph, trd[i & 1]); gmd[i & 1] = fma(ph, z, gmd[i & 1]); }
        else { tro[e & 3] = fma(A[e], ph, tro[e & 3]); gmo[e & 3] = fma(ph, z, gmo[e & 3]); }
      }
#undef AS
    const double alpha = a.alpha, one_m_alpha = 1.0 - a.alpha;
    double xi = fma(2.0, (tro[0] + tro[1]) + (tro[2] + tro[3]), trd[0] + trd[1]);           // tr(A Phi_yy_old)
    xi = fma(alpha, xi, one_m_alpha * syu);                                                 // tr(A Phi_yy')
    xi = fma(a.eps, trA, xi - (double)M);
    double gam = fma(2.0, (gmo[0] + gmo[1]) + (gmo[2] + gmo[3]), gmd[0] + gmd[1]);
    gam = fma(alpha, gam, one_m_alpha * (fma(d1, d1, d2 * d2) + fma(d3, d3, d4 * d4)));
    gam = fma(a.eps, uu + uu2, gam - syu);
    xi = fmin(fmax(xi, a.snr_min), a.snr_max);                               // :286-287
    gam = fmin(fmax(gam, a.snr_min), a.snr_max);

    // ---- posterior SPP                                                       :124-138
    const double xi1 = 1.0 + xi;
    const double rxi1 = CHAIN_RCP(xi1);
    double p = CHAIN_RCP(1.0 + q * CHAIN_RCP(1.0 - q) * xi1 * exp_nonpos(-1.0 * (gam * rxi1)));
    p = fmin(fmax(p, a.p_min), a.p_max);
    p_post = p;

    // ---- OMLSA gain and output  Y = (w^H y) G,  w = A a / den                :140-155
    double scale = CHAIN_RCP(den);
    if (a.apply_gain) {
      const float pf = (float)p;
      double G = (double)expf(pf * logf((float)(xi * rxi1)) + (1.0f - pf) * (float)a.logGmin);
      G = fmax(fmin(G, 1.0), a.Gmin);
      if (k < 2) G = 0.0;
      scale *= G;
    }
    const float2 yout = make_float2((float)((Yr + Yr2) * scale), (float)((Yi + Yi2) * scale));

    // ---- both covariance recursions as streaming passes (Phi_yy' depends on nothing adaptive)     :84-90, :299-319
    const double at = a.alpha_d + (1.0 - a.alpha_d) * p;
    const double one_m_at = 1.0 * (1.0 - at);
    double yr[M], yi[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { yr[m] = (double)yf[m].x; yi[m] = (double)yf[m].y; }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const double tr_i = one_m_alpha * yr[i], ti_i = one_m_alpha * yi[i];
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        smy[e * NT] = fma(tr_i, yr[j], fma(ti_i, yi[j], alpha * smy[e * NT]));
      }
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const double tr_i = one_m_at * yr[i], ti_i = one_m_at * yi[i];
#pragma unroll
      for (int j = i; j < M; ++j) {
        const int e = pidx<M>(i, j);
        smv[e * NT] = fma(tr_i, yr[j], fma(ti_i, yi[j], at * smv[e * NT]));
      }
    }
    return yout;
}

}  // namespace ds
