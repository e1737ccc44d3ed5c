// chain_step.cuh -- one frame of the McSppBase + MVDR + OMLSA chain for one frequency bin
// (output-only state: packed real parts of Phi_yy / Phi_vv in shared memory).  Shared by the
// stand-alone per-bin kernel (mcspp_fast.cu) and the fused STFT->chain->ISTFT kernel
// (chain_fused.cu).  Reference citations: mcspp_base.py:262-297, :140-155, beamformer.py:133-155.
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

template <int M> __host__ __device__ constexpr int psym(int i, int j) { return i <= j ? pidx<M>(i, j) : pidx<M>(j, i); }

// In-place inverse of an SPD matrix in packed upper storage by symmetric Gauss-Jordan
// sweeps (sweep operator): after sweeping every pivot the array holds -A^-1, which is
// negated on the way out.  Compared with Cholesky (U, U^-1, U^-1 U^-T) the dependent
// chain per pivot is one reciprocal + two FMA levels, and the 28 rank-1 updates of a
// pivot are independent -- this is what the fp64 pipe needs at 2-3 warps per scheduler.
template <int M> __device__ __forceinline__ void spd_inverse_packed(double (&a)[M * (M + 1) / 2]) {
  sfor<0, M>([&](auto kc) {
    constexpr int k = SIDX(kc);
    const double r = rcp_pos(a[pidx<M>(k, k)]);
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
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_f64_once(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

struct McraRegs { double S, Smin, Stmp, p, lam; };

// yf[m]: spectrum of this frame at bin k (complex64), ynb0/ynb1: channel-0 spectrum at k-1 / k+1.
// smy/smv/smc: this thread's Phi_yy / Phi_vv / C columns, element e at [e * NT].
// a0: steering vector of this bin in global memory ((re, im) pairs, mic stride 2K doubles).
// Returns the beamformed (and gained) output bin.
template <int M, int NT, bool USE_C>
__device__ __forceinline__ float2 chain_bin_step(const float2 (&yf)[M], float2 ynb0, float2 ynb1, int k, int K, int frm,
                                                 bool reset, McraRegs &mc, double *smy, double *smv, const double *smc,
                                                 const double *a0, const McsppArgs &a) {
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
      mcra_step(mc.S, mc.Smin, mc.Stmp, mc.p, mc.lam, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
      q = fmin(fmax(sqrt_pos(1.0 - mc.p), a.q_min), a.q_max);
    }
#define AS(i, j) (((i) <= (j)) ? A[pidx<M>(i, j)] : A[pidx<M>(j, i)])

    // ---- MVDR denominator den = a^H A a = sum_{i<=j} A_ij C_ij with the per-bin constants
    //      C_ij = (2 - delta_ij) Re(conj(a_i) a_j) staged in shared memory      beamformer.py:152-153
    // (all long reductions below use several independent accumulators: with two warps per
    //  scheduler the kernel is bound by dependent-issue latency, not by fp64 throughput)
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
    double Yr = 0.0, Yi = 0.0, Yr2 = 0.0, Yi2 = 0.0;   // numerator (A a)^H y = a^H u, accumulated below from u = A y

    // ---- P4: u = A y, Phi_yy update, Xr = Re(Phi_yy - Phi_vv), xi = tr(A Xr), gamma = Re(u^H Xr u)   :84-90,274-284
    // gamma runs on the packed triangle as sum_ij Xr_ij Z_ij with Z_ij = Re(conj(u_i) u_j): one product
    // pair per element instead of one quadratic form per real / imaginary half.
    const double alpha = a.alpha, one_m_alpha = 1.0 - a.alpha;
    double trd[2] = {0.0, 0.0}, tro[4] = {0.0, 0.0, 0.0, 0.0}, gmd[2] = {0.0, 0.0}, gmo[4] = {0.0, 0.0, 0.0, 0.0};
    {
      double ur[M], ui[M];
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
#pragma unroll
      for (int m = 0; m < M; ++m) {          // numerator a^H u
        const double ar = ld_f64_once(a0 + 2 * m * K), ai = ld_f64_once(a0 + 2 * m * K + 1);
        Yr = fma(ar, ur[m], Yr);   Yi = fma(-ai, ur[m], Yi);
        Yr2 = fma(ai, ui[m], Yr2); Yi2 = fma(ar, ui[m], Yi2);
      }
      // every product chain starts at the shared-memory operand, so nothing can be
      // pre-computed (and spilled) ahead of the loads by the instruction scheduler
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const double tr_i = one_m_alpha * yr[i], ti_i = one_m_alpha * yi[i];
#pragma unroll
        for (int j = i; j < M; ++j) {
          const int e = pidx<M>(i, j);
          const double pyy = fma(tr_i, yr[j], fma(ti_i, yi[j], alpha * smy[e * NT]));
          smy[e * NT] = pyy;
          const double x = pyy - smv[e * NT];
          const double z = fma(ui[i], ui[j], ur[i] * ur[j]);
          if (i == j) { trd[i & 1] = fma(A[e], x, trd[i & 1]); gmd[i & 1] = fma(x, z, gmd[i & 1]); }
          else { tro[e & 3] = fma(A[e], x, tro[e & 3]); gmo[e & 3] = fma(x, z, gmo[e & 3]); }
        }
      }
    }
#undef AS
    double xi = fma(2.0, (tro[0] + tro[1]) + (tro[2] + tro[3]), trd[0] + trd[1]);
    double gam = fma(2.0, (gmo[0] + gmo[1]) + (gmo[2] + gmo[3]), gmd[0] + gmd[1]);
    xi = fmin(fmax(xi, a.snr_min), a.snr_max);                               // :286-287
    gam = fmin(fmax(gam, a.snr_min), a.snr_max);

    // ---- P6: posterior SPP                                                   :124-138
    const double xi1 = 1.0 + xi;
    const double rxi1 = rcp_pos(xi1);
    double p = rcp_pos(1.0 + q * rcp_pos(1.0 - q) * xi1 * exp(-1.0 * (gam * rxi1)));
    p = fmin(fmax(p, a.p_min), a.p_max);

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

    // ---- OMLSA gain and output  Y = (w^H y) G,  w = A a / den                :140-155
    double scale = rcp_pos(den);
    if (a.apply_gain) {
      // the gain only scales the output (no feedback into the recursions): fp32 exp/log are enough
      const float pf = (float)p;
      double G = (double)expf(pf * logf((float)(xi * rxi1)) + (1.0f - pf) * (float)a.logGmin);
      G = fmax(fmin(G, 1.0), a.Gmin);
      if (k < 2) G = 0.0;
      scale *= G;
    }
    return make_float2((float)((Yr + Yr2) * scale), (float)((Yi + Yi2) * scale));
}

}  // namespace ds
