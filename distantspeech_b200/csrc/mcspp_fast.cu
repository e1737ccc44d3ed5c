// mcspp_fast.cu -- hot-path variant of the McSppBase + MVDR + OMLSA per-bin
// kernel (config 4): output-only state, complex64 spectrum in, no taps.
//
// Same arithmetic as mcspp_kernel (mcspp.cu) -- see the reference citations
// there -- restructured for the fp64 pipe:
//   * every quadratic form is evaluated against the packed upper triangle
//     (xi = sum A_ij X_ij, gamma = sum X_ij Re(conj(u_i) u_j),
//      a^H A a = sum A_ij Re(conj(a_i) a_j)), so no M-vector temporaries
//     besides u = A y survive the inverse;
//   * the live set stays under 168 registers => 3 CTAs x 128 threads per SM
//     next to 3 x 72 KB of shared-memory state, no local-memory spills.
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

template <int M, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) mcspp_fast_kernel(McsppArgs a) {
  constexpr int NP = M * (M + 1) / 2;
  constexpr int NE = mcspp_state_elems<M>();
  constexpr int OFF_YR = 0, OFF_VR = NP, OFF_MC = 2 * NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  const int K = a.K;
  const int Kp = K - a.k_first;
  const long long g = (long long)blockIdx.x * NT + tid;
  if (g >= (long long)a.S * Kp) return;          // threads are independent: no block-wide sync below
  const int s = (int)(g / Kp), k = a.k_first + (int)(g % Kp);
  double *blob = a.state + (long long)s * NE * K + k;
  double *smy = sm + tid;                 // Phi_yy (real part), element e at smy[e * NT]
  double *smv = sm + NP * NT + tid;       // Phi_vv (real part)
  double *smc = sm + 2 * NP * NT + tid;   // (2 - delta_ij) Re(conj(a_i) a_j), constant per bin
#pragma unroll
  for (int e = 0; e < NP; ++e) { smy[e * NT] = blob[(long long)(OFF_YR + e) * K]; smv[e * NT] = blob[(long long)(OFF_VR + e) * K]; }
  double mS = blob[(long long)(OFF_MC + 0) * K], mSmin = blob[(long long)(OFF_MC + 1) * K], mStmp = blob[(long long)(OFF_MC + 2) * K],
         mp = blob[(long long)(OFF_MC + 3) * K], mlam = blob[(long long)(OFF_MC + 4) * K];

  {
    double ar[M], ai[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { const double2 v = a.a0[(long long)m * K + k]; ar[m] = v.x; ai[m] = v.y; }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j)
        smc[pidx<M>(i, j) * NT] = ((i == j) ? 1.0 : 2.0) * fma(ai[i], ai[j], ar[i] * ar[j]);
  }
  int frm = a.frm_cnt, ell = a.ell;
  const float2 *Xp = reinterpret_cast<const float2 *>(a.X) + (long long)s * a.T * M * K + k;
  float2 *Yp = a.Yout + (long long)s * a.T * K + k;
  const double *a0 = reinterpret_cast<const double *>(a.a0 + k);     // (re, im) pairs, mic stride 2K doubles

#pragma unroll 1
  for (int t = 0; t < a.T; ++t) {
    // ---- P0: issue the loads of this frame's spectrum (kept as float until the inverse is done,
    //          so the HBM latency hides behind the Gauss-Jordan sweeps instead of stalling the warp)
    float2 yf[M], ynb[2];
#pragma unroll
    for (int m = 0; m < M; ++m) yf[m] = ld_f2_once(Xp + m * K);
    ynb[0] = (k > 0) ? ld_f2_once(Xp - 1) : make_float2(0.f, 0.f);
    ynb[1] = (k < K - 1) ? ld_f2_once(Xp + 1) : make_float2(0.f, 0.f);
    Xp += M * K;

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
      const double Ym1 = (k > 0) ? power_c((double)ynb[0].x, (double)ynb[0].y) : 0.0;
      const double Yp1 = (k < K - 1) ? power_c((double)ynb[1].x, (double)ynb[1].y) : 0.0;
      const double Y0 = power_c(yr[0], yi[0]);
      const bool reset = (frm > 0) && (ell % a.mc.L == 0);
      mcra_step(mS, mSmin, mStmp, mp, mlam, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
      if (reset) ell = 0;
      ++ell; ++frm;
      q = fmin(fmax(sqrt_pos(1.0 - mp), a.q_min), a.q_max);
    }
#define AS(i, j) (((i) <= (j)) ? A[pidx<M>(i, j)] : A[pidx<M>(j, i)])

    // ---- MVDR denominator den = a^H A a = sum_{i<=j} A_ij C_ij with the per-bin constants
    //      C_ij = (2 - delta_ij) Re(conj(a_i) a_j) staged in shared memory      beamformer.py:152-153
    // (all long reductions below use several independent accumulators: with two warps per
    //  scheduler the kernel is bound by dependent-issue latency, not by fp64 throughput)
    double den4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int e = 0; e < NP; ++e) den4[e & 3] = fma(A[e], smc[e * NT], den4[e & 3]);
    const double den = (den4[0] + den4[1]) + (den4[2] + den4[3]);
    double Yr = 0.0, Yi = 0.0, Yr2 = 0.0, Yi2 = 0.0;   // numerator (A a)^H y = a^H u, accumulated below from u = A y

    // ---- P4: Phi_yy update, Xr = Re(Phi_yy - Phi_vv), xi = tr(A Xr), real half of gamma   :84-90,274-284
    const double alpha = a.alpha, one_m_alpha = 1.0 - a.alpha;
    double trd[2] = {0.0, 0.0}, tro[4] = {0.0, 0.0, 0.0, 0.0}, gmd[2] = {0.0, 0.0}, gmo[4] = {0.0, 0.0, 0.0, 0.0};
    {
      double ur[M];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double sr = 0.0, sr2 = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) { if (j & 1) sr2 = fma(AS(i, j), yr[j], sr2); else sr = fma(AS(i, j), yr[j], sr); }
        ur[i] = sr + sr2;
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {          // conj(a) * u, real half of u
        Yr = fma(ld_f64_once(a0 + 2 * m * K), ur[m], Yr);
        Yi = fma(-ld_f64_once(a0 + 2 * m * K + 1), ur[m], Yi);
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
          if (i == j) { trd[i & 1] = fma(A[e], x, trd[i & 1]); gmd[i & 1] = fma(x * ur[i], ur[j], gmd[i & 1]); }
          else { tro[e & 3] = fma(A[e], x, tro[e & 3]); gmo[e & 3] = fma(x * ur[i], ur[j], gmo[e & 3]); }
        }
      }
    }
    // ---- P5: imaginary half of gamma = Re(u^H Xr u), u = A y
    {
      double ui[M];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double si = 0.0, si2 = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) { if (j & 1) si2 = fma(AS(i, j), yi[j], si2); else si = fma(AS(i, j), yi[j], si); }
        ui[i] = si + si2;
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {          // conj(a) * u, imaginary half of u
        Yr2 = fma(ld_f64_once(a0 + 2 * m * K + 1), ui[m], Yr2);
        Yi2 = fma(ld_f64_once(a0 + 2 * m * K), ui[m], Yi2);
      }
#pragma unroll
      for (int i = 0; i < M; ++i) {
#pragma unroll
        for (int j = i; j < M; ++j) {
          const int e = pidx<M>(i, j);
          const double x = smy[e * NT] - smv[e * NT];
          if (i == j) gmd[i & 1] = fma(x * ui[i], ui[j], gmd[i & 1]);
          else gmo[e & 3] = fma(x * ui[i], ui[j], gmo[e & 3]);
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
    *Yp = make_float2((float)((Yr + Yr2) * scale), (float)((Yi + Yi2) * scale));
    if (a.k_first == 2 && k == 2) { Yp[-1] = make_float2(0.f, 0.f); Yp[-2] = make_float2(0.f, 0.f); }
    Yp += K;
  }

#pragma unroll
  for (int e = 0; e < NP; ++e) { blob[(long long)(OFF_YR + e) * K] = smy[e * NT]; blob[(long long)(OFF_VR + e) * K] = smv[e * NT]; }
  blob[(long long)(OFF_MC + 0) * K] = mS; blob[(long long)(OFF_MC + 1) * K] = mSmin; blob[(long long)(OFF_MC + 2) * K] = mStmp;
  blob[(long long)(OFF_MC + 3) * K] = mp; blob[(long long)(OFF_MC + 4) * K] = mlam;
}

template <int M>
static int launch_fast_m(const McsppArgs &a, cudaStream_t st) {
  constexpr int NT = 64;
  constexpr int NP = M * (M + 1) / 2;
  constexpr int MINB = 4;            // 4 x 64 threads at <= 255 registers: fewer spills beat more warps here (measured)
  const size_t smem = (size_t)3 * NP * NT * sizeof(double);
  auto kern = mcspp_fast_kernel<M, NT, MINB>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)a.S * (a.K - a.k_first);
  kern<<<(unsigned)((items + NT - 1) / NT), NT, smem, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int launch_mcspp_fast(int M, const McsppArgs &a, cudaStream_t st) {
  switch (M) {
    case 2: return launch_fast_m<2>(a, st);
    case 3: return launch_fast_m<3>(a, st);
    case 4: return launch_fast_m<4>(a, st);
    case 5: return launch_fast_m<5>(a, st);
    case 6: return launch_fast_m<6>(a, st);
    case 7: return launch_fast_m<7>(a, st);
    case 8: return launch_fast_m<8>(a, st);
  }
  set_error("mcspp: n_mics %d outside the compiled range 2..8", M);
  return DS_EUNSUPPORTED;
}

}  // namespace ds
