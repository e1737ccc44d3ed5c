// mcspp_fast.cu -- hot-path variant of the McSppBase + MVDR + OMLSA per-bin
// kernel (config 4): output-only state, complex64 spectrum in, no taps.
//
// Same arithmetic as mcspp_kernel (mcspp.cu) -- see the reference citations
// there -- restructured for the fp64 pipe:
//   * every quadratic form is evaluated against the packed upper triangle, xi and gamma through
//     A (Phi_vv + eps I) = I (chain_step.cuh), so no M-vector temporaries besides u = A y survive the
//     inverse and the updated Phi_yy takes over A's registers;
//   * the frame loop is one basic block (branch-free MCRA / exp / sqrt, no integer division);
//   * 4 CTAs x 64 threads per SM at 255 registers next to 4 x 55 KB of shared-memory state, no
//     local-memory traffic inside the frame loop; the CTAs are de-phased once at start-up.
#include "chain_step.cuh"

namespace ds {

#ifdef DS_FAST_TRACE
// timing experiment only: start clock of frames 100..163 of every warp, plus the SM and hardware warp slot it ran on
__device__ long long g_trace[8192 * 66];
#endif

#ifndef DS_FAST_SKEW
#define DS_FAST_SKEW 8000       // cycles, about one frame of the M = 8 kernel; 0 disables
#endif
#ifndef DS_CHAIN_V2
#define DS_CHAIN_V2 0
#endif
#ifndef DS_FAST_MINB
#define DS_FAST_MINB 4
#endif
#ifndef DS_FAST_USE_C
#define DS_FAST_USE_C 1
#endif

#ifdef DS_FAST_MAXNREG
#define DS_FAST_BOUNDS __maxnreg__(DS_FAST_MAXNREG)
#else
#define DS_FAST_BOUNDS __launch_bounds__(NT, MINB)
#endif

// TAP_P: also write the posterior p per frame (the mask of the mask-based beamformers); a separate instantiation,
// because even a predicated-off store costs the headline kernel 2 % (registers: 16.84 -> 17.17 ms measured)
// PAIRED: 256-thread CTAs (one per SM); warps w and w + 4 share a scheduler and meet once per frame at a named barrier, one
// at its frame start, the other after the MVDR denominator -- which pins their phase offset at a favourable value (the
// frame time of a warp depends on where its scheduler partner is: profiles/ab_runs_r02.txt) instead of letting it wander:
// 16.87 -> 16.49 ms.  No thread may leave early then: threads past the end redo the last item.
template <int M, int NT, int MINB, bool TAP_P, bool MIXED = false, bool PAIRED = false>
__global__ void DS_FAST_BOUNDS mcspp_fast_kernel(McsppArgs a) {
  constexpr bool USE_C = DS_FAST_USE_C != 0;
  constexpr int NP = M * (M + 1) / 2;
  constexpr int NE = mcspp_state_elems<M>();
  constexpr int OFF_YR = 0, OFF_VR = NP, OFF_MC = 2 * NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  const int K = a.K;
  const int Kp = K - a.k_first;
  long long g = (long long)blockIdx.x * NT + tid;
  if (g >= (long long)a.S * Kp) {
    // PAIRED: nobody may leave before the last rendezvous -- threads past the end redo the last item (same inputs, same
    // results, the same values stored twice); otherwise threads are independent and simply leave
    if constexpr (PAIRED) g = (long long)a.S * Kp - 1; else return;
  }
  const int s = (int)(g / Kp), k = a.k_first + (int)(g % Kp);
  double *blob = a.state + (long long)s * NE * K + k;
  double *smy = sm + tid;                 // Phi_yy (real part), element e at smy[e * NT]
  double *smv = sm + NP * NT + tid;       // Phi_vv (real part)
  double *smc = sm + 2 * NP * NT + tid;   // (2 - delta_ij) Re(conj(a_i) a_j), constant per bin
#pragma unroll
  for (int e = 0; e < NP; ++e) { smy[e * NT] = blob[(long long)(OFF_YR + e) * K]; smv[e * NT] = blob[(long long)(OFF_VR + e) * K]; }
  double mS = blob[(long long)(OFF_MC + 0) * K], mSmin = blob[(long long)(OFF_MC + 1) * K], mStmp = blob[(long long)(OFF_MC + 2) * K],
         mp = blob[(long long)(OFF_MC + 3) * K], mlam = blob[(long long)(OFF_MC + 4) * K];

  if constexpr (USE_C) {
    double ar[M], ai[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { const double2 v = a.a0[(long long)m * K + k]; ar[m] = v.x; ai[m] = v.y; }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j)
        smc[pidx<M>(i, j) * NT] = ((i == j) ? 1.0 : 2.0) * fma(ai[i], ai[j], ar[i] * ar[j]);
  }
  // ell % L without a division per frame: r follows ell modulo L (mcra.py:52-56 resets ell inside the frame loop)
  int frm = a.frm_cnt, ell = a.ell, ell_mod = a.ell % a.mc.L;
  const float2 *Xp = reinterpret_cast<const float2 *>(a.X) + (long long)s * a.T * M * K + k;
  float2 *Yp = a.Yout + (long long)s * a.T * K + k;
  const double *a0 = reinterpret_cast<const double *>(a.a0 + k);     // (re, im) pairs, mic stride 2K doubles

  McraRegs mc = {mS, mSmin, mStmp, mp, mlam};
#if DS_FAST_SKEW > 0
  if constexpr (!PAIRED) {
    // De-phase the CTAs once at start-up: every warp runs the same ~1800-instruction frame body, and warps that
    // start together stay in lock-step for a long time (same phase => they want the fp64 pipe, the LSU and the
    // MUFU at the same moments).  A pseudo-random delay of 0..7/8 of a frame time per warp costs < 4 us per CTA
    // and is worth 1.9 % of the step (A/B on the B200: 23.12 -> 22.70 ms; 2 600 / 3 900 / 5 200-cycle offsets by
    // scheduler slot, by launch round or random all land within 0.03 ms of each other).  Results are unaffected.
    const unsigned lvl = ((blockIdx.x * 2u + (tid >> 5)) * 2654435761u) >> 29;
    const long long wait = (long long)lvl * (DS_FAST_SKEW / 8), t0 = clock64();
    while (clock64() - t0 < wait) { }
  }
#endif
  const int pair_bar = 1 + ((tid >> 5) & 3);     // PAIRED: warps w and w + 4 of the CTA
  const bool role_b = (tid >> 5) >= 4;
#pragma unroll 1
  for (int t = 0; t < a.T; ++t) {
    if constexpr (PAIRED) pair_sync(pair_bar, !role_b);
    // issue the loads of this frame's spectrum first: they are consumed only after the matrix
    // inverse, so the HBM latency hides behind the Gauss-Jordan sweeps
#ifdef DS_FAST_TRACE
    if ((tid & 31) == 0 && t >= 100 && t < 164) {
      const unsigned gw = blockIdx.x * (NT / 32) + (tid >> 5);
      if (gw < 8192) {
        g_trace[gw * 66 + (t - 100)] = clock64();
        if (t == 100) { unsigned smid, wid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); asm("mov.u32 %0, %%warpid;" : "=r"(wid)); g_trace[gw * 66 + 64] = smid; g_trace[gw * 66 + 65] = wid; }
      }
    }
#endif
    float2 yf[M];
#pragma unroll
    for (int m = 0; m < M; ++m) yf[m] = ld_f2_once(Xp + m * K);
    const float2 ynb0 = (k > 0) ? ld_f2_once(Xp - 1) : make_float2(0.f, 0.f);
    const float2 ynb1 = (k < K - 1) ? ld_f2_once(Xp + 1) : make_float2(0.f, 0.f);
    Xp += M * K;
#ifndef DS_X_PREFETCH
#define DS_X_PREFETCH 4
#endif
    // pull the spectrum of a later frame into L2 now: the loads above then hit L2 instead of
    // HBM (the profile showed the float->double conversions of yf waiting on the long scoreboard)
    if (DS_X_PREFETCH > 0 && t + DS_X_PREFETCH < a.T) {
#pragma unroll
      for (int m = 0; m < M; ++m)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Xp + (long long)(DS_X_PREFETCH - 1) * M * K + m * K));
    }
    const bool reset = (frm > 0) && (ell_mod == 0);
    double p_post;
#if DS_CHAIN_V2
    const float2 yo = chain_bin_step_v2<M, NT, USE_C>(yf, ynb0, ynb1, k, K, frm, reset, mc, smy, smv, smc, a0, a, p_post);
#else
    const float2 yo = chain_bin_step<M, NT, USE_C, MIXED, PAIRED>(yf, ynb0, ynb1, k, K, frm, reset, mc, smy, smv, smc, a0, a, p_post,
                                                                  (PAIRED && role_b) ? pair_bar : 0);
#endif
    *Yp = yo;
    if constexpr (TAP_P) a.tp[((long long)s * a.T + t) * K + k] = p_post;     // the only tap this kernel serves
    if (a.k_first == 2 && k == 2) { Yp[-1] = make_float2(0.f, 0.f); Yp[-2] = make_float2(0.f, 0.f); }
    if (reset) ell = 0;
    ++ell; ++frm;
    ell_mod = (ell_mod + 1 == a.mc.L) ? 0 : ell_mod + 1;
    Yp += K;
  }
  mS = mc.S; mSmin = mc.Smin; mStmp = mc.Stmp; mp = mc.p; mlam = mc.lam;

#pragma unroll
  for (int e = 0; e < NP; ++e) { blob[(long long)(OFF_YR + e) * K] = smy[e * NT]; blob[(long long)(OFF_VR + e) * K] = smv[e * NT]; }
  blob[(long long)(OFF_MC + 0) * K] = mS; blob[(long long)(OFF_MC + 1) * K] = mSmin; blob[(long long)(OFF_MC + 2) * K] = mStmp;
  blob[(long long)(OFF_MC + 3) * K] = mp; blob[(long long)(OFF_MC + 4) * K] = mlam;
}

template <int M>
static int launch_fast_m(const McsppArgs &a, cudaStream_t st) {
#if !defined(DS_FAST_NT) && !defined(DS_FAST_NOPAIR) && !defined(DS_CHAIN_MIXED) && !DS_CHAIN_V2
  // 8 microphones (the headline): one 256-thread CTA per SM with paired warps -- see mcspp_fast_kernel
  constexpr bool PAIRED = (M == 8) && (DS_FAST_USE_C != 0);
#else
  constexpr bool PAIRED = false;
#endif
#ifndef DS_FAST_NT
#define DS_FAST_NT 64
#endif
  constexpr int NT = PAIRED ? 256 : DS_FAST_NT;
  constexpr int NP = M * (M + 1) / 2;
  constexpr int MINB = PAIRED ? 1 : DS_FAST_MINB; // 4 x 64 threads at <= 255 registers: fewer spills beat more warps here (measured)
#ifndef DS_FAST_SMEM_PAD_KB
#define DS_FAST_SMEM_PAD_KB 0     // occupancy experiments only: extra dynamic shared memory per CTA
#endif
  const size_t smem = (size_t)(DS_FAST_USE_C ? 3 : 2) * NP * NT * sizeof(double) + (size_t)DS_FAST_SMEM_PAD_KB * 1024;
#ifdef DS_CHAIN_MIXED      // A/B build only (tools/build_variant.sh x mcspp_fast.cu -DDS_CHAIN_MIXED): see chain_step.cuh
  auto kern = a.tp ? mcspp_fast_kernel<M, NT, MINB, true> : mcspp_fast_kernel<M, NT, MINB, false, true>;
#else
  auto kern = a.tp ? mcspp_fast_kernel<M, NT, MINB, true, false, PAIRED> : mcspp_fast_kernel<M, NT, MINB, false, false, PAIRED>;
#endif
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

#ifdef DS_FAST_TRACE
extern "C" int ds_debug_trace_read(long long *host, size_t bytes) { return (int)cudaMemcpyFromSymbol(host, ds::g_trace, bytes); }
#endif
