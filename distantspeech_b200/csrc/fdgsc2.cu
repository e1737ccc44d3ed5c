// fdgsc2.cu -- FDGSC.process(postfilter=False) (beamformer/FDGSC.py:201-317) as a pipeline of kernels cut along the
// data dependences of the algorithm (same arithmetic as the one-CTA-per-stream kernel in fdgsc.cu, which stays as the
// workspace-free entry point ds_fdgsc_run):
//
//   feed-forward, parallel over (stream, block) -- nothing here depends on an adaptive filter:
//     fd_prologue     delay-line state -> head of the extended buffers
//     fd_fir          time-alignment FIR (fixedbeamformer.py:13-48) + mean beamformer (FDGSC.py:123-138)
//     fd_spec         X_f = rfft([fbf_prev | fbf]) (FastFreqLms.py:157) and |rfft(w [x0_prev | x0])|^2 (FDGSC.py:239-244)
//     fd_recur        per (stream, bin), blocks sequential: BM input power (FastFreqLms.py:158,189), MCRA (L = 60)
//     fd_control      per (stream, block): p[:32] raised when mean(p[32:128]) > .8 (FDGSC.py:247-251), AIC step size
//   recurrent:
//     fd_bm           one WARP per (stream, microphone): the M blocking filters of a stream never talk to each other
//                     (gsc_bm.py:61-122) -- 4 FFTs per block, no CTA barrier, weights in shared memory for the whole call
//     fd_aic          one CTA per stream, one warp per channel: the M-channel canceller (gsc_aic.py:54-108); the channels
//                     couple through sum_ch X W and the norm constraint: block barriers, warp-shuffle reductions
//     fd_epilogue     tails of the buffers -> delay-line state
//
// The monolith spends its time in CTA-wide barriers around ~46 FFTs per block with most warps idle (ncu: 2.4 barrier
// stalls per issue, warps active 21 %); here the 24 blocking-filter FFTs run barrier-free at full occupancy and the
// canceller keeps only the barriers its data flow needs.  Intermediate signals (aligned channels, blocking outputs,
// X_f, powers, p) travel through a caller-provided HBM workspace: about 3x the input, read and written once.
#include "common.cuh"
#include "fft.cuh"
#include "perbin.cuh"

namespace ds {

constexpr int F2_L = 256, F2_N = 512, F2_K = 257, F2_H = 256, F2_FLMAX = 128;
// resident CTAs per SM the register allocation aims for.  A/B on the B200 (4096 streams x 10 s x 6 mics, CUDA events):
// no bound (79 / 117 registers, 24 / 12 warps per SM) 272 ms, (2, 2) 272 ms, (4, 4) 210 ms, (4, 3) 206 ms, (3, 3) 204.6 ms
#ifndef FD_BM_MINB
#define FD_BM_MINB 3
#endif
#ifndef FD_BM_PF
#define FD_BM_PF "prefetch.global.L2"
#endif
#ifndef FD_AIC_MINB
#define FD_AIC_MINB 4
#endif

struct Fd2Args {
  double *state;              // [S][elems]
  const double *h;            // [M][FL]
  const float *x;             // [S][M][Ns] DC-notched input
  float *y;                   // [S][Ns]
  float *bm_out;              // [S][M][Ns] or null
  float *fix_out;             // [S][Ns] or null
  double *p_out;              // [S][nblk][K] or null
  const double *window;       // [512]
  unsigned char *ws;          // workspace
  int S, M, Ns, FL, frm_cnt, ell, nblk;
  double mu_bm, mu_aic, alpha, maxnorm, delta;
  McraConst mc;
};

// state offsets (doubles), identical to fdgsc.cu's load / save order
struct Fd2State {
  size_t Wbm, Waic, Pf, Pa, fbf_prev, bm_prev, x0_prev, cache, dl_al, dl_fbf, mcra, notch, total;
  __host__ __device__ explicit Fd2State(int M) {
    size_t o = 0;
    Wbm = o; o += (size_t)2 * M * F2_K;
    Waic = o; o += (size_t)2 * M * F2_K;
    Pf = o; o += F2_K;
    Pa = o; o += F2_K;
    fbf_prev = o; o += F2_L;
    bm_prev = o; o += (size_t)M * F2_L;
    x0_prev = o; o += F2_L;
    cache = o; o += (size_t)M * (F2_FLMAX - 1);
    dl_al = o; o += (size_t)M * (F2_L / 2);
    dl_fbf = o; o += F2_L;
    mcra = o; o += 5 * F2_K;
    notch = o; o += 2 * (size_t)M;
    total = o;
  }
};

// workspace layout for element type T
template <typename T> struct Fd2Ws {
  typedef typename V2<T>::type C2;
  size_t A, F, B, Xf, Pf, P0, step, total;
  __host__ __device__ Fd2Ws(int S, int M, int Ns) {
    const size_t nblk = Ns / F2_L;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    A = o; o += al((size_t)S * M * (F2_L / 2 + Ns) * sizeof(T));      // aligned channels behind their 128-sample delay line
    F = o; o += al((size_t)S * (F2_L + Ns) * sizeof(T));              // fixed-beamformer output behind its 256-sample delay line
    B = o; o += al((size_t)S * M * (F2_L + Ns) * sizeof(T));          // blocking-matrix outputs behind the previous block
    Xf = o; o += al((size_t)S * nblk * F2_K * sizeof(C2));
    Pf = o; o += al((size_t)S * nblk * F2_K * sizeof(T));
    P0 = o; o += al((size_t)2 * S * nblk * F2_K * sizeof(double));    // two planes: |X_0|^2 of microphone 0, MCRA p
    step = o; o += al((size_t)S * nblk * sizeof(T));
    total = o;
  }
};

// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void fd_prologue_kernel(Fd2Args a) {
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  T *A = reinterpret_cast<T *>(a.ws + w.A), *F = reinterpret_cast<T *>(a.ws + w.F), *B = reinterpret_cast<T *>(a.ws + w.B);
  const int s = blockIdx.x;
  const double *st = a.state + (size_t)s * so.total;
  for (int i = threadIdx.x; i < a.M * (F2_L / 2); i += blockDim.x) {
    const int m = i / (F2_L / 2), n = i % (F2_L / 2);
    A[((size_t)s * a.M + m) * (F2_L / 2 + a.Ns) + n] = (T)st[so.dl_al + i];
  }
  for (int i = threadIdx.x; i < F2_L; i += blockDim.x) F[(size_t)s * (F2_L + a.Ns) + i] = (T)st[so.dl_fbf + i];
  for (int i = threadIdx.x; i < a.M * F2_L; i += blockDim.x) {
    const int m = i / F2_L, n = i % F2_L;
    B[((size_t)s * a.M + m) * (F2_L + a.Ns) + n] = (T)st[so.bm_prev + i];
  }
}

// time-alignment FIR + mean beamformer: CTA = (tile of 1024 samples, stream), thread = 4 consecutive samples of every
// microphone.  Taps run in groups of four over a register window of the input: per group one 16-byte shared-memory load
// of samples (conflict-free) and one broadcast load of taps feed 16 multiply-adds; every output still sums its taps in
// ascending order like the sequential FIR.
#ifndef FIR_NPT
#define FIR_NPT 4          // consecutive outputs per thread (A/B, FDGSC pipeline ms: 4 -> 129.1, 8 -> 135.8 (the 32-byte lane stride
                           // makes the 16-byte sample loads two-way bank conflicted), 16 -> 157.0)
#endif
constexpr int FIR_TS = 1024, FIR_NT = FIR_TS / FIR_NPT;
template <typename T> __device__ __forceinline__ void load4(const T *p, T (&v)[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 q = *reinterpret_cast<const float4 *>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    const double2 q0 = *reinterpret_cast<const double2 *>(p), q1 = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
  }
}
template <typename T>
__global__ void __launch_bounds__(FIR_NT) fd_fir_kernel(Fd2Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const int M = a.M, FL = a.FL, s = blockIdx.y, tile0 = blockIdx.x * FIR_TS;
  const int nt = min(FIR_TS, a.Ns - tile0);
  constexpr int XS = FIR_TS + F2_FLMAX;            // row stride of the staged input: xs[m][F2_FLMAX + i] = x[tile0 + i]
  T *xs = reinterpret_cast<T *>(smem_raw);         // [M][XS], history of F2_FLMAX - 1 samples at [1, F2_FLMAX)
  T *hs = xs + (size_t)M * XS;                     // [M][F2_FLMAX], zero beyond FL
  const double *st = a.state + (size_t)s * so.total;
  for (int i = threadIdx.x; i < M * F2_FLMAX; i += FIR_NT) {
    const int m = i / F2_FLMAX, k = i % F2_FLMAX;
    hs[i] = (k < FL) ? (T)a.h[m * FL + k] : (T)0;
  }
  for (int m = 0; m < M; ++m) {
    const float *xrow = a.x + ((size_t)s * M + m) * a.Ns;
    for (int i = threadIdx.x; i < FIR_TS + F2_FLMAX; i += FIR_NT) {
      const int g = tile0 - F2_FLMAX + i;          // sample index; negative: carried FIR cache (last FLMAX-1 samples)
      T v = (T)0;
      if (g >= 0) { if (g < a.Ns) v = (T)xrow[g]; }
      else if (g >= -(F2_FLMAX - 1)) v = (T)st[so.cache + (size_t)m * (F2_FLMAX - 1) + (F2_FLMAX - 1) + g];
      xs[(size_t)m * XS + i] = v;
    }
  }
  __syncthreads();
  const int n0 = threadIdx.x * FIR_NPT;
  if (n0 >= nt) return;                            // nt is a multiple of 256 (Ns is): whole groups of FIR_NPT
  T mean[FIR_NPT];
#pragma unroll
  for (int j = 0; j < FIR_NPT; ++j) mean[j] = (T)0;
  T *A = reinterpret_cast<T *>(a.ws + w.A);
  const int nq = (FL + 3) / 4;
  for (int m = 0; m < M; ++m) {
    const T *row = xs + (size_t)m * XS + F2_FLMAX + n0;
    const T *hm = hs + m * F2_FLMAX;
    T acc[FIR_NPT];
#pragma unroll
    for (int j = 0; j < FIR_NPT; ++j) acc[j] = (T)0;
    // sliding register window X[i] = x[n0 - 4q - 4 + i], i = 0 .. FIR_NPT + 3: per group of four taps one 16-byte load of
    // samples and one broadcast load of taps feed 4 FIR_NPT multiply-adds
    T X[FIR_NPT + 4], h4[4];
#pragma unroll
    for (int g = 0; g < FIR_NPT / 4; ++g) { T v[4]; load4<T>(row + 4 * g, v); X[4 + 4 * g] = v[0]; X[5 + 4 * g] = v[1]; X[6 + 4 * g] = v[2]; X[7 + 4 * g] = v[3]; }
#pragma unroll 3
    for (int q = 0; q < nq; ++q) {
      { T v[4]; load4<T>(row - 4 * q - 4, v); X[0] = v[0]; X[1] = v[1]; X[2] = v[2]; X[3] = v[3]; }   // x[n0-4q-4 .. n0-4q-1]
      load4<T>(hm + 4 * q, h4);
      // y[n0+j] += h[k] x[n0 + j - k],  k = 4q + t  ->  X[4 + j - t]; taps in ascending order for every output
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int j = 0; j < FIR_NPT; ++j) acc[j] += h4[t] * X[4 + j - t];
#pragma unroll
      for (int i = FIR_NPT + 3; i >= 4; --i) X[i] = X[i - 4];
    }
    T *dst = A + ((size_t)s * M + m) * (F2_L / 2 + a.Ns) + F2_L / 2 + tile0 + n0;
#pragma unroll
    for (int j = 0; j < FIR_NPT; ++j) { dst[j] = acc[j]; mean[j] += acc[j]; }
  }
  T *F = reinterpret_cast<T *>(a.ws + w.F) + (size_t)s * (F2_L + a.Ns) + F2_L + tile0 + n0;
#pragma unroll
  for (int j = 0; j < FIR_NPT; ++j) {
    const T v = mean[j] / (T)M;                    // np.mean(x, axis=1)  (FDGSC.py:138)
    F[j] = v;
    if (a.fix_out) a.fix_out[(size_t)s * a.Ns + tile0 + n0 + j] = (float)v;
  }
}

// spectra of the blocking-matrix reference and of raw microphone 0: one warp per (stream, block)
constexpr int SPEC_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(SPEC_WARPS * 32) fd_spec_kernel(Fd2Args a, const typename V2<T>::type *__restrict__ tw_h_g,
                                                                  const typename V2<T>::type *__restrict__ tw_n_g) {
  typedef typename V2<T>::type C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = F2_N, H = F2_H, K = F2_K, L = F2_L, BE = fft_buf_elems(F2_N);
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  C2 *tw_h = reinterpret_cast<C2 *>(smem_raw);
  C2 *tw_n = tw_h + H;
  T *win = reinterpret_cast<T *>(tw_n + (H / 2 + 2));
  C2 *buf = reinterpret_cast<C2 *>(win + N) + (size_t)warp * BE;
  T *fb = reinterpret_cast<T *>(buf);
  for (int i = threadIdx.x; i < H; i += blockDim.x) tw_h[i] = tw_h_g[i];
  for (int i = threadIdx.x; i <= H / 2; i += blockDim.x) tw_n[i] = tw_n_g[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) win[i] = (T)a.window[i];
  __syncthreads();
  const long long item = (long long)blockIdx.x * SPEC_WARPS + warp;
  if (item >= (long long)a.S * a.nblk) return;
  const int s = (int)(item / a.nblk), b = (int)(item % a.nblk);
#define FIDX(n) (2 * FPAD<T>((n) >> 1) + ((n) & 1))
  typedef FftFirst<H, T> F1;                           // first radix-8 pass fed from registers (coalesced 8-byte global loads)
  // X_f = rfft([fbf_prev | fbf]): 512 contiguous samples of the extended buffer
  const T *F = reinterpret_cast<const T *>(a.ws + w.F) + (size_t)s * (L + a.Ns) + (size_t)b * L;
  {
    C2 v[F1::PER][F1::R];
    const C2 *src = reinterpret_cast<const C2 *>(F);
#pragma unroll
    for (int i = 0; i < F1::PER; ++i)
#pragma unroll
      for (int r = 0; r < F1::R; ++r) v[i][r] = src[lane + 32 * i + r * F1::NB];
    F1::run(v, buf, tw_h, lane);
  }
  C2 *Xf = reinterpret_cast<C2 *>(a.ws + w.Xf) + ((size_t)s * a.nblk + b) * K;
#pragma unroll
  for (int i = 0; i < 5; ++i) {                        // real-FFT split straight to global memory
    const int k = lane + 32 * i;
    if (k <= H / 2) {
      if (k == 0) {
        const C2 z = buf[FPAD<T>(0)];
        Xf[0] = mk2<T>(z.x + z.y, (T)0); Xf[H] = mk2<T>(z.x - z.y, (T)0);
      } else {
        C2 x1, x2;
        rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], x1, x2);
        Xf[k] = x1; Xf[H - k] = x2;
      }
    }
  }
  __syncwarp();
  // |rfft(window * [x0_prev | x0])|^2 with the reference's complex64 rounding (transform.py:212)
  const float *x0 = a.x + (size_t)s * a.M * a.Ns;
  const double *st = a.state + (size_t)s * so.total;
  if (b >= 1) {
    C2 v[F1::PER][F1::R];
    const float2 *src = reinterpret_cast<const float2 *>(x0 + (size_t)(b - 1) * L);
    const C2 *w2 = reinterpret_cast<const C2 *>(win);
#pragma unroll
    for (int i = 0; i < F1::PER; ++i)
#pragma unroll
      for (int r = 0; r < F1::R; ++r) {
        const int e = lane + 32 * i + r * F1::NB;
        const float2 xv = src[e];
        const C2 wv = w2[e];
        v[i][r] = mk2<T>(mul_rn((T)xv.x, wv.x), mul_rn((T)xv.y, wv.y));      // rounded on its own, like the staged path
      }
    F1::run(v, buf, tw_h, lane);
  } else {
    for (int n = lane; n < N; n += 32) {
      const int g = (b - 1) * L + n;
      const T xv = (g >= 0) ? (T)x0[g] : (T)st[so.x0_prev + n];
      fb[FIDX(n)] = mul_rn(xv, win[n]);
    }
    __syncwarp();
    warp_cfft<H, T>(buf, tw_h, lane);
  }
  double *P0 = reinterpret_cast<double *>(a.ws + w.P0) + ((size_t)s * a.nblk + b) * K;
  auto pw = [](C2 v) { const double re = (double)(float)v.x, im = (double)(float)v.y; return re * re + im * im; };
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int k = lane + 32 * i;
    if (k <= H / 2) {
      if (k == 0) {
        const C2 z = buf[FPAD<T>(0)];
        P0[0] = pw(mk2<T>(z.x + z.y, (T)0)); P0[H] = pw(mk2<T>(z.x - z.y, (T)0));
      } else {
        C2 x1, x2;
        rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], x1, x2);
        P0[k] = pw(x1); P0[H - k] = pw(x2);
      }
    }
  }
#undef FIDX
}

// per (stream, bin), blocks sequential: BM input power with its floor, MCRA on microphone 0
template <typename T>
__global__ void __launch_bounds__(128) fd_recur_kernel(Fd2Args a) {
  typedef typename V2<T>::type C2;
  constexpr int K = F2_K;
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *st = a.state + (size_t)s * so.total;
  T pf = (T)st[so.Pf + k];
  double m0 = st[so.mcra + 0 * K + k], m1 = st[so.mcra + 1 * K + k], m2 = st[so.mcra + 2 * K + k], m3 = st[so.mcra + 3 * K + k],
         m4 = st[so.mcra + 4 * K + k];
  const C2 *Xf = reinterpret_cast<const C2 *>(a.ws + w.Xf) + (size_t)s * a.nblk * K + k;
  T *Pf = reinterpret_cast<T *>(a.ws + w.Pf) + (size_t)s * a.nblk * K + k;
  double *P0 = reinterpret_cast<double *>(a.ws + w.P0) + (size_t)s * a.nblk * K + k;
  int frm = a.frm_cnt, ell = a.ell % a.mc.L;
  double *Pp = P0 + (size_t)a.S * a.nblk * K;      // p goes to the second plane (neighbour bins still read the powers)
  for (int b = 0; b < a.nblk; ++b) {
    const C2 v = Xf[(size_t)b * K];
    T pn = (T)a.alpha * pf + ((T)1 - (T)a.alpha) * (v.x * v.x + v.y * v.y);        // FastFreqLms.py:158
    pf = (pn < (T)1e-4) ? (T)1e-4 : pn;                                            // :189
    Pf[(size_t)b * K] = pf;
    const double Y0 = P0[(size_t)b * K];
    const double Ym1 = (k > 0) ? P0[(size_t)b * K - 1] : 0.0, Yp1 = (k < K - 1) ? P0[(size_t)b * K + 1] : 0.0;
    const bool reset = (frm > 0) && (ell == 0);
    mcra_step(m0, m1, m2, m3, m4, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
    if (reset) ell = 0;
    ++ell; ++frm;
    if (ell == a.mc.L) ell = 0;
    Pp[(size_t)b * K] = m3;
  }
  st[so.Pf + k] = (double)pf;
  st[so.mcra + 0 * K + k] = m0; st[so.mcra + 1 * K + k] = m1; st[so.mcra + 2 * K + k] = m2; st[so.mcra + 3 * K + k] = m3;
  st[so.mcra + 4 * K + k] = m4;
}

// per (stream, block): adaptation-control heuristics across bins, AIC step size; one warp each
template <typename T>
__global__ void __launch_bounds__(256) fd_control_kernel(Fd2Args a) {
  constexpr int K = F2_K;
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 8 + warp;
  if (item >= (long long)a.S * a.nblk) return;
  double *Pp = reinterpret_cast<double *>(a.ws + w.P0) + (size_t)a.S * a.nblk * K + (size_t)item * K;
  double p[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { const int k = lane + 32 * i; p[i] = (k < K) ? Pp[k] : 0.0; }
  // mean(p[32:128]) > 0.8  ->  p[:32] = max(p[:32], 0.8)            (FDGSC.py:247-249)
  double mid = p[1] + p[2] + p[3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mid += __shfl_xor_sync(0xffffffffu, mid, o);
  mid /= 96.0;
  if (mid > 0.8 && p[0] < 0.8) p[0] = 0.8;
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) tot += p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  const double pbar = tot / (double)K;
  if (lane == 0) reinterpret_cast<T *>(a.ws + w.step)[item] = (T)((1.0 - pbar) * a.mu_aic);       // p * mu (gsc_aic.py:82, FDGSC.py:279)
  if (a.p_out) {
    double *po = a.p_out + (size_t)item * K;
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int k = lane + 32 * i; if (k < K) po[k] = p[i]; }
  }
}

// blocking matrix: one warp per (stream, microphone), blocks sequential, no CTA-wide synchronisation
constexpr int BM_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(BM_WARPS * 32, FD_BM_MINB) fd_bm_kernel(Fd2Args a, const typename V2<T>::type *__restrict__ tw_h_g,
                                                              const typename V2<T>::type *__restrict__ tw_n_g) {
  typedef typename V2<T>::type C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = F2_N, H = F2_H, K = F2_K, L = F2_L, BE = fft_buf_elems(F2_N);
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, M = a.M;
  C2 *tw_h = reinterpret_cast<C2 *>(smem_raw);
  C2 *tw_n = tw_h + H;
  C2 *buf = tw_n + (H / 2 + 2) + (size_t)warp * BE;
  C2 *W = tw_n + (H / 2 + 2) + (size_t)BM_WARPS * BE + (size_t)warp * (K + 1);
  T *fb = reinterpret_cast<T *>(buf);
  for (int i = threadIdx.x; i < H; i += blockDim.x) tw_h[i] = tw_h_g[i];
  for (int i = threadIdx.x; i <= H / 2; i += blockDim.x) tw_n[i] = tw_n_g[i];
  __syncthreads();
  const long long item = (long long)blockIdx.x * BM_WARPS + warp;
  if (item >= (long long)a.S * M) return;
  const int s = (int)(item / M), m = (int)(item % M);
  double *st = a.state + (size_t)s * so.total + so.Wbm + (size_t)2 * m * K;
  for (int k = lane; k < K; k += 32) W[k] = mk2<T>((T)st[2 * k], (T)st[2 * k + 1]);
  __syncwarp();
  const C2 *Xf = reinterpret_cast<const C2 *>(a.ws + w.Xf) + (size_t)s * a.nblk * K;
  const T *Pf = reinterpret_cast<const T *>(a.ws + w.Pf) + (size_t)s * a.nblk * K;
  const T *xad = reinterpret_cast<const T *>(a.ws + w.A) + ((size_t)s * M + m) * (L / 2 + a.Ns);   // delayed by L/2 through the buffer head
  T *bm = reinterpret_cast<T *>(a.ws + w.B) + ((size_t)s * M + m) * (L + a.Ns) + L;
  float *bmo = a.bm_out ? a.bm_out + ((size_t)s * M + m) * a.Ns : nullptr;
  const T invN = (T)1 / (T)N, step_bm = (T)(1.0 * a.mu_bm);                  // p = 1.0 (gsc_bm.py:90, FDGSC.py:260)
#define FIDX(n) (2 * FPAD<T>((n) >> 1) + ((n) & 1))
  for (int b = 0; b < a.nblk; ++b) {
    if (b + 1 < a.nblk) {
      // next block's operands towards L2 now (ncu: 3.4 long-scoreboard stalls per issue without it): 128-byte lines
      const char *px = reinterpret_cast<const char *>(Xf + (size_t)(b + 1) * K), *pp = reinterpret_cast<const char *>(Pf + (size_t)(b + 1) * K);
      const char *pd = reinterpret_cast<const char *>(xad + (size_t)(b + 1) * L);
      if (lane * 128 < (int)(K * sizeof(C2))) asm volatile(FD_BM_PF " [%0];" ::"l"(px + lane * 128));
      if (lane * 128 < (int)(K * sizeof(T))) asm volatile(FD_BM_PF " [%0];" ::"l"(pp + lane * 128));
      if (lane * 128 < (int)(L * sizeof(T))) asm volatile(FD_BM_PF " [%0];" ::"l"(pd + lane * 128));
    }
    // The real-FFT split / merge passes are fused with the elementwise work around them (same operations and rounding as
    // warp_rfft / warp_irfft_unscaled, 15 instead of 22 shared-memory round trips per block): each lane owns the bin
    // pairs (k, H - k), k = lane + 32 i.
    const C2 *Xb = Xf + (size_t)b * K;
    const T *Pb = Pf + (size_t)b * K;
    // (1) y = irfft(X_f W): products of a pair, merged, straight into the transform buffer
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = lane + 32 * i;
      if (k <= H / 2) {
        if (k == 0) {
          const T p0 = cmul(Xb[0], W[0]).x, ph = cmul(Xb[H], W[H]).x;
          buf[FPAD<T>(0)] = mk2<T>(p0 + ph, -(p0 - ph));
        } else {
          C2 z1, z2;
          irfft_merge_pair<T, C2>(cmul(Xb[k], W[k]), cmul(Xb[H - k], W[H - k]), tw_n[k], z1, z2);
          buf[FPAD<T>(k)] = z1; buf[FPAD<T>(H - k)] = z2;
        }
      }
    }
    __syncwarp();
    warp_cfft<H, T>(buf, tw_h, lane);                                       // x[2n] = r.x, x[2n+1] = -r.y (sign folded below)
    {
      // e = d - y on the last hop_len samples (:161, :174), computed by the lane that owns those samples in the next
      // transform's first radix-8 pass (complex element c = samples 2c, 2c + 1; the upper half of the frame), so the
      // zero-padded error [0 | e] (:185) goes into that transform from registers
      typedef FftFirst<H, T> F1;
      C2 v[F1::PER][F1::R];
      const C2 *xd2 = reinterpret_cast<const C2 *>(xad + (size_t)b * L);
      C2 *bm2 = reinterpret_cast<C2 *>(bm + (size_t)b * L);
#pragma unroll
      for (int i = 0; i < F1::PER; ++i)
#pragma unroll
        for (int r = 0; r < F1::R; ++r) {
          const int c = lane + 32 * i + r * F1::NB;
          C2 z = mk2<T>((T)0, (T)0);
          if (2 * c >= L) {
            const C2 q = buf[FPAD<T>(c)], d = xd2[c - L / 2];
            z = mk2<T>(d.x - q.x * invN, d.y - (-q.y) * invN);
            bm2[c - L / 2] = z;
            if (bmo) *reinterpret_cast<float2 *>(bmo + (size_t)b * L + 2 * c - L) = make_float2((float)z.x, (float)z.y);
          }
          v[i][r] = z;
        }
      __syncwarp();
      F1::run(v, buf, tw_h, lane);
    }
    // (2) E = split, W' = W + mu conj(X_f) E / P_f, merged for the constraint's inverse transform -- all per pair in registers
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = lane + 32 * i;
      if (k <= H / 2) {
        if (k == 0) {
          const C2 z = buf[FPAD<T>(0)];
          const C2 E0 = mk2<T>(z.x + z.y, (T)0), Eh = mk2<T>(z.x - z.y, (T)0);
          const C2 g0 = cmulc(E0, Xb[0]), gh = cmulc(Eh, Xb[H]);
          const T w0 = W[0].x + step_bm * (g0.x * ((T)1 / Pb[0])), wh = W[H].x + step_bm * (gh.x * ((T)1 / Pb[H]));
          buf[FPAD<T>(0)] = mk2<T>(w0 + wh, -(w0 - wh));
        } else {
          C2 E1, E2, z1, z2;
          rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], E1, E2);
          const C2 g1 = cmulc(E1, Xb[k]), g2 = cmulc(E2, Xb[H - k]);        // conj(X) * E
          const T ip1 = (T)1 / Pb[k], ip2 = (T)1 / Pb[H - k];
          C2 w1 = W[k], w2 = W[H - k];
          w1.x += step_bm * (g1.x * ip1); w1.y += step_bm * (g1.y * ip1);
          w2.x += step_bm * (g2.x * ip2); w2.y += step_bm * (g2.y * ip2);
          irfft_merge_pair<T, C2>(w1, w2, tw_n[k], z1, z2);
          buf[FPAD<T>(k)] = z1; buf[FPAD<T>(H - k)] = z2;
        }
      }
    }
    __syncwarp();
    warp_cfft<H, T>(buf, tw_h, lane);
    // (3) constraint in the time domain (sign of the inverse folded in): zero tail, tap bounds -- read in the order of the
    //     next transform's first radix-8 pass, which is then fed from registers
    {
      typedef FftFirst<H, T> F1;
      C2 v[F1::PER][F1::R];
#pragma unroll
      for (int i = 0; i < F1::PER; ++i)
#pragma unroll
        for (int r = 0; r < F1::R; ++r) {
          const int c = lane + 32 * i + r * F1::NB;                         // complex element = samples 2c, 2c + 1
          C2 z = mk2<T>((T)0, (T)0);
          if (2 * c < L) {                                                  // w[-hop_len:] = 0 (:94)
            const C2 q = buf[FPAD<T>(c)];
            T t0 = q.x * invN, t1 = (-q.y) * invN;
            const int d0 = 2 * c - N / 4, d1 = d0 + 1;
            T ub0 = (T)a.delta, ub1 = (T)a.delta;
            if (d0 == 0) ub0 = (T)0.9; else if (d0 == 1 || d0 == -1) ub0 = (T)0.3; else if (d0 == 2 || d0 == -2) ub0 = (T)0.05;
            if (d1 == 0) ub1 = (T)0.9; else if (d1 == 1 || d1 == -1) ub1 = (T)0.3; else if (d1 == 2 || d1 == -2) ub1 = (T)0.05;
            z = mk2<T>(fmin(fmax(t0, -(T)a.delta), ub0), fmin(fmax(t1, -(T)a.delta), ub1));      // tap bounds (:48-59, :96-108)
          }
          v[i][r] = z;
        }
      __syncwarp();                                                         // every lane has read before any lane overwrites
      F1::run(v, buf, tw_h, lane);
    }
    // (4) W = split, straight into the weight array
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = lane + 32 * i;
      if (k <= H / 2) {
        if (k == 0) {
          const C2 z = buf[FPAD<T>(0)];
          W[0] = mk2<T>(z.x + z.y, (T)0); W[H] = mk2<T>(z.x - z.y, (T)0);
        } else {
          C2 x1, x2;
          rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], x1, x2);
          W[k] = x1; W[H - k] = x2;
        }
      }
    }
    __syncwarp();
  }
#undef FIDX
  for (int k = lane; k < K; k += 32) { st[2 * k] = (double)W[k].x; st[2 * k + 1] = (double)W[k].y; }
}

// transform buffer elements the canceller really needs: the XOR swizzle of 8-byte elements stays inside 16-element groups
// (272 for H + 1 = 257), the padded layout of 16-byte elements needs fft_buf_elems -- the difference lets four canceller
// CTAs of six warps share an SM
template <typename T> __host__ __device__ constexpr int aic_buf_elems() { return sizeof(T) == 4 ? ((F2_K + 15) / 16) * 16 : fft_buf_elems(F2_N); }

// interference canceller: CTA per stream, warp per channel.  The reference spectra X_a of block b + 1 depend only on the
// blocking outputs, so they are computed by the otherwise idle warps WHILE warp 0 runs the serial output / error-spectrum
// step of block b (ncu on the unpipelined version: 3.8 barrier stalls per issue): 4 transform times per block on the
// critical path instead of 5, X_a double-buffered in shared memory.
template <typename T>
__global__ void __launch_bounds__(256, FD_AIC_MINB) fd_aic_kernel(Fd2Args a, const typename V2<T>::type *__restrict__ tw_h_g,
                                                     const typename V2<T>::type *__restrict__ tw_n_g) {
  typedef typename V2<T>::type C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = F2_N, H = F2_H, K = F2_K, L = F2_L, BE = aic_buf_elems<T>();
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const int M = a.M, NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, s = blockIdx.x;
  C2 *tw_h = reinterpret_cast<C2 *>(smem_raw);
  C2 *tw_n = tw_h + H;
  C2 *bufs = tw_n + (H / 2 + 2);                       // [M][BE]
  C2 *Waic = bufs + (size_t)M * BE;                    // [M][K]
  C2 *XaBuf = Waic + (size_t)M * K;                    // [2][M][K]
  C2 *Ef = XaBuf + (size_t)2 * M * K;                  // [K]
  T *Pa = reinterpret_cast<T *>(Ef + K + 1);           // [K]
  double *red = reinterpret_cast<double *>((reinterpret_cast<size_t>(Pa + K) + 15) & ~(size_t)15);    // [8]
  double *st = a.state + (size_t)s * so.total;
  for (int i = tid; i < H; i += NT) tw_h[i] = tw_h_g[i];
  for (int i = tid; i <= H / 2; i += NT) tw_n[i] = tw_n_g[i];
  for (int i = tid; i < M * K; i += NT) Waic[i] = mk2<T>((T)st[so.Waic + 2 * i], (T)st[so.Waic + 2 * i + 1]);
  for (int i = tid; i < K; i += NT) Pa[i] = (T)st[so.Pa + i];
  __syncthreads();
  C2 *buf = bufs + (size_t)warp * BE;
  T *fb = reinterpret_cast<T *>(buf);
  const T *Bs = reinterpret_cast<const T *>(a.ws + w.B) + (size_t)s * M * (L + a.Ns);               // [M][L + Ns]: [bm_prev | bm ...]
  const T *Fd = reinterpret_cast<const T *>(a.ws + w.F) + (size_t)s * (L + a.Ns);                   // fbf delayed by one block
  const T *stepv = reinterpret_cast<const T *>(a.ws + w.step) + (size_t)s * a.nblk;
  const T invN = (T)1 / (T)N;
#define FIDX(n) (2 * FPAD<T>((n) >> 1) + ((n) & 1))
  // X_a = rfft([bm_prev | bm]) of channel m for block b into plane `pl`
  auto spectrum = [&](int m, int b, int pl) {
    const C2 *src = reinterpret_cast<const C2 *>(Bs + (size_t)m * (L + a.Ns) + (size_t)b * L);
    {
      typedef FftFirst<H, T> F1;                       // first radix-8 pass fed from registers (coalesced 8-byte global loads)
      C2 v[F1::PER][F1::R];
#pragma unroll
      for (int i = 0; i < F1::PER; ++i)
#pragma unroll
        for (int r = 0; r < F1::R; ++r) v[i][r] = src[lane + 32 * i + r * F1::NB];
      F1::run(v, buf, tw_h, lane);
    }
    C2 *dst = XaBuf + ((size_t)pl * M + m) * K;                             // real-FFT split straight into the plane
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = lane + 32 * i;
      if (k <= H / 2) {
        if (k == 0) {
          const C2 z = buf[FPAD<T>(0)];
          dst[0] = mk2<T>(z.x + z.y, (T)0); dst[H] = mk2<T>(z.x - z.y, (T)0);
        } else {
          C2 x1, x2;
          rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], x1, x2);
          dst[k] = x1; dst[H - k] = x2;
        }
      }
    }
    __syncwarp();
  };
  spectrum(warp, 0, 0);
  __syncthreads();
  for (int b = 0; b < a.nblk; ++b) {
    const int cur = b & 1;
    const C2 *Xa = XaBuf + (size_t)cur * M * K;
    // (b) power, sum_ch X W
    for (int k = tid; k < K; k += NT) {
      T pw = (T)0;
      C2 acc = mk2<T>((T)0, (T)0);
      for (int m = 0; m < M; ++m) {
        const C2 v = Xa[(size_t)m * K + k];
        pw += v.x * v.x + v.y * v.y;
        const C2 pr = cmul(v, Waic[(size_t)m * K + k]);
        acc.x += pr.x; acc.y += pr.y;
      }
      const T pa = (T)a.alpha * Pa[k] + ((T)1 - (T)a.alpha) * pw;
      Pa[k] = (pa < (T)1e-4) ? (T)1e-4 : pa;
      Ef[k] = acc;                                                          // sum_ch X W  (:161)
    }
    __syncthreads();
    if (warp == 0) {
      // (c) output block and error spectrum (merge / split fused with the copies in and out, sign of the inverse folded in)
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int k = lane + 32 * i;
        if (k <= H / 2) {
          if (k == 0) {
            const T p0 = Ef[0].x, ph = Ef[H].x;
            buf[FPAD<T>(0)] = mk2<T>(p0 + ph, -(p0 - ph));
          } else {
            C2 z1, z2;
            irfft_merge_pair<T, C2>(Ef[k], Ef[H - k], tw_n[k], z1, z2);
            buf[FPAD<T>(k)] = z1; buf[FPAD<T>(H - k)] = z2;
          }
        }
      }
      __syncwarp();
      warp_cfft<H, T>(buf, tw_h, lane);
      {
        typedef FftFirst<H, T> F1;                     // e = d - y and [0 | e] straight into the next transform's first pass
        C2 v[F1::PER][F1::R];
        const C2 *fd2 = reinterpret_cast<const C2 *>(Fd + (size_t)b * L);
        float2 *y2 = reinterpret_cast<float2 *>(a.y + (size_t)s * a.Ns + (size_t)b * L);
#pragma unroll
        for (int i = 0; i < F1::PER; ++i)
#pragma unroll
          for (int r = 0; r < F1::R; ++r) {
            const int c = lane + 32 * i + r * F1::NB;
            C2 z = mk2<T>((T)0, (T)0);
            if (2 * c >= L) {
              const C2 q = buf[FPAD<T>(c)], d = fd2[c - L / 2];
              z = mk2<T>(d.x - q.x * invN, d.y - (-q.y) * invN);            // e = d - y
              y2[c - L / 2] = make_float2((float)z.x, (float)z.y);
            }
            v[i][r] = z;
          }
        __syncwarp();
        F1::run(v, buf, tw_h, lane);
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int k = lane + 32 * i;
        if (k <= H / 2) {
          if (k == 0) {
            const C2 z = buf[FPAD<T>(0)];
            Ef[0] = mk2<T>(z.x + z.y, (T)0); Ef[H] = mk2<T>(z.x - z.y, (T)0);
          } else {
            C2 x1, x2;
            rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], x1, x2);
            Ef[k] = x1; Ef[H - k] = x2;
          }
        }
      }
    } else if (b + 1 < a.nblk) {
      // meanwhile: reference spectra of the next block (warp 1 also takes channel 0)
      spectrum(warp, b + 1, cur ^ 1);
      if (warp == 1) spectrum(0, b + 1, cur ^ 1);
    }
    __syncthreads();
    // (d) weight update, norm of the updated weights
    const T step_aic = stepv[b];
    double nrm = 0.0;
    for (int i = tid; i < M * K; i += NT) {
      const int k = i % K;
      const C2 g = cmulc(Ef[k], Xa[i]);                                     // conj(X) * E
      const T ip = (T)1 / Pa[k];
      C2 wv = Waic[i];
      wv.x += step_aic * (g.x * ip); wv.y += step_aic * (g.y * ip);
      Waic[i] = wv;
      nrm += (double)wv.x * (double)wv.x + (double)wv.y * (double)wv.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) red[warp] = nrm;
    __syncthreads();
    nrm = 0.0;
    for (int q = 0; q < M; ++q) nrm += red[q];
    nrm = nrm / (double)N / (double)N;                                      // :86
    const T sc = (nrm > a.maxnorm) ? (T)sqrt(a.maxnorm / nrm) : (T)1;
    // (e) constraint per channel: irfft, scale, zero the second half, rfft (:92-97)
    C2 *Wm = Waic + (size_t)warp * K;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = lane + 32 * i;
      if (k <= H / 2) {
        if (k == 0) {
          const T p0 = Wm[0].x, ph = Wm[H].x;
          buf[FPAD<T>(0)] = mk2<T>(p0 + ph, -(p0 - ph));
        } else {
          C2 z1, z2;
          irfft_merge_pair<T, C2>(Wm[k], Wm[H - k], tw_n[k], z1, z2);
          buf[FPAD<T>(k)] = z1; buf[FPAD<T>(H - k)] = z2;
        }
      }
    }
    __syncwarp();
    warp_cfft<H, T>(buf, tw_h, lane);
    {
      typedef FftFirst<H, T> F1;                       // scale / zero-tail pass feeds the next transform's first pass from registers
      C2 v[F1::PER][F1::R];
#pragma unroll
      for (int i = 0; i < F1::PER; ++i)
#pragma unroll
        for (int r = 0; r < F1::R; ++r) {
          const int c = lane + 32 * i + r * F1::NB;
          C2 z = mk2<T>((T)0, (T)0);
          if (2 * c < L) { const C2 q = buf[FPAD<T>(c)]; z = mk2<T>(q.x * invN * sc, (-q.y) * invN * sc); }
          v[i][r] = z;
        }
      __syncwarp();
      F1::run(v, buf, tw_h, lane);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int k = lane + 32 * i;
      if (k <= H / 2) {
        if (k == 0) {
          const C2 z = buf[FPAD<T>(0)];
          Wm[0] = mk2<T>(z.x + z.y, (T)0); Wm[H] = mk2<T>(z.x - z.y, (T)0);
        } else {
          C2 x1, x2;
          rfft_split_pair<T, C2>(buf[FPAD<T>(k)], buf[FPAD<T>(H - k)], tw_n[k], x1, x2);
          Wm[k] = x1; Wm[H - k] = x2;
        }
      }
    }
    __syncthreads();
  }
#undef FIDX
  for (int i = tid; i < M * K; i += NT) { st[so.Waic + 2 * i] = (double)Waic[i].x; st[so.Waic + 2 * i + 1] = (double)Waic[i].y; }
  for (int i = tid; i < K; i += NT) st[so.Pa + i] = (double)Pa[i];
}

// tails of the call -> delay-line state (everything the next call's prologue / first block reads)
template <typename T>
__global__ void fd_epilogue_kernel(Fd2Args a) {
  const Fd2State so(a.M);
  const Fd2Ws<T> w(a.S, a.M, a.Ns);
  const T *A = reinterpret_cast<const T *>(a.ws + w.A), *F = reinterpret_cast<const T *>(a.ws + w.F), *B = reinterpret_cast<const T *>(a.ws + w.B);
  const int s = blockIdx.x, M = a.M, Ns = a.Ns;
  double *st = a.state + (size_t)s * so.total;
  for (int i = threadIdx.x; i < M * (F2_L / 2); i += blockDim.x) {
    const int m = i / (F2_L / 2), n = i % (F2_L / 2);
    st[so.dl_al + i] = (double)A[((size_t)s * M + m) * (F2_L / 2 + Ns) + Ns + n];
  }
  for (int i = threadIdx.x; i < F2_L; i += blockDim.x) {
    const double v = (double)F[(size_t)s * (F2_L + Ns) + Ns + i];
    st[so.dl_fbf + i] = v;
    st[so.fbf_prev + i] = v;
    st[so.x0_prev + i] = (double)a.x[(size_t)s * M * Ns + (Ns - F2_L) + i];
  }
  for (int i = threadIdx.x; i < M * F2_L; i += blockDim.x) {
    const int m = i / F2_L, n = i % F2_L;
    st[so.bm_prev + i] = (double)B[((size_t)s * M + m) * (F2_L + Ns) + Ns + n];
  }
  // FIR cache <- last FLMAX-1 samples of concat(cache, x): Ns >= 256 > FLMAX-1, so they all come from x
  for (int i = threadIdx.x; i < M * (F2_FLMAX - 1); i += blockDim.x) {
    const int m = i / (F2_FLMAX - 1), k = i % (F2_FLMAX - 1);
    st[so.cache + i] = (double)a.x[((size_t)s * M + m) * Ns + Ns - (F2_FLMAX - 1) + k];
  }
}

template <typename T>
static int launch_fdgsc2(const Fd2Args &a, const TwiddleSet &tw, cudaStream_t st) {
  typedef typename V2<T>::type C2;
  constexpr int BE = fft_buf_elems(F2_N);
  const int M = a.M;
  fd_prologue_kernel<T><<<a.S, 256, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  {
    const size_t smem = ((size_t)M * (FIR_TS + F2_FLMAX) + (size_t)M * F2_FLMAX) * sizeof(T);
    auto k = fd_fir_kernel<T>;
    DS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((a.Ns + FIR_TS - 1) / FIR_TS, a.S);
    k<<<grid, FIR_NT, smem, st>>>(a);
    DS_LAUNCH_CHECK();
  }
  const size_t tw_bytes = (size_t)(F2_H + F2_H / 2 + 2) * sizeof(C2);
  const long long sb = (long long)a.S * a.nblk;
  {
    const size_t smem = tw_bytes + (size_t)F2_N * sizeof(T) + (size_t)SPEC_WARPS * BE * sizeof(C2);
    auto k = fd_spec_kernel<T>;
    DS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)((sb + SPEC_WARPS - 1) / SPEC_WARPS), SPEC_WARPS * 32, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw));
    DS_LAUNCH_CHECK();
  }
  {
    const long long items = (long long)a.S * F2_K;
    fd_recur_kernel<T><<<(unsigned)((items + 127) / 128), 128, 0, st>>>(a);
    DS_LAUNCH_CHECK();
    fd_control_kernel<T><<<(unsigned)((sb + 7) / 8), 256, 0, st>>>(a);
    DS_LAUNCH_CHECK();
  }
  {
    const size_t smem = tw_bytes + (size_t)BM_WARPS * (BE + F2_K + 1) * sizeof(C2);
    auto k = fd_bm_kernel<T>;
    DS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = (long long)a.S * M;
    k<<<(unsigned)((items + BM_WARPS - 1) / BM_WARPS), BM_WARPS * 32, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw));
    DS_LAUNCH_CHECK();
  }
  {
    const size_t smem = tw_bytes + ((size_t)M * aic_buf_elems<T>() + (size_t)3 * M * F2_K + F2_K + 1) * sizeof(C2) + (size_t)F2_K * sizeof(T) + 16 + 8 * sizeof(double);
    auto k = fd_aic_kernel<T>;
    DS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<a.S, 32 * M, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw));
    DS_LAUNCH_CHECK();
  }
  fd_epilogue_kernel<T><<<a.S, 256, 0, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

int fdgsc_notch_launch(float *x, double *state, int S, int M, int Ns, double r, cudaStream_t st);      // fdgsc.cu

}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_fdgsc_workspace_bytes(const ds_fdgsc_params *p) {
  if (!p || p->n_streams < 1 || p->n_mics < 2 || p->n_samples < F2_L) return 0;
  return p->fp64 ? Fd2Ws<double>(p->n_streams, p->n_mics, p->n_samples).total : Fd2Ws<float>(p->n_streams, p->n_mics, p->n_samples).total;
}

int ds_fdgsc_run_ws(const ds_fdgsc_params *p, const double *delay_filter, const double *window, void *state, void *workspace,
                    float *x, float *y, float *bm_out, float *fix_out, double *p_out, void *stream) {
  DS_CHECK_ARG(p && delay_filter && window && state && workspace && x && y, "ds_fdgsc_run_ws: null argument");
  DS_CHECK_ARG(p->frame_len == F2_L, "ds_fdgsc_run_ws: only frameLen = 256 is compiled");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_mics >= 2 && p->n_mics <= 8, "ds_fdgsc_run_ws: n_mics must be 2..8");
  DS_CHECK_ARG(p->n_samples >= F2_L && p->n_samples % F2_L == 0, "ds_fdgsc_run_ws: n_samples must be a positive multiple of 256");
  DS_CHECK_ARG(p->filter_len >= 1 && p->filter_len <= F2_FLMAX, "ds_fdgsc_run_ws: alignment filter longer than %d taps", F2_FLMAX);
  TwiddleSet tw;
  int rc = get_twiddles(F2_N, &tw);
  if (rc != DS_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dc_notch) {
    rc = fdgsc_notch_launch(x, (double *)state, p->n_streams, p->n_mics, p->n_samples, p->notch_radius, st);
    if (rc != DS_OK) return rc;
  }
  Fd2Args a;
  a.state = (double *)state; a.h = delay_filter; a.x = x; a.y = y; a.bm_out = bm_out; a.fix_out = fix_out; a.p_out = p_out;
  a.window = window; a.ws = (unsigned char *)workspace;
  a.S = p->n_streams; a.M = p->n_mics; a.Ns = p->n_samples; a.FL = p->filter_len; a.frm_cnt = p->frm_cnt; a.ell = p->ell;
  a.nblk = p->n_samples / F2_L;
  a.mu_bm = p->mu_bm; a.mu_aic = p->mu_aic; a.alpha = p->alpha; a.maxnorm = p->maxnorm; a.delta = p->delta;
  a.mc.alpha_d = p->mcra_alpha_d; a.mc.alpha_s = p->mcra_alpha_s; a.mc.delta_s = p->mcra_delta_s;
  a.mc.alpha_p = p->mcra_alpha_p; a.mc.p_min = p->mcra_p_min; a.mc.p_max = p->mcra_p_max; a.mc.L = p->mcra_L;
  return p->fp64 ? launch_fdgsc2<double>(a, tw, st) : launch_fdgsc2<float>(a, tw, st);
}

}  // extern "C"
