// transform.cu -- batched STFT / ISTFT kernels and the fused fixed beamformer.
//
// Reference behaviour restated (file:line relative to the reference tree):
//   stft            DistantSpeech/transform/transform.py:10-221   (complex64 result, :212)
//   istft           transform.py:237-404 + __overlap_add :224-234 (float32 OLA, no window-sum norm)
//   Transform       transform.py:407-496 (streaming history/tail, hop/W0 scaling :479)
//   FixedBeamformer beamformer/fixedbeamformer.py:147-207 (Y = sum_m conj(W) X, :163)
//
// Two families of kernels: the half-warp "square" transform for n_fft = 512 with the fp32 FFT (stft_sq_kernel,
// istft_sq_kernel, fixedbf_sq_kernel; fft16.cuh) -- the shape of every BASELINE configuration on this path -- and the
// Stockham warp transform for every other size and for the fp64 FFT (stft_kernel, istft_seq_kernel / istft_kernel,
// fixedbf_seq_kernel / fixedbf_kernel; fft.cuh).
#include "common.cuh"
#include "fft.cuh"
#include "fft16.cuh"

namespace ds {

// ===========================================================================
// STFT
// ===========================================================================
struct StftArgs {
  const float *x; float *history; void *X; const double *window;
  int S, C, Ns, T, hop, mode, out_c128;
  const short *x16;     // int16 PCM input instead of x (fused load_audio scaling, beamformer/utils.py:184-185)
};

// load_audio's scaling float32(pcm) / 32767.0f (utils.py:184-185; iinfo(int16).max, not 32768) without a division:
// q = x r, rem = fma(-q, 32767, x), q' = fma(rem, r, q) with r = fl(1 / 32767) equals the correctly rounded
// quotient for every int16 value (checked exhaustively on the host, tests/test_pcm_fused_cpu.py).
__device__ __forceinline__ float pcm16_scale(short v) {
  const float x = (float)v, r = 1.0f / 32767.0f;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-q, 32767.0f, x), r, q);
}
// sample g of a row that is float32 or int16 PCM
template <bool PCM> __device__ __forceinline__ float load_sample(const float *xs, const short *xs16, int g) {
  if constexpr (PCM) return pcm16_scale(xs16[g]); else return xs[g];
}

constexpr int STFT_WARPS = 8;

// One warp per frame.  Twiddles and the window are staged once per CTA in shared
// memory; interior frames of 16-byte aligned rows are read with float4 loads
// ([stream, mic, sample] rows are contiguous), everything else (history, reflect
// padding, odd alignment) takes the scalar path.
template <int N, typename T, bool PCM>
__global__ void __launch_bounds__(STFT_WARPS * 32) stft_kernel(StftArgs a, const typename V2<T>::type *__restrict__ tw_h_g,
                                                              const typename V2<T>::type *__restrict__ tw_n_g) {
  typedef typename V2<T>::type C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = N / 2, K = H + 1, BE = fft_buf_elems(N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  C2 *tw_h = reinterpret_cast<C2 *>(smem_raw);          // [H]
  C2 *tw_n = tw_h + H;                                  // [H/2 + 1] (+1 pad)
  T *win = reinterpret_cast<T *>(tw_n + H / 2 + 2);     // [N]
  C2 *buf = reinterpret_cast<C2 *>(win + N) + warp * BE;
  T *fbuf = reinterpret_cast<T *>(buf);
  for (int i = threadIdx.x; i < H; i += blockDim.x) tw_h[i] = tw_h_g[i];
  for (int i = threadIdx.x; i <= H / 2; i += blockDim.x) tw_n[i] = tw_n_g[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) win[i] = (T)a.window[i];
  __syncthreads();
  const unsigned total = (unsigned)a.S * a.C * a.T;      // host guarantees < 2^31
  const int ov = N - a.hop;
  for (unsigned f = blockIdx.x * STFT_WARPS + warp; f < total; f += gridDim.x * STFT_WARPS) {
    const unsigned sc = f / (unsigned)a.T;
    const int t = (int)(f - sc * (unsigned)a.T);
    const float *xs = PCM ? nullptr : a.x + (size_t)sc * a.Ns;
    const short *xs16 = PCM ? a.x16 + (size_t)sc * a.Ns : nullptr;
    int g0;
    if (a.mode == DS_STFT_STREAMING) g0 = t * a.hop - ov;
    else if (a.mode == DS_STFT_CENTER) g0 = t * a.hop - N / 2;
    else g0 = t * a.hop;
    const unsigned s = sc / (unsigned)a.C;
    const unsigned c = sc - s * (unsigned)a.C;
    const size_t o = (((size_t)s * a.T + t) * a.C + c) * K;
    const bool interior = (g0 >= 0) && (g0 + N <= a.Ns) &&
                          (PCM ? ((reinterpret_cast<size_t>(xs16 + g0) & 3) == 0) : ((reinterpret_cast<size_t>(xs + g0) & 7) == 0));
    if (interior) {
      // frame read straight from global memory in first-pass order (coalesced float2 / short2 rows)
      typedef FftFirst<H, T> F1;
      C2 v[F1::PER][F1::R];
      const float2 *src = reinterpret_cast<const float2 *>(xs + g0);
      const short2 *src16 = reinterpret_cast<const short2 *>(xs16 + g0);
      const C2 *w2 = reinterpret_cast<const C2 *>(win);
#pragma unroll
      for (int i = 0; i < F1::PER; ++i) {
        const int j = lane + 32 * i;
        if (F1::NB % 32 == 0 || j < F1::NB) {
#pragma unroll
          for (int r = 0; r < F1::R; ++r) {
            const int e = j + r * F1::NB;
            float2 xv;
            if constexpr (PCM) { const short2 pv = __ldg(src16 + e); xv = make_float2(pcm16_scale(pv.x), pcm16_scale(pv.y)); }
            else xv = __ldg(src + e);
            const C2 wv = w2[e];
            v[i][r] = mk2<T>(mul_rn((T)xv.x, wv.x), mul_rn((T)xv.y, wv.y));
          }
        }
      }
      F1::run(v, buf, tw_h, lane);
    } else {
      const float *hs = a.history ? a.history + (size_t)sc * ov : nullptr;
      for (int n = lane; n < N; n += 32) {
        int g = g0 + n;
        float v;
        if (a.mode == DS_STFT_STREAMING) {
          v = (g < 0) ? hs[ov + g] : load_sample<PCM>(xs, xs16, g);
        } else {
          if (g < 0) g = -g;                              // np.pad(mode="reflect")
          if (g >= a.Ns) g = 2 * (a.Ns - 1) - g;
          v = load_sample<PCM>(xs, xs16, g);
        }
        fbuf[2 * FPAD<T>(n >> 1) + (n & 1)] = (T)v * win[n];
      }
      __syncwarp();
      warp_cfft<H, T>(buf, tw_h, lane);
    }
    // real-FFT split, written straight to the spectrum (complex64 rounding, transform.py:212)
    if (a.out_c128) warp_rfft_split_store<N, T, double2>(buf, tw_n, reinterpret_cast<double2 *>(a.X) + o, lane);
    else warp_rfft_split_store<N, T, float2>(buf, tw_n, reinterpret_cast<float2 *>(a.X) + o, lane);
    __syncwarp();
  }
}

// history <- last `ov` samples of concat(history, x)        (transform.py:451)
template <bool PCM>
__global__ void stft_history_kernel(const float *x, const short *x16, float *history, int Ns, int ov) {
  const long long sc = blockIdx.x;
  const float *xs = PCM ? nullptr : x + sc * Ns;
  const short *xs16 = PCM ? x16 + sc * Ns : nullptr;
  float *hs = history + sc * ov;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = threadIdx.x + i * blockDim.x;
    if (idx < ov) {
      int g = Ns + idx;  // index into concat(history[ov], x[Ns])
      v[i] = (g < ov) ? hs[g] : load_sample<PCM>(xs, xs16, g - ov);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = threadIdx.x + i * blockDim.x;
    if (idx < ov) hs[idx] = v[i];
  }
}

// ---------------------------------------------------------------------------
// n_fft = 512 with the fp32 transform (the headline shape): a HALF-warp per frame, "square" 16 x 16 FFT (fft16.cuh).
// Lane j of a half-warp owns z[j + 16 r] (r = 0..15) in the first radix-16 pass -- loaded straight from global memory,
// 128 contiguous bytes per half-warp and r -- and Z[j + 16 q] after the second one; the 16 x 16 transpose in between is the
// only shared-memory round trip (STS.64 columns, LDS.128 rows, both conflict free at a row stride of 18 elements).  The
// real-FFT split pairs Z[k] with Z[256 - k], which the mirror lane (16 - j) & 15 holds in register 15 - q: two shuffles
// per pair instead of a trip through shared memory.  Window, second-pass twiddles and split twiddles are lane-invariant
// and stay in registers for the life of the (persistent) warp; the two half-warps take consecutive frames of one row, so
// the 50 % overlap of their input is served by L1.  Non-interior frames (history, reflect padding, odd alignment) fill
// the same registers through the scalar loader, so a frame has the same bits whichever way its samples arrived --
// chunked streaming stays bit-identical to one call.
// ---------------------------------------------------------------------------
constexpr int SQ_WARPS = 4;

// real-FFT split of the pairs (k, 256 - k), k = j + 16 q < 128, written straight to the spectrum
template <typename OutC>
__device__ __forceinline__ void sq_split_store(const float2 (&u)[16], const float2 (&tws)[8], OutC *__restrict__ out, int j, int mirror,
                                               bool valid) {
  constexpr int H = 256;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float2 za = u[P16(q)];
    float2 zb;
    zb.x = __shfl_sync(0xffffffffu, u[P16(15 - q)].x, mirror);
    zb.y = __shfl_sync(0xffffffffu, u[P16(15 - q)].y, mirror);
    const float2 own = u[P16((16 - q) & 15)];                // lane 0 pairs with itself: Z[16 q] <-> Z[16 (16 - q)]
    zb.x = (j == 0) ? own.x : zb.x;
    zb.y = (j == 0) ? own.y : zb.y;
    float2 X1, X2;
    rfft_split_pair<float>(za, zb, tws[q], X1, X2);
    OutC o1, o2;
    o1.x = X1.x; o1.y = X1.y; o2.x = X2.x; o2.y = X2.y;
    if (valid) { out[j + 16 * q] = o1; out[H - j - 16 * q] = o2; }
  }
  if (j == 0 && valid) {                                    // k = 128 pairs with itself
    const float2 za = u[P16(8)];
    float2 X1, X2;
    rfft_split_pair<float>(za, za, make_float2(0.f, -1.f), X1, X2);
    OutC o1;
    o1.x = X1.x; o1.y = X1.y;
    out[H / 2] = o1;
  }
}

#ifndef SQ_MINB
#define SQ_MINB 3          // CTAs (of SQ_WARPS warps) per SM the register budget is set for
#endif
// Measured on the B200 (config 4, 1024 streams x 10 s x 8 mics, analysis pass of ds_chain_run_profiled; Stockham kernel 4.15 ms;
// profiles/ab_runs_r02.txt): constants in registers, 12 warps/SM 3.74 ms; window + split twiddles in shared memory, 16 warps
// 3.35 ms; all constants in shared memory, 20 warps 3.15 ms (24 warps 3.36 ms); register prefetch of the next pair, 8 warps
// 3.53 ms; prefetch + window / split twiddles in shared memory, 12 warps **3.00 ms** (default); prefetch + everything in shared
// memory, 12 warps 3.04 ms, 16 warps 3.23 ms; prefetch.global.L2 instead of the register prefetch 3.4 - 3.7 ms.  15.7 GB in
// 3.00 ms = 5.2 TB/s = 80 % of the measured copy peak: what is left is the HBM system, not the transform.
#ifndef SQ_CONST_SMEM
#define SQ_CONST_SMEM 1    // 1: window and split twiddles from shared memory; 2: second-pass twiddles too; 0: all in registers
#endif
#ifndef SQ_REGPF
#define SQ_REGPF 1         // 1: the next frame pair's samples are loaded (into registers) before the current pair is transformed
#endif
#ifndef SQ_L2PF
#define SQ_L2PF 0          // 1: prefetch.global.L2 of the next frame pair's samples
#endif

// samples of frame pair p for this lane: raw[r] = (x[2 e], x[2 e + 1]), e = j + 16 r, before the window; interior frames of
// aligned rows by coalesced 8-byte (4-byte for int16 PCM) loads, everything else through the scalar loader
template <bool PCM>
__device__ __forceinline__ void sq_load(const StftArgs &a, unsigned p, unsigned Tp, int half, int j, float2 (&raw)[16], bool &valid,
                                        size_t &o) {
  constexpr int N = 512, K = 257;
  const int ov = N - a.hop;
  const unsigned sc = p / Tp;
  const int t = 2 * (int)(p - sc * Tp) + half;
  valid = t < a.T;
  const float *xs = PCM ? nullptr : a.x + (size_t)sc * a.Ns;
  const short *xs16 = PCM ? a.x16 + (size_t)sc * a.Ns : nullptr;
  int g0;
  if (a.mode == DS_STFT_STREAMING) g0 = t * a.hop - ov;
  else if (a.mode == DS_STFT_CENTER) g0 = t * a.hop - N / 2;
  else g0 = t * a.hop;
  const unsigned s = sc / (unsigned)a.C;
  const unsigned c = sc - s * (unsigned)a.C;
  o = (((size_t)s * a.T + (valid ? t : 0)) * a.C + c) * K;
  const bool interior = valid && (g0 >= 0) && (g0 + N <= a.Ns) &&
                        (PCM ? ((reinterpret_cast<size_t>(xs16 + g0) & 3) == 0) : ((reinterpret_cast<size_t>(xs + g0) & 7) == 0));
  if (interior) {
    const float2 *src = reinterpret_cast<const float2 *>(xs + g0) + j;
    const short2 *src16 = reinterpret_cast<const short2 *>(xs16 + g0) + j;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if constexpr (PCM) { const short2 pv = __ldg(src16 + 16 * r); raw[r] = make_float2(pcm16_scale(pv.x), pcm16_scale(pv.y)); }
      else raw[r] = __ldg(src + 16 * r);
    }
  } else if (valid) {
    const float *hs = a.history ? a.history + (size_t)sc * ov : nullptr;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float sv[2];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        int g = g0 + 2 * (j + 16 * r) + cc;
        if (a.mode == DS_STFT_STREAMING) {
          sv[cc] = (g < 0) ? hs[ov + g] : load_sample<PCM>(xs, xs16, g);
        } else {
          if (g < 0) g = -g;                              // np.pad(mode="reflect")
          if (g >= a.Ns) g = 2 * (a.Ns - 1) - g;
          sv[cc] = load_sample<PCM>(xs, xs16, g);
        }
      }
      raw[r] = make_float2(sv[0], sv[1]);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 16; ++r) raw[r] = make_float2(0.f, 0.f);
  }
}

template <bool PCM>
__global__ void __launch_bounds__(SQ_WARPS * 32, SQ_MINB) stft_sq_kernel(StftArgs a, const float2 *__restrict__ tw_h_g,
                                                                        const float2 *__restrict__ tw_n_g) {
  constexpr int H = 256;
  __shared__ __align__(16) float2 xbuf[SQ_WARPS][2 * 16 * SQ_RS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, j = lane & 15;
  float2 *xb = xbuf[warp] + half * 16 * SQ_RS;
  // lane-invariant constants: registers (default) or shared memory (more resident warps)
#if SQ_CONST_SMEM >= 1
  __shared__ float2 s_win[H], s_tws[H / 2];
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_win[i] = make_float2((float)a.window[2 * i], (float)a.window[2 * i + 1]);
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) s_tws[i] = tw_n_g[i];
#else
  float2 win[16], tws[8];
#pragma unroll
  for (int r = 0; r < 16; ++r) win[r] = make_float2((float)a.window[2 * (j + 16 * r)], (float)a.window[2 * (j + 16 * r) + 1]);
#pragma unroll
  for (int q = 0; q < 8; ++q) tws[q] = tw_n_g[j + 16 * q];
#endif
#if SQ_CONST_SMEM >= 2
  __shared__ float2 s_tw2[H];                               // [r][j] = W256^(r j)
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_tw2[i] = tw_h_g[((i >> 4) * (i & 15)) & (H - 1)];
#else
  float2 tw2[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) tw2[r] = tw_h_g[(r * j) & (H - 1)];
#endif
#if SQ_CONST_SMEM >= 1
  __syncthreads();
#endif
  const unsigned Tp = ((unsigned)a.T + 1u) >> 1;
  const unsigned total = (unsigned)a.S * a.C * Tp;       // host guarantees S * C * T < 2^31
  const int mirror = (lane & 16) | ((16 - j) & 15);
  const unsigned stride = gridDim.x * SQ_WARPS;
  unsigned p = blockIdx.x * SQ_WARPS + warp;
  float2 raw[16];
  bool valid = false;
  size_t o = 0;
#if SQ_REGPF
  if (p < total) sq_load<PCM>(a, p, Tp, half, j, raw, valid, o);
#endif
  for (; p < total; p += stride) {
#if !SQ_REGPF
    sq_load<PCM>(a, p, Tp, half, j, raw, valid, o);
#endif
#if SQ_L2PF
    if (p + stride < total) {
      // the next pair's samples (3 KB for float32 at hop 256) towards L2: one 128-byte line per lane of the first 24
      const unsigned pn = p + stride, scn = pn / Tp;
      const int tn = 2 * (int)(pn - scn * Tp);
      const long long gn = (long long)scn * a.Ns + (long long)tn * a.hop + (long long)lane * (PCM ? 64 : 32);
      if (lane < 24 && gn + 32 < (long long)a.S * a.C * a.Ns) {
        if constexpr (PCM) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x16 + gn));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x + gn));
      }
    }
#endif
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#if SQ_CONST_SMEM >= 1
      const float2 wv = s_win[j + 16 * r];
#else
      const float2 wv = win[r];
#endif
      v[r] = make_float2(mul_rn(raw[r].x, wv.x), mul_rn(raw[r].y, wv.y));
    }
    const bool cur_valid = valid;
    const size_t cur_o = o;
#if SQ_REGPF
    if (p + stride < total) sq_load<PCM>(a, p + stride, Tp, half, j, raw, valid, o);
#endif
    // pass 1: V_j[q] = sum_r z[j + 16 r] W16^(r q), stored transposed at [q][j]
    dft16(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) xb[q * SQ_RS + j] = v[P16(q)];
    __syncwarp();
    // pass 2: lane j reads row j (V_r[j], r = 0..15), applies W256^(r j) and transforms over r: Z[j + 16 q]
    float2 u[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 w4 = *reinterpret_cast<const float4 *>(xb + j * SQ_RS + 2 * i);
      u[2 * i] = make_float2(w4.x, w4.y); u[2 * i + 1] = make_float2(w4.z, w4.w);
    }
    __syncwarp();
#pragma unroll
    for (int r = 1; r < 16; ++r) {
#if SQ_CONST_SMEM >= 2
      u[r] = cmul(u[r], s_tw2[r * 16 + j]);
#else
      u[r] = cmul(u[r], tw2[r]);
#endif
    }
    dft16(u);
#if SQ_CONST_SMEM >= 1
    float2 tws[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) tws[q] = s_tws[j + 16 * q];
#endif
    if (a.out_c128) sq_split_store(u, tws, reinterpret_cast<double2 *>(a.X) + cur_o, j, mirror, cur_valid);
    else sq_split_store(u, tws, reinterpret_cast<float2 *>(a.X) + cur_o, j, mirror, cur_valid);
  }
}

template <typename T> struct SqStft {
  static bool applies(int, const StftArgs &) { return false; }
  static int launch(const StftArgs &, const TwiddleSet &, cudaStream_t) { return DS_EUNSUPPORTED; }
};
template <> struct SqStft<float> {
  static bool applies(int n_fft, const StftArgs &) {
#ifdef DS_STFT_NO_SQ
    return false;
#else
    return n_fft == 512;
#endif
  }
  static int launch(const StftArgs &a, const TwiddleSet &tw, cudaStream_t st) {
    const long long total = (long long)a.S * a.C * a.T;
    if (total >= (1LL << 31)) { set_error("stft: more than 2^31 frames in one call"); return DS_EUNSUPPORTED; }
    const long long pairs = (long long)a.S * a.C * ((a.T + 1) / 2);
    long long blocks = (pairs + SQ_WARPS - 1) / SQ_WARPS;
    if (blocks > 148LL * SQ_MINB) blocks = 148LL * SQ_MINB;
    if (blocks < 1) blocks = 1;
    if (a.x16) stft_sq_kernel<true><<<(unsigned)blocks, SQ_WARPS * 32, 0, st>>>(a, tw.h32, tw.n32);
    else stft_sq_kernel<false><<<(unsigned)blocks, SQ_WARPS * 32, 0, st>>>(a, tw.h32, tw.n32);
    DS_LAUNCH_CHECK();
    return DS_OK;
  }
};

template <int N, typename T>
static int launch_stft(const StftArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  typedef typename V2<T>::type C2;
  constexpr int H = N / 2;
  const size_t smem = (size_t)(H + H / 2 + 2) * sizeof(C2) + (size_t)N * sizeof(T) + (size_t)STFT_WARPS * fft_buf_elems(N) * sizeof(C2);
  auto kern = a.x16 ? stft_kernel<N, T, true> : stft_kernel<N, T, false>;
  if (smem > 48 * 1024) DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)a.S * a.C * a.T;
  if (total >= (1LL << 31)) { set_error("stft: more than 2^31 frames in one call"); return DS_EUNSUPPORTED; }
  long long blocks = (total + STFT_WARPS - 1) / STFT_WARPS;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, STFT_WARPS * 32, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw));
  DS_LAUNCH_CHECK();
  return DS_OK;
}

template <typename T>
static int dispatch_stft(int n_fft, const StftArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  if (SqStft<T>::applies(n_fft, a)) return SqStft<T>::launch(a, tw, st);
  switch (n_fft) {
    case 128: return launch_stft<128, T>(a, tw, st);
    case 256: return launch_stft<256, T>(a, tw, st);
    case 512: return launch_stft<512, T>(a, tw, st);
    case 1024: return launch_stft<1024, T>(a, tw, st);
    case 2048: return launch_stft<2048, T>(a, tw, st);
  }
  set_error("unsupported n_fft %d", n_fft);
  return DS_EUNSUPPORTED;
}

// ===========================================================================
// ISTFT
// ===========================================================================
struct IstftArgs {
  const void *Y; float *y; float *tail; const double *window;
  int S, C, T, hop, mode, n_out, in_c128;  // n_out = samples written per (s,c)
  double scale;
  short *y16;       // int16 PCM output instead of y (fused save_audio scaling, beamformer/utils.py:193)
  long long sS, sT, sC, sK;   // element strides of Y over (stream, frame, channel, bin); default [S][T][C][K] contiguous
};

// save_audio on the float32 sample the float path would have stored: (audio * 32767).astype(int16) with the product in
// double (Transform.istft hands save_audio a float64 array, transform.py:474-481) and the cast truncating toward zero.
// Out-of-range products saturate (NumPy leaves that cast undefined).
__device__ __forceinline__ short pcm16_from_sample(float v) {
  const double p = (double)v * 32767.0;
  return (short)max(-32768, min(32767, __double2int_rz(p)));
}

constexpr int ISTFT_WARPS = 4;
constexpr int ISTFT_TILE = 8;   // output hop-blocks per CTA
constexpr int ISTFT_MAXR = 8;   // max overlapping frames per sample

// Windowed inverse transform of frame t of (s,c) into dst[0..N) (float).
template <int N, typename T>
__device__ __forceinline__ void istft_frame(const IstftArgs &a, int s, int c, int t, typename V2<T>::type *buf,
                                            const typename V2<T>::type *tw_h, const typename V2<T>::type *tw_n,
                                            float *dst, int lane) {
  typedef typename V2<T>::type C2;
  constexpr int H = N / 2, K = H + 1;
  const long long ibase = (long long)s * a.sS + (long long)t * a.sT + (long long)c * a.sC;
  if (a.in_c128) {
    const double2 *in = reinterpret_cast<const double2 *>(a.Y) + ibase;
    for (int k = lane; k < K; k += 32) { double2 v = in[k * a.sK]; buf[FPAD<T>(k)] = mk2<T>((T)v.x, (T)v.y); }
  } else {
    const float2 *in = reinterpret_cast<const float2 *>(a.Y) + ibase;
    for (int k = lane; k < K; k += 32) { float2 v = in[k * a.sK]; buf[FPAD<T>(k)] = mk2<T>((T)v.x, (T)v.y); }
  }
  __syncwarp();
  warp_irfft_unscaled<N, T>(buf, tw_h, tw_n, lane);
  const T *fb = reinterpret_cast<const T *>(buf);
  const T inv_n = (T)1 / (T)N;
  for (int n = lane; n < N; n += 32) {
    T v = fb[2 * FPAD<T>(n >> 1) + (n & 1)] * inv_n;      // numpy irfft value
    dst[n] = (float)((double)v * a.window[n]);         // ifft_window * irfft   (:368)
  }
  __syncwarp();
}

// One CTA = (s, c, tile of ISTFT_TILE hop-blocks).  Frames that touch the tile
// are inverse-transformed into shared memory (halo frames recomputed), then each
// output sample adds its frames in increasing frame order in float32, exactly
// like the reference's float32 += accumulation.
template <int N, typename T>
__global__ void __launch_bounds__(ISTFT_WARPS * 32) istft_kernel(IstftArgs a, const typename V2<T>::type *__restrict__ tw_h,
                                                                const typename V2<T>::type *__restrict__ tw_n,
                                                                int R, int tail_pass) {
  typedef typename V2<T>::type C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  C2 *fftbuf = reinterpret_cast<C2 *>(smem_raw) + warp * fft_buf_elems(N);
  float *frames = reinterpret_cast<float *>(reinterpret_cast<C2 *>(smem_raw) + ISTFT_WARPS * fft_buf_elems(N));
  const int ov = N - a.hop;
  const int sc = blockIdx.y;
  const int s = sc / a.C, c = sc % a.C;
  const int total_len = a.T * a.hop + ov;                     // OLA buffer length
  // tile of hop-blocks [b0, b1)
  int b0, b1;
  if (!tail_pass) {
    b0 = blockIdx.x * ISTFT_TILE;
    b1 = min(b0 + ISTFT_TILE, (a.n_out + a.hop - 1) / a.hop);
  } else {  // blocks covering [T*hop, T*hop+ov)
    b0 = a.T;
    b1 = (total_len + a.hop - 1) / a.hop;
  }
  const int tf0 = max(0, b0 - R + 1), tf1 = min(a.T - 1, b1 - 1);   // frames needed (inclusive)
  for (int t = tf0 + warp; t <= tf1; t += ISTFT_WARPS)
    istft_frame<N, T>(a, s, c, t, fftbuf, tw_h, tw_n, frames + (size_t)(t - tf0) * N, lane);
  __syncthreads();
  const int g0 = b0 * a.hop;
  const int g1 = tail_pass ? total_len : min(b1 * a.hop, a.n_out);
  float *ys = a.y16 ? nullptr : a.y + (long long)sc * a.n_out;
  short *ys16 = a.y16 ? a.y16 + (long long)sc * a.n_out : nullptr;
  float *tl = a.tail ? a.tail + (long long)sc * ov : nullptr;
  // tail pass: every thread first computes its values (reads old tail), then all write
  float keep[(ISTFT_MAXR * 2048 / 8) / (ISTFT_WARPS * 32) + 1];
  int nkeep = 0;
  for (int g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
    int t_hi = min(a.T - 1, g / a.hop);
    int t_lo = (g - N + 1 + a.hop - 1);
    t_lo = (t_lo <= 0) ? 0 : t_lo / a.hop;
    float acc = 0.0f;
    for (int t = t_lo; t <= t_hi; ++t) acc = acc + frames[(size_t)(t - tf0) * N + (g - t * a.hop)];
    if (a.mode == DS_STFT_STREAMING) {
      if (g < ov) acc = acc + tl[g];                          // x[:overlap] += previous_output (:476)
      if (!tail_pass) {                                       // :479
        const float o = (float)((double)acc * a.scale);
        if (ys16) ys16[g] = pcm16_from_sample(o); else ys[g] = o;
      } else keep[nkeep++] = acc;
    } else {
      if (ys16) ys16[g] = pcm16_from_sample(acc); else ys[g] = acc;
    }
  }
  if (tail_pass) {
    __syncthreads();
    nkeep = 0;
    for (int g = g0 + threadIdx.x; g < g1; g += blockDim.x) tl[g - a.T * a.hop] = keep[nkeep++];
  }
}

// hop = N/2 fast path: one warp walks a run of consecutive hop-blocks of one (stream, channel), keeping the
// second half of the previous frame in shared memory -- every frame is inverse-transformed once (plus one
// halo frame per run), there is no CTA-wide synchronisation, and output rows are written as they complete.
// Same arithmetic and float32 accumulation order as istft_kernel: acc = (0 + f[b-1]) + f[b] (+ tail).
constexpr int ISEQ_WARPS = 8;

template <int N, typename T>
__global__ void __launch_bounds__(ISEQ_WARPS * 32) istft_seq_kernel(IstftArgs a, const typename V2<T>::type *__restrict__ tw_h_g,
                                                                   const typename V2<T>::type *__restrict__ tw_n_g, int G, int nseg) {
  typedef typename V2<T>::type C2;
  constexpr int H = N / 2, K = H + 1, BE = fft_buf_elems(N), HOP = N / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *wind = reinterpret_cast<double *>(smem_raw);                    // [N]
  C2 *tw_h = reinterpret_cast<C2 *>(wind + N);                            // [H]
  C2 *tw_n = tw_h + H;                                                    // [H/2 + 1] (+1 pad)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  C2 *buf = tw_n + H / 2 + 2 + warp * BE;
  float *prev = reinterpret_cast<float *>(tw_n + H / 2 + 2 + ISEQ_WARPS * BE) + warp * HOP;
  for (int i = threadIdx.x; i < N; i += blockDim.x) wind[i] = a.window[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) tw_h[i] = tw_h_g[i];
  for (int i = threadIdx.x; i <= H / 2; i += blockDim.x) tw_n[i] = tw_n_g[i];
  __syncthreads();
  const long long w = (long long)blockIdx.x * ISEQ_WARPS + warp;
  if (w >= (long long)a.S * a.C * nseg) return;
  const int sc = (int)(w / nseg), seg = (int)(w % nseg);
  const int s = sc / a.C, c = sc % a.C;
  const int nblk = (a.n_out + HOP - 1) / HOP;
  const int b0 = seg * G, b1 = min(b0 + G, nblk);
  const T *fb = reinterpret_cast<const T *>(buf);
  const T inv_n = (T)1 / (T)N;
  float *ys = a.y16 ? nullptr : a.y + (long long)sc * a.n_out;
  short *ys16 = a.y16 ? a.y16 + (long long)sc * a.n_out : nullptr;
  const bool streaming = a.mode == DS_STFT_STREAMING;

  auto inverse = [&](int t) {          // frame t -> buf holds irfft * N
    const long long ibase = (long long)s * a.sS + (long long)t * a.sT + (long long)c * a.sC;
    if (a.in_c128) {
      const double2 *in = reinterpret_cast<const double2 *>(a.Y) + ibase;
      for (int k = lane; k < K; k += 32) { const double2 v = in[k * a.sK]; buf[FPAD<T>(k)] = mk2<T>((T)v.x, (T)v.y); }
    } else {
      const float2 *in = reinterpret_cast<const float2 *>(a.Y) + ibase;
      for (int k = lane; k < K; k += 32) { const float2 v = in[k * a.sK]; buf[FPAD<T>(k)] = mk2<T>((T)v.x, (T)v.y); }
    }
    __syncwarp();
    warp_irfft_unscaled<N, T>(buf, tw_h, tw_n, lane);
  };
  auto sample = [&](int n) -> float {  // windowed sample n of the frame in buf      (:368)
    const T v = fb[2 * FPAD<T>(n >> 1) + (n & 1)] * inv_n;
    return (float)((double)v * wind[n]);
  };

  if (b0 >= 1) {                       // halo: second half of frame b0-1
    inverse(b0 - 1);
    for (int j = lane; j < HOP; j += 32) prev[j] = sample(HOP + j);
    __syncwarp();
  }
  for (int b = b0; b < b1; ++b) {
    if (b < a.T) inverse(b);
    for (int j = lane; j < HOP; j += 32) {
      const int g = b * HOP + j;
      float acc = 0.0f;
      if (b >= 1) acc = acc + prev[j];
      if (b < a.T) { acc = acc + sample(j); prev[j] = sample(HOP + j); }
      if (g < a.n_out) {
        if (streaming) {
          if (b == 0) acc = acc + a.tail[(long long)sc * HOP + j];           // x[:overlap] += previous_output (:476)
          acc = (float)((double)acc * a.scale);                             // :479
        }
        if (ys16) ys16[g] = pcm16_from_sample(acc); else ys[g] = acc;
      }
    }
    __syncwarp();
  }
}

// n_fft = 512, hop = 256, fp32 transform: the inverse STFT on the half-warp "square" transform (fft16.cuh).  A warp owns a run
// of consecutive hop-blocks of one (stream, channel) and takes the frames two at a time (one per half-warp): merge of the bin
// pairs (k, 256 - k) (the partner half of Z from the mirror lane by shuffles), the forward transform of conj(Z), the
// synthesis window in double like the reference's float64 product (transform.py:368), and the overlap-add in registers --
// the first frame's second half reaches the second half-warp by shuffles, the second frame's is carried to the next pair.
// tail_only: a second launch (one warp per sequence) recomputes the last frame and stores its second half as the new
// streaming tail -- the same arithmetic as the main pass, so chunked streaming stays bit-identical to one call.
constexpr int ISQ_WARPS = 4;

__global__ void __launch_bounds__(ISQ_WARPS * 32, 3) istft_sq_kernel(IstftArgs a, const float2 *__restrict__ tw_h_g,
                                                                    const float2 *__restrict__ tw_n_g, int G, int nseg, int tail_only) {
  constexpr int N = 512, H = 256, HOP = 256;
  __shared__ __align__(16) float2 xbuf[ISQ_WARPS][2 * 16 * SQ_RS];
  __shared__ double2 s_win[H];                                          // (w[2e], w[2e+1])
  __shared__ float2 s_tws[H / 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, j = lane & 15;
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_win[i] = make_double2(a.window[2 * i], a.window[2 * i + 1]);
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) s_tws[i] = tw_n_g[i];
  float2 tw2[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) tw2[r] = tw_h_g[(r * j) & (H - 1)];
  __syncthreads();
  const long long w = (long long)blockIdx.x * ISQ_WARPS + warp;
  if (w >= (long long)a.S * a.C * nseg) return;                          // no CTA-wide synchronisation below
  float2 *xb = xbuf[warp] + half * 16 * SQ_RS;
  const int sc = (int)(w / nseg), seg = (int)(w % nseg);
  const int s = sc / a.C, c = sc % a.C;
  const int nblk = (a.n_out + HOP - 1) / HOP;
  const int b0 = tail_only ? a.T : seg * G, b1 = tail_only ? a.T : min(b0 + G, nblk);
  const float inv_n = 1.0f / (float)N;
  float *ys = a.y16 ? nullptr : a.y + (long long)sc * a.n_out;
  short *ys16 = a.y16 ? a.y16 + (long long)sc * a.n_out : nullptr;
  const bool streaming = a.mode == DS_STFT_STREAMING;
  const int mirror = (lane & 16) | ((16 - j) & 15);
  // pv: second half of the frame before the pair, as this lane's samples 2 (j + 16 q), +1 of the block
  float2 pv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    pv[q] = (b0 == 0 && streaming) ? *reinterpret_cast<const float2 *>(a.tail + (long long)sc * HOP + 2 * (j + 16 * q))   // x[:overlap] += previous_output (:476)
                                   : make_float2(0.f, 0.f);
  const int f_end = tail_only ? a.T : min(b1, a.T + 1);                   // blocks / frames f < f_end are visited
  for (int tt = (b0 >= 1 ? b0 - 1 : 0); tt < f_end; tt += 2) {
    const int f = tt + half;
    const bool fvalid = f < a.T && f < f_end;                             // frame f exists (block T of the plain mode has only a carry)
    float2 z[16], Z2[8], r2[16];
    float2 ya[8], yb[8], yq = make_float2(0.f, 0.f);
    if (fvalid) {
      const long long ibase = (long long)s * a.sS + (long long)f * a.sT + (long long)c * a.sC;
      if (a.in_c128) {
        const double2 *in = reinterpret_cast<const double2 *>(a.Y) + ibase;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double2 v1 = in[(long long)(j + 16 * q) * a.sK], v2 = in[(long long)(H - j - 16 * q) * a.sK];
          ya[q] = make_float2((float)v1.x, (float)v1.y); yb[q] = make_float2((float)v2.x, (float)v2.y);
        }
        if (j == 0) { const double2 v = in[(long long)(H / 2) * a.sK]; yq = make_float2((float)v.x, (float)v.y); }
      } else {
        const float2 *in = reinterpret_cast<const float2 *>(a.Y) + ibase;
#pragma unroll
        for (int q = 0; q < 8; ++q) { ya[q] = in[(long long)(j + 16 * q) * a.sK]; yb[q] = in[(long long)(H - j - 16 * q) * a.sK]; }
        if (j == 0) yq = in[(long long)(H / 2) * a.sK];
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) { ya[q] = make_float2(0.f, 0.f); yb[q] = make_float2(0.f, 0.f); }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) irfft_merge_pair<float>(ya[q], yb[q], s_tws[j + 16 * q], z[q], Z2[q]);
    float2 zq, zq2;
    irfft_merge_pair<float>(yq, yq, make_float2(0.f, -1.f), zq, zq2);
#pragma unroll
    for (int r = 8; r < 16; ++r) {
      float2 rc;
      rc.x = __shfl_sync(0xffffffffu, Z2[15 - r].x, mirror);
      rc.y = __shfl_sync(0xffffffffu, Z2[15 - r].y, mirror);
      const float2 own = (r == 8) ? zq : Z2[(16 - r) & 7];
      z[r].x = (j == 0) ? own.x : rc.x;
      z[r].y = (j == 0) ? own.y : rc.y;
    }
    if (j == 0) {                                                       // bins 0 and 256 (imaginary parts ignored like numpy.fft.irfft)
      const float a0 = ya[0].x, b0v = yb[0].x;
      z[0] = make_float2(a0 + b0v, -(a0 - b0v));
    }
    sq_cfft256(z, r2, xb, tw2, j);
    // r2[P16(q)] = conj(x[2n] + i x[2n+1]) * N, n = j + 16 q; windowed sample (:368)
    float2 cw[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const double2 wv = s_win[j + 16 * q];
      cw[q] = make_float2((float)((double)(r2[P16(q)].x * inv_n) * wv.x), (float)((double)(-r2[P16(q)].y * inv_n) * wv.y));
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float2 s0, nx;
      s0.x = __shfl_sync(0xffffffffu, cw[q + 8].x, j);                   // second half of the first frame of the pair
      s0.y = __shfl_sync(0xffffffffu, cw[q + 8].y, j);
      nx.x = __shfl_sync(0xffffffffu, cw[q + 8].x, j | 16);              // second half of the second frame
      nx.y = __shfl_sync(0xffffffffu, cw[q + 8].y, j | 16);
      const float2 ad = half ? s0 : pv[q];
      float vx = ad.x + cw[q].x, vy = ad.y + cw[q].y;                    // (0 + f[b-1]) + f[b]
      const int g = f * HOP + 2 * (j + 16 * q);
      if (f >= b0 && f < b1 && g < a.n_out) {                            // n_out is a multiple of the hop: both samples or none
        if (streaming) { vx = (float)((double)vx * a.scale); vy = (float)((double)vy * a.scale); }   // :479
        if (ys16) *reinterpret_cast<short2 *>(ys16 + g) = make_short2(pcm16_from_sample(vx), pcm16_from_sample(vy));
        else *reinterpret_cast<float2 *>(ys + g) = make_float2(vx, vy);
      }
      pv[q] = (tt + 1 < f_end) ? nx : s0;
    }
  }
  if (tail_only && half == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<float2 *>(a.tail + (long long)sc * HOP + 2 * (j + 16 * q)) = pv[q];
  }
}

template <int N, typename T>
static int launch_istft(const IstftArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  const int R = (N + a.hop - 1) / a.hop;
  if (R > ISTFT_MAXR) { set_error("istft: hop %d < n_fft/8 unsupported", a.hop); return DS_EUNSUPPORTED; }
  const int max_frames = ISTFT_TILE + R - 1 + R;   // tail pass may need up to 2R-2 frames
  const size_t smem = (size_t)ISTFT_WARPS * fft_buf_elems(N) * sizeof(typename V2<T>::type) + (size_t)max_frames * N * sizeof(float);
  auto kern = istft_kernel<N, T>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = (a.n_out + a.hop - 1) / a.hop;
#ifndef DS_ISTFT_NO_SQ
  if constexpr (N == 512 && sizeof(T) == 4) {
    // the square kernel stores sample pairs: 8-byte aligned float32 (4-byte aligned int16) output and tail rows
    const bool aligned = (a.y16 ? (reinterpret_cast<size_t>(a.y16) & 3) == 0 : (reinterpret_cast<size_t>(a.y) & 7) == 0) &&
                         (reinterpret_cast<size_t>(a.tail) & 7) == 0;
    if (aligned && a.hop * 2 == N && (a.mode != DS_STFT_STREAMING || a.tail)) {
      const long long seqs = (long long)a.S * a.C;
      long long nseg = (2LL * 148 * 16 + seqs - 1) / seqs;
      const long long max_seg = (nblk + 15) / 16;
      if (nseg > max_seg) nseg = max_seg;
      if (nseg < 1) nseg = 1;
      const int G = (int)((nblk + nseg - 1) / nseg);
      nseg = (nblk + G - 1) / G;
      const long long warps = seqs * nseg;
      istft_sq_kernel<<<(unsigned)((warps + ISQ_WARPS - 1) / ISQ_WARPS), ISQ_WARPS * 32, 0, st>>>(a, tw.h32, tw.n32, G, (int)nseg, 0);
      DS_LAUNCH_CHECK();
      if (a.mode == DS_STFT_STREAMING && a.tail) {
        istft_sq_kernel<<<(unsigned)((seqs + ISQ_WARPS - 1) / ISQ_WARPS), ISQ_WARPS * 32, 0, st>>>(a, tw.h32, tw.n32, 0, 1, 1);
        DS_LAUNCH_CHECK();
      }
      return DS_OK;
    }
  }
#endif
  if (a.hop * 2 == N && (a.mode != DS_STFT_STREAMING || a.tail)) {
    // one warp per run of G hop-blocks; enough runs to fill the machine twice over
    typedef typename V2<T>::type C2;
    const long long seqs = (long long)a.S * a.C;
    long long nseg = (2LL * 148 * 32 + seqs - 1) / seqs;
    const long long max_seg = (nblk + 7) / 8;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    const int G = (int)((nblk + nseg - 1) / nseg);
    nseg = (nblk + G - 1) / G;
    const size_t smem2 = (size_t)N * sizeof(double) + (size_t)(N / 2 + N / 4 + 2) * sizeof(C2) +
                         (size_t)ISEQ_WARPS * (fft_buf_elems(N) * sizeof(C2) + (N / 2) * sizeof(float));
    auto kseq = istft_seq_kernel<N, T>;
    DS_CUDA(cudaFuncSetAttribute(kseq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const long long warps = seqs * nseg;
    kseq<<<(unsigned)((warps + ISEQ_WARPS - 1) / ISEQ_WARPS), ISEQ_WARPS * 32, smem2, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw), G, (int)nseg);
    DS_LAUNCH_CHECK();
  } else {
    dim3 grid((nblk + ISTFT_TILE - 1) / ISTFT_TILE, a.S * a.C);
    if (grid.x > 0 && grid.y > 0) {
      kern<<<grid, ISTFT_WARPS * 32, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw), R, 0);
      DS_LAUNCH_CHECK();
    }
  }
  if (a.mode == DS_STFT_STREAMING && a.tail && N > a.hop) {
    dim3 g2(1, a.S * a.C);
    kern<<<g2, ISTFT_WARPS * 32, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw), R, 1);
    DS_LAUNCH_CHECK();
  }
  return DS_OK;
}

template <typename T>
static int dispatch_istft(int n_fft, const IstftArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  switch (n_fft) {
    case 128: return launch_istft<128, T>(a, tw, st);
    case 256: return launch_istft<256, T>(a, tw, st);
    case 512: return launch_istft<512, T>(a, tw, st);
    case 1024: return launch_istft<1024, T>(a, tw, st);
    case 2048: return launch_istft<2048, T>(a, tw, st);
  }
  set_error("unsupported n_fft %d", n_fft);
  return DS_EUNSUPPORTED;
}

// ===========================================================================
// fused fixed beamformer:  STFT -> sum_m conj(W) X -> ISTFT/OLA, one kernel
// ===========================================================================
// State blob: int parity (+pad to 16 B), then two halves, each
//   history [S][M][ov] float32 ; tail [S][B][ov] float32.
// The kernel reads half `parity`, writes half `1-parity`; a trailing 1-thread
// kernel flips the parity, so segments of one stream never race on the state.
struct FixedBfArgs {
  const float *x; float *y; const float2 *W; const double *window;
  unsigned char *state;
  int S, M, B, Ns, T, hop, segs, frames_per_seg;
  double scale;
};

constexpr int FBF_WARPS = 9;     // 8 analysis warps + 1 synthesis warp

__host__ __device__ inline size_t fbf_half_bytes(int S, int M, int B, int ov) {
  return ((size_t)S * M * ov + (size_t)S * B * ov) * sizeof(float);
}

template <int N>
__global__ void __launch_bounds__(FBF_WARPS * 32, 3) fixedbf_kernel(FixedBfArgs a, const float2 *__restrict__ tw_h_g,
                                                                const float2 *__restrict__ tw_n_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = N / 2, K = H + 1;
  constexpr int BE = fft_buf_elems(N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int ov = N - a.hop;
  const int R = N / a.hop;                       // hop divides N (checked on host)
  // smem carve-up
  float2 *micbuf = reinterpret_cast<float2 *>(smem_raw);           // [M][BE]
  float2 *beambuf = micbuf + (size_t)a.M * BE;                     // [B][BE]
  float *ola = reinterpret_cast<float *>(beambuf + (size_t)a.B * BE);   // [B][N] ring accumulator
  float *win = ola + (size_t)a.B * N;                              // [N]
  float2 *tw_h = reinterpret_cast<float2 *>(win + N);             // [H]   twiddles staged once per CTA
  float2 *tw_n = tw_h + H;                                         // [H/2 + 1]
  for (int i = tid; i < H; i += blockDim.x) tw_h[i] = tw_h_g[i];
  for (int i = tid; i <= H / 2; i += blockDim.x) tw_n[i] = tw_n_g[i];
  for (int n = tid; n < N; n += blockDim.x) win[n] = (float)a.window[n];
  for (int i = tid; i < a.B * N; i += blockDim.x) ola[i] = 0.0f;

  const int s = blockIdx.x / a.segs, seg = blockIdx.x % a.segs;
  const int t0 = seg * a.frames_per_seg;
  const int t1 = min(a.T, t0 + a.frames_per_seg);
  const int parity = *reinterpret_cast<const int *>(a.state);
  const size_t half = fbf_half_bytes(a.S, a.M, a.B, ov);
  const float *hist_in = reinterpret_cast<const float *>(a.state + 16 + (size_t)parity * half);
  const float *tail_in = hist_in + (size_t)a.S * a.M * ov;
  float *hist_out = reinterpret_cast<float *>(a.state + 16 + (size_t)(1 - parity) * half);
  float *tail_out = hist_out + (size_t)a.S * a.M * ov;
  __syncthreads();

  const float inv_n = 1.0f / (float)N;
  // frames before t0 that still overlap the segment's first output block are
  // recomputed (halo) so that each segment is self-contained.
  const int th0 = max(0, t0 - (R - 1));
  // Software pipeline over frames: in iteration t the analysis warps (0 .. FBF_WARPS-2) transform
  // frame t while the synthesis warp (FBF_WARPS-1) inverse-transforms and overlap-adds frame t-1;
  // the weight apply of frame t follows.  Two block barriers per frame, no idle phase.
  constexpr int AW = FBF_WARPS - 1;
  for (int t = th0; t <= t1; ++t) {
    if (warp < AW) {
      if (t < t1) {
        for (int m = warp; m < a.M; m += AW) {
          const float *xs = a.x + ((long long)s * a.M + m) * a.Ns;
          float2 *buf = micbuf + (size_t)m * BE;
          const int g0 = t * a.hop - ov;
          if (g0 >= 0 && ((reinterpret_cast<size_t>(xs + g0) & 7) == 0)) {
            typedef FftFirst<H, float> F1;
            float2 v[F1::PER][F1::R];
            const float2 *src = reinterpret_cast<const float2 *>(xs + g0);
            const float2 *w2 = reinterpret_cast<const float2 *>(win);
#pragma unroll
            for (int i = 0; i < F1::PER; ++i) {
              const int j = lane + 32 * i;
              if (F1::NB % 32 == 0 || j < F1::NB) {
#pragma unroll
                for (int r = 0; r < F1::R; ++r) {
                  const int e = j + r * F1::NB;
                  const float2 xv = __ldg(src + e);
                  const float2 wv = w2[e];
                  v[i][r] = make_float2(mul_rn(xv.x, wv.x), mul_rn(xv.y, wv.y));
                }
              }
            }
            F1::run(v, buf, tw_h, lane);
          } else {
            const float *hs = hist_in + ((long long)s * a.M + m) * ov;
            float *fb = reinterpret_cast<float *>(buf);
            for (int n = lane; n < N; n += 32) {
              const int g = g0 + n;
              const float v = (g < 0) ? hs[ov + g] : xs[g];
              fb[2 * FPAD<float>(n >> 1) + (n & 1)] = v * win[n];
            }
            __syncwarp();
            warp_cfft<H, float>(buf, tw_h, lane);
          }
        }
      }
    } else if (t > th0) {
      // ---- synthesis + overlap-add of frame tp = t-1 ----------------------------------
      const int tp = t - 1;
      for (int b = 0; b < a.B; ++b) {
        float2 *buf = beambuf + (size_t)b * BE;
        warp_irfft_unscaled<N, float>(buf, tw_h, tw_n, lane);
        const float *fb = reinterpret_cast<const float *>(buf);
        float *acc = ola + (size_t)b * N;
        for (int n = lane; n < N; n += 32) {               // ring: sample g lives at acc[g % N]
          const float v = fb[2 * FPAD<float>(n >> 1) + (n & 1)] * inv_n * win[n];
          const int pos = (tp * a.hop + n) & (N - 1);
          acc[pos] = acc[pos] + v;
        }
        __syncwarp();
        if (tp >= t0) {                                    // samples [tp*hop, (tp+1)*hop) are complete now
          float *ys = a.y + ((long long)s * a.B + b) * a.Ns;
          const float *tl = tail_in + ((long long)s * a.B + b) * ov;
          for (int j = lane; j < a.hop; j += 32) {
            const int g = tp * a.hop + j;
            const int pos = g & (N - 1);
            float v = acc[pos];
            if (g < ov) v = v + tl[g];
            ys[g] = (float)((double)v * a.scale);
            acc[pos] = 0.0f;
          }
        } else {
          for (int j = lane; j < a.hop; j += 32) acc[(tp * a.hop + j) & (N - 1)] = 0.0f;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- weights fused with the real-FFT split: micbuf holds the H-point complex spectra Z_m;
    //      X_m[k] = s/2 - e, X_m[H-k] = conj(s/2 + e) (see warp_rfft); Y_b[k] = sum_m conj(W[b,k,m]) X_m[k]
    if (t < t1) {
      for (int kk = tid; kk <= H / 2; kk += blockDim.x) {
        const float2 wn = tw_n[kk];
        for (int b = 0; b < a.B; ++b) {
          const float2 *w1 = a.W + ((size_t)b * K + kk) * a.M;
          const float2 *w2p = a.W + ((size_t)b * K + (H - kk)) * a.M;
          float y1r = 0.f, y1i = 0.f, y2r = 0.f, y2i = 0.f;
          for (int m = 0; m < a.M; ++m) {
            const float2 za = micbuf[(size_t)m * BE + FPAD<float>(kk)];
            const float2 zb = micbuf[(size_t)m * BE + FPAD<float>((H - kk) & (H - 1))];
            float x1r, x1i, x2r, x2i;
            if (kk == 0) {
              x1r = za.x + za.y; x1i = 0.f; x2r = za.x - za.y; x2i = 0.f;
            } else {
              const float sx = 0.5f * (za.x + zb.x), sy = 0.5f * (za.y - zb.y);
              const float dx = 0.5f * (za.x - zb.x), dy = 0.5f * (za.y + zb.y);
              const float px = wn.x * dx - wn.y * dy, py = wn.x * dy + wn.y * dx;
              x1r = sx + py; x1i = sy - px;          // s/2 - e, e = (-py, px)
              x2r = sx - py; x2i = -(sy + px);       // conj(s/2 + e)
            }
            const float2 wa = __ldg(w1 + m), wb = __ldg(w2p + m);
            y1r += wa.x * x1r + wa.y * x1i; y1i += wa.x * x1i - wa.y * x1r;
            y2r += wb.x * x2r + wb.y * x2i; y2i += wb.x * x2i - wb.y * x2r;
          }
          beambuf[(size_t)b * BE + FPAD<float>(kk)] = make_float2(y1r, y1i);
          beambuf[(size_t)b * BE + FPAD<float>(H - kk)] = make_float2(y2r, y2i);
        }
      }
    }
    __syncthreads();
  }
  // ---- new state (last segment of each stream) ----------------------------
  if (t1 == a.T && seg == a.segs - 1) {
    for (int i = tid; i < a.B * ov; i += blockDim.x) {
      int b = i / ov, j = i - b * ov;
      int g = a.T * a.hop + j;
      float v = ola[(size_t)b * N + (g & (N - 1))];
      if (g < ov) v = v + tail_in[((long long)s * a.B + b) * ov + g];
      tail_out[((long long)s * a.B + b) * ov + j] = v;
    }
    for (int i = tid; i < a.M * ov; i += blockDim.x) {
      int m = i / ov, j = i - m * ov;
      int g = a.Ns + j;  // index into concat(history, x)
      const float *hs = hist_in + ((long long)s * a.M + m) * ov;
      const float *xs = a.x + ((long long)s * a.M + m) * a.Ns;
      hist_out[((long long)s * a.M + m) * ov + j] = (g < ov) ? hs[g] : xs[g - ov];
    }
  }
}

// hop = N/2, up to 3 beams: one warp owns a run of consecutive frames of one stream and does everything
// for them -- M analysis FFTs whose real-FFT split feeds the weight sums held in registers (each lane owns
// the bin pairs (k, H-k), k = lane + 32 i), one inverse FFT per beam, overlap-add against the previous
// half-frame kept in shared memory.  No CTA-wide synchronisation, the spectrum exists only in registers.
// Same arithmetic as fixedbf_kernel.
constexpr int FSEQ_WARPS = 8;

template <int N, int NB>
__global__ void __launch_bounds__(FSEQ_WARPS * 32, NB == 1 ? 4 : 2) fixedbf_seq_kernel(FixedBfArgs a, const float2 *__restrict__ tw_h_g,
                                                                     const float2 *__restrict__ tw_n_g, int G, int nseg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = N / 2, K = H + 1, HOP = N / 2, BE = fft_buf_elems(N);
  constexpr int NPAIR = (H / 2) / 32 + 1;        // pair slots per lane: kk = lane + 32 i <= H/2
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  float *win = reinterpret_cast<float *>(smem_raw);                 // [N]
  float2 *tw_h = reinterpret_cast<float2 *>(win + N);               // [H]
  float2 *tw_n = tw_h + H;                                          // [H/2 + 1] (+1 pad)
  float2 *buf = tw_n + H / 2 + 2 + (size_t)warp * BE;
  float *prev = reinterpret_cast<float *>(tw_n + H / 2 + 2 + (size_t)FSEQ_WARPS * BE) + (size_t)warp * NB * HOP;
  float2 *Ws = reinterpret_cast<float2 *>(reinterpret_cast<float *>(tw_n + H / 2 + 2 + (size_t)FSEQ_WARPS * BE) +
                                          (size_t)FSEQ_WARPS * NB * HOP);       // [NB][M][K]: weights, bin index innermost
  for (int i = tid; i < H; i += blockDim.x) tw_h[i] = tw_h_g[i];
  for (int i = tid; i <= H / 2; i += blockDim.x) tw_n[i] = tw_n_g[i];
  for (int n = tid; n < N; n += blockDim.x) win[n] = (float)a.window[n];
  for (int i = tid; i < NB * K * a.M; i += blockDim.x) {              // a.W is [NB][K][M]
    const int m = i % a.M, k = (i / a.M) % K, b = i / (a.M * K);
    Ws[((size_t)b * a.M + m) * K + k] = a.W[i];
  }
  __syncthreads();
  const long long w = (long long)blockIdx.x * FSEQ_WARPS + warp;
  if (w >= (long long)a.S * nseg) return;
  const int s = (int)(w / nseg), seg = (int)(w % nseg);
  const int t0 = seg * G, t1 = min(a.T, t0 + G);
  const int parity = *reinterpret_cast<const int *>(a.state);
  const size_t half = fbf_half_bytes(a.S, a.M, a.B, HOP);
  const float *hist_in = reinterpret_cast<const float *>(a.state + 16 + (size_t)parity * half);
  const float *tail_in = hist_in + (size_t)a.S * a.M * HOP;
  float *hist_out = reinterpret_cast<float *>(a.state + 16 + (size_t)(1 - parity) * half);
  float *tail_out = hist_out + (size_t)a.S * a.M * HOP;
  const float inv_n = 1.0f / (float)N;
  float *fb = reinterpret_cast<float *>(buf);

  if (t0 == 0) {
    for (int i = lane; i < NB * HOP; i += 32) prev[i] = tail_in[(long long)s * NB * HOP + i];   // x[:overlap] += previous_output
    __syncwarp();
  }
  for (int t = (t0 > 0 ? t0 - 1 : 0); t < t1; ++t) {
    float2 y1[NB][NPAIR], y2[NB][NPAIR];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) { y1[b][i] = make_float2(0.f, 0.f); y2[b][i] = make_float2(0.f, 0.f); }
    const int g0 = t * HOP - HOP;
    for (int m = 0; m < a.M; ++m) {
      const float *xs = a.x + ((long long)s * a.M + m) * a.Ns;
      if (g0 >= 0 && ((reinterpret_cast<size_t>(xs + g0) & 7) == 0)) {
        typedef FftFirst<H, float> F1;
        float2 v[F1::PER][F1::R];
        const float2 *src = reinterpret_cast<const float2 *>(xs + g0);
        const float2 *w2 = reinterpret_cast<const float2 *>(win);
#pragma unroll
        for (int i = 0; i < F1::PER; ++i) {
          const int j = lane + 32 * i;
          if (F1::NB % 32 == 0 || j < F1::NB) {
#pragma unroll
            for (int r = 0; r < F1::R; ++r) {
              const int e = j + r * F1::NB;
              const float2 xv = __ldg(src + e);
              const float2 wv = w2[e];
              v[i][r] = make_float2(mul_rn(xv.x, wv.x), mul_rn(xv.y, wv.y));
            }
          }
        }
        F1::run(v, buf, tw_h, lane);
      } else {
        const float *hs = hist_in + ((long long)s * a.M + m) * HOP;
        for (int n = lane; n < N; n += 32) {
          const int g = g0 + n;
          const float v = (g < 0) ? hs[HOP + g] : xs[g];
          fb[2 * FPAD<float>(n >> 1) + (n & 1)] = v * win[n];
        }
        __syncwarp();
        warp_cfft<H, float>(buf, tw_h, lane);
      }
      // real-FFT split fused with the weight sums: Y_b[k] += conj(W[b,k,m]) X_m[k]
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) {
        const int kk = lane + 32 * i;
        if (kk <= H / 2) {
          const float2 za = buf[FPAD<float>(kk)];
          const float2 zb = buf[FPAD<float>((H - kk) & (H - 1))];
          float x1r, x1i, x2r, x2i;
          if (kk == 0) {
            x1r = za.x + za.y; x1i = 0.f; x2r = za.x - za.y; x2i = 0.f;
          } else {
            const float2 wn = tw_n[kk];
            const float sx = 0.5f * (za.x + zb.x), sy = 0.5f * (za.y - zb.y);
            const float dx = 0.5f * (za.x - zb.x), dy = 0.5f * (za.y + zb.y);
            const float px = wn.x * dx - wn.y * dy, py = wn.x * dy + wn.y * dx;
            x1r = sx + py; x1i = sy - px;          // s/2 - e, e = (-py, px)
            x2r = sx - py; x2i = -(sy + px);       // conj(s/2 + e)
          }
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const float2 wa = Ws[((size_t)b * a.M + m) * K + kk];
            const float2 wb = Ws[((size_t)b * a.M + m) * K + (H - kk)];
            y1[b][i].x += wa.x * x1r + wa.y * x1i; y1[b][i].y += wa.x * x1i - wa.y * x1r;
            y2[b][i].x += wb.x * x2r + wb.y * x2i; y2[b][i].y += wb.x * x2i - wb.y * x2r;
          }
        }
      }
      __syncwarp();
    }
    // synthesis of every beam, overlap-add with the previous half-frame
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
      for (int i = 0; i < NPAIR; ++i) {
        const int kk = lane + 32 * i;
        if (kk <= H / 2) { buf[FPAD<float>(kk)] = y1[b][i]; buf[FPAD<float>(H - kk)] = y2[b][i]; }
      }
      __syncwarp();
      warp_irfft_unscaled<N, float>(buf, tw_h, tw_n, lane);
      float *pv = prev + b * HOP;
      float *ys = a.y + ((long long)s * NB + b) * a.Ns;
      for (int j = lane; j < HOP; j += 32) {
        const float c0 = fb[2 * FPAD<float>(j >> 1) + (j & 1)] * inv_n * win[j];
        const float c1 = fb[2 * FPAD<float>((HOP + j) >> 1) + (j & 1)] * inv_n * win[HOP + j];
        if (t >= t0) {
          const float v = (t > 0) ? pv[j] + c0 : c0 + pv[j];      // (0 + f[t-1]) + f[t];  t = 0: f[0] + previous_output
          ys[(long long)t * HOP + j] = (float)((double)v * a.scale);
        }
        pv[j] = c1;
      }
      __syncwarp();
    }
  }
  // ---- new state (last run of each stream) ----------------------------------------
  if (t1 == a.T) {
    for (int i = lane; i < NB * HOP; i += 32) tail_out[(long long)s * NB * HOP + i] = prev[i];
    for (int i = lane; i < a.M * HOP; i += 32) {
      const int m = i / HOP, j = i - m * HOP;
      const int g = a.Ns + j;  // index into concat(history, x)
      const float *hs = hist_in + ((long long)s * a.M + m) * HOP;
      const float *xs = a.x + ((long long)s * a.M + m) * a.Ns;
      hist_out[((long long)s * a.M + m) * HOP + j] = (g < HOP) ? hs[g] : xs[g - HOP];
    }
  }
}

template <int N, int NB>
static int launch_fixedbf_seq(const FixedBfArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  long long nseg = (2LL * 148 * 24 + a.S - 1) / a.S;
  const long long max_seg = a.T / 16 > 0 ? a.T / 16 : 1;
  if (nseg > max_seg) nseg = max_seg;
  if (nseg < 1) nseg = 1;
  const int G = (int)((a.T + nseg - 1) / nseg);
  nseg = (a.T + G - 1) / G;
  const size_t smem = (size_t)N * sizeof(float) + (size_t)(N / 2 + N / 4 + 2) * sizeof(float2) +
                      (size_t)FSEQ_WARPS * (fft_buf_elems(N) * sizeof(float2) + (size_t)NB * (N / 2) * sizeof(float)) +
                      (size_t)NB * a.M * (N / 2 + 1) * sizeof(float2);
  if (smem > 227 * 1024) return DS_EUNSUPPORTED;
  auto kern = fixedbf_seq_kernel<N, NB>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long warps = (long long)a.S * nseg;
  kern<<<(unsigned)((warps + FSEQ_WARPS - 1) / FSEQ_WARPS), FSEQ_WARPS * 32, smem, st>>>(a, tw.h32, tw.n32, G, (int)nseg);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

// n_fft = 512, hop = 256, up to 3 beams: the fused fixed beamformer on the half-warp "square" transform (fft16.cuh).  A warp
// owns a run of consecutive frames of one stream and takes them two at a time (one per half-warp): per microphone one
// 256-point transform whose real-FFT split (partner bins from the mirror lane by shuffles) feeds the weight sums
// sum_m conj(W) X held in registers (lane j owns the bin pairs (k, 256 - k), k = j + 16 q < 128); per beam the inverse
// transform (merge, mirror shuffles, the same forward transform on conj(Z)), the synthesis window and the overlap-add:
// the first frame's second half reaches the second frame's half-warp by shuffles, the second frame's second half is
// carried in registers to the next pair.  One shared-memory round trip per transform, nothing else leaves the registers.
constexpr int FSQ_WARPS = 4;

#ifndef FSQ_MINB1
#define FSQ_MINB1 2         // with the prefetch: 254 registers, 8 warps per SM (A/B on the B200, config 2: no prefetch / 12 warps 3.41 ms,
                            // prefetch / 12 warps (spills) 3.53 ms, prefetch / 8 warps 2.86 ms; Stockham warp kernel 5.19 ms)
#endif
template <int NB>
__global__ void __launch_bounds__(FSQ_WARPS * 32, NB == 1 ? FSQ_MINB1 : 2) fixedbf_sq_kernel(FixedBfArgs a, const float2 *__restrict__ tw_h_g,
                                                                                    const float2 *__restrict__ tw_n_g, int G, int nseg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = 512, H = 256, K = H + 1, HOP = 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x, half = lane >> 4, j = lane & 15;
  float2 *s_win = reinterpret_cast<float2 *>(smem_raw);              // [H]  (w[2e], w[2e+1])
  float2 *s_tws = s_win + H;                                        // [H/2]
  float2 *xbuf = s_tws + H / 2;                                     // [FSQ_WARPS][2][16 * SQ_RS]
  float2 *Ws = xbuf + FSQ_WARPS * 2 * 16 * SQ_RS;                   // [NB][M][K]: weights, bin index innermost
  for (int i = tid; i < H; i += blockDim.x) s_win[i] = make_float2((float)a.window[2 * i], (float)a.window[2 * i + 1]);
  for (int i = tid; i < H / 2; i += blockDim.x) s_tws[i] = tw_n_g[i];
  for (int i = tid; i < NB * K * a.M; i += blockDim.x) {              // a.W is [NB][K][M]
    const int m = i % a.M, k = (i / a.M) % K, b = i / (a.M * K);
    Ws[((size_t)b * a.M + m) * K + k] = a.W[i];
  }
  float2 tw2[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) tw2[r] = tw_h_g[(r * j) & (H - 1)];
  __syncthreads();
  const long long w = (long long)blockIdx.x * FSQ_WARPS + warp;
  if (w >= (long long)a.S * nseg) return;                             // no CTA-wide synchronisation below
  float2 *xb = xbuf + (warp * 2 + half) * 16 * SQ_RS;
  const int s = (int)(w / nseg), seg = (int)(w % nseg);
  const int t0 = seg * G, t1 = min(a.T, t0 + G);
  const int parity = *reinterpret_cast<const int *>(a.state);
  const size_t halfb = fbf_half_bytes(a.S, a.M, a.B, HOP);
  const float *hist_in = reinterpret_cast<const float *>(a.state + 16 + (size_t)parity * halfb);
  const float *tail_in = hist_in + (size_t)a.S * a.M * HOP;
  float *hist_out = reinterpret_cast<float *>(a.state + 16 + (size_t)(1 - parity) * halfb);
  float *tail_out = hist_out + (size_t)a.S * a.M * HOP;
  const float inv_n = 1.0f / (float)N;
  const int mirror = (lane & 16) | ((16 - j) & 15);

  // pv: second half of the frame before the pair (samples 2 (j + 16 q), +1 of the block), needed by the first half-warp
  float2 pv[NB][8];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int q = 0; q < 8; ++q)
      pv[b][q] = (t0 == 0) ? *reinterpret_cast<const float2 *>(tail_in + ((long long)s * NB + b) * HOP + 2 * (j + 16 * q))   // x[:overlap] += previous_output
                           : make_float2(0.f, 0.f);

  for (int tt = (t0 > 0 ? t0 - 1 : 0); tt < t1; tt += 2) {
    const int t = tt + half;
    const bool valid = t < t1;
    float2 y1[NB][8], y2[NB][8], yq[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      yq[b] = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) { y1[b][q] = make_float2(0.f, 0.f); y2[b][q] = make_float2(0.f, 0.f); }
    }
    const int g0 = t * HOP - HOP;
#ifndef FSQ_PF
#define FSQ_PF 1          // the next microphone's samples are loaded before the current one is transformed
#endif
    // samples of microphone m for this lane (before the window): raw[r] = (x[2 e], x[2 e + 1]), e = j + 16 r
    auto load_mic = [&](int m, float2 (&raw)[16]) {
      const float *xs = a.x + ((long long)s * a.M + m) * a.Ns;
      if (valid && g0 >= 0 && ((reinterpret_cast<size_t>(xs + g0) & 7) == 0)) {
        const float2 *src = reinterpret_cast<const float2 *>(xs + g0) + j;
#pragma unroll
        for (int r = 0; r < 16; ++r) raw[r] = __ldg(src + 16 * r);
      } else if (valid) {
        const float *hs = hist_in + ((long long)s * a.M + m) * HOP;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int g = g0 + 2 * (j + 16 * r);
          raw[r] = make_float2((g < 0) ? hs[HOP + g] : xs[g], (g + 1 < 0) ? hs[HOP + g + 1] : xs[g + 1]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) raw[r] = make_float2(0.f, 0.f);
      }
    };
    float2 raw[16];
#if FSQ_PF
    load_mic(0, raw);
#endif
    for (int m = 0; m < a.M; ++m) {
      float2 v[16], u[16];
#if !FSQ_PF
      load_mic(m, raw);
#endif
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 wv = s_win[j + 16 * r];
        v[r] = make_float2(mul_rn(raw[r].x, wv.x), mul_rn(raw[r].y, wv.y));
      }
#if FSQ_PF
      if (m + 1 < a.M) load_mic(m + 1, raw);
#endif
      sq_cfft256(v, u, xb, tw2, j);
      // real-FFT split fused with the weight sums: Y_b[k] += conj(W[b,k,m]) X_m[k]
      const float2 *wm = Ws + (size_t)m * K;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 za = u[P16(q)];
        float2 zb;
        zb.x = __shfl_sync(0xffffffffu, u[P16(15 - q)].x, mirror);
        zb.y = __shfl_sync(0xffffffffu, u[P16(15 - q)].y, mirror);
        const float2 own = u[P16((16 - q) & 15)];
        zb.x = (j == 0) ? own.x : zb.x;
        zb.y = (j == 0) ? own.y : zb.y;
        float2 X1, X2;
        rfft_split_pair<float>(za, zb, s_tws[j + 16 * q], X1, X2);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float2 wa = wm[(size_t)b * a.M * K + j + 16 * q];
          const float2 wb = wm[(size_t)b * a.M * K + H - j - 16 * q];
          y1[b][q].x += wa.x * X1.x + wa.y * X1.y; y1[b][q].y += wa.x * X1.y - wa.y * X1.x;
          y2[b][q].x += wb.x * X2.x + wb.y * X2.y; y2[b][q].y += wb.x * X2.y - wb.y * X2.x;
        }
      }
      {                                                               // k = 128 pairs with itself (lane 0's register 8): X = conj(Z)
        const float2 za = u[P16(8)];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float2 wq = wm[(size_t)b * a.M * K + H / 2];
          yq[b].x += wq.x * za.x - wq.y * za.y; yq[b].y += -wq.x * za.y - wq.y * za.x;
        }
      }
    }
    // synthesis of every beam, overlap-add
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float2 z[16], Z2[8], r2[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) irfft_merge_pair<float>(y1[b][q], y2[b][q], s_tws[j + 16 * q], z[q], Z2[q]);
      float2 zq, zq2;
      irfft_merge_pair<float>(yq[b], yq[b], make_float2(0.f, -1.f), zq, zq2);
#pragma unroll
      for (int r = 8; r < 16; ++r) {
        float2 rc;
        rc.x = __shfl_sync(0xffffffffu, Z2[15 - r].x, mirror);
        rc.y = __shfl_sync(0xffffffffu, Z2[15 - r].y, mirror);
        const float2 own = (r == 8) ? zq : Z2[(16 - r) & 7];
        z[r].x = (j == 0) ? own.x : rc.x;
        z[r].y = (j == 0) ? own.y : rc.y;
      }
      if (j == 0) {                                                   // bins 0 and 256 (imaginary parts ignored like numpy.fft.irfft)
        const float a0 = y1[b][0].x, b0 = y2[b][0].x;
        z[0] = make_float2(a0 + b0, -(a0 - b0));
      }
      sq_cfft256(z, r2, xb, tw2, j);
      // r2[P16(q)] = conj(x[2n] + i x[2n+1]) * N, n = j + 16 q
      float2 c[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float2 wv = s_win[j + 16 * q];
        c[q] = make_float2(r2[P16(q)].x * inv_n * wv.x, -r2[P16(q)].y * inv_n * wv.y);
      }
      float *ys = a.y + ((long long)s * NB + b) * a.Ns + (long long)t * HOP;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float2 s0, nx;
        s0.x = __shfl_sync(0xffffffffu, c[q + 8].x, j);               // second half of the first frame of the pair
        s0.y = __shfl_sync(0xffffffffu, c[q + 8].y, j);
        nx.x = __shfl_sync(0xffffffffu, c[q + 8].x, j | 16);          // second half of the second frame
        nx.y = __shfl_sync(0xffffffffu, c[q + 8].y, j | 16);
        const float2 ad = half ? s0 : pv[b][q];
        const float vx = ad.x + c[q].x, vy = ad.y + c[q].y;           // (0 + f[t-1]) + f[t];  t = 0: previous_output + f[0]
        if (valid && t >= t0)
          *reinterpret_cast<float2 *>(ys + 2 * (j + 16 * q)) = make_float2((float)((double)vx * a.scale), (float)((double)vy * a.scale));
        pv[b][q] = (tt + 1 < t1) ? nx : s0;
      }
    }
  }
  // ---- new state (last run of each stream) ----------------------------------------
  if (t1 == a.T) {
    if (half == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float2 *>(tail_out + ((long long)s * NB + b) * HOP + 2 * (j + 16 * q)) = pv[b][q];
    }
    for (int i = lane; i < a.M * HOP; i += 32) {
      const int m = i / HOP, jj = i - m * HOP;
      const int g = a.Ns + jj;  // index into concat(history, x)
      const float *hs = hist_in + ((long long)s * a.M + m) * HOP;
      const float *xs = a.x + ((long long)s * a.M + m) * a.Ns;
      hist_out[((long long)s * a.M + m) * HOP + jj] = (g < HOP) ? hs[g] : xs[g - HOP];
    }
  }
}

template <int NB>
static int launch_fixedbf_sq(const FixedBfArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  long long nseg = (2LL * 148 * 24 + a.S - 1) / a.S;
  const long long max_seg = a.T / 16 > 0 ? a.T / 16 : 1;
  if (nseg > max_seg) nseg = max_seg;
  if (nseg < 1) nseg = 1;
  const int G = (int)((a.T + nseg - 1) / nseg);
  nseg = (a.T + G - 1) / G;
  const size_t smem = (size_t)(256 + 128 + FSQ_WARPS * 2 * 16 * SQ_RS) * sizeof(float2) + (size_t)NB * a.M * 257 * sizeof(float2);
  if (smem > 75 * 1024) return DS_EUNSUPPORTED;     // three (two) CTAs per SM
  auto kern = fixedbf_sq_kernel<NB>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long warps = (long long)a.S * nseg;
  kern<<<(unsigned)((warps + FSQ_WARPS - 1) / FSQ_WARPS), FSQ_WARPS * 32, smem, st>>>(a, tw.h32, tw.n32, G, (int)nseg);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

__global__ void flip_parity_kernel(unsigned char *state) {
  int *p = reinterpret_cast<int *>(state);
  *p = 1 - *p;
}

template <int N>
static int launch_fixedbf(const FixedBfArgs &a0, const TwiddleSet &tw, cudaStream_t st) {
  FixedBfArgs a = a0;
  if constexpr (N <= 512) {
    if (a.hop * 2 == N && a.B <= 3) {
      int rc = DS_EUNSUPPORTED;
#ifndef DS_FIXEDBF_NO_SQ
      if constexpr (N == 512) if ((reinterpret_cast<size_t>(a.y) & 7) == 0) rc = a.B == 1 ? launch_fixedbf_sq<1>(a, tw, st) : a.B == 2 ? launch_fixedbf_sq<2>(a, tw, st) : launch_fixedbf_sq<3>(a, tw, st);
#endif
      if (rc == DS_EUNSUPPORTED) rc = a.B == 1 ? launch_fixedbf_seq<N, 1>(a, tw, st) : a.B == 2 ? launch_fixedbf_seq<N, 2>(a, tw, st) : launch_fixedbf_seq<N, 3>(a, tw, st);
      if (rc == DS_OK) {
        flip_parity_kernel<<<1, 1, 0, st>>>(a.state);
        DS_LAUNCH_CHECK();
        return DS_OK;
      }
      if (rc != DS_EUNSUPPORTED) return rc;        // too many microphones for the weight tile: CTA-pipelined kernel below
    }
  }
  constexpr int BE = fft_buf_elems(N);
  const size_t smem = ((size_t)a.M + a.B) * BE * sizeof(float2) + (size_t)a.B * N * sizeof(float) + (size_t)N * sizeof(float) +
                      (size_t)(N / 2 + N / 4 + 2) * sizeof(float2);
  if (smem > 227 * 1024) { set_error("fixedbf: M=%d B=%d n_fft=%d does not fit shared memory", a.M, a.B, N); return DS_EUNSUPPORTED; }
  auto kern = fixedbf_kernel<N>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // split the frame axis into segments when there are too few streams to fill the GPU
  int segs = 1;
  const int target_ctas = 148 * 2;
  if (a.S < target_ctas) {
    segs = (target_ctas + a.S - 1) / a.S;
    int max_segs = a.T / 16;                     // keep the halo overhead small
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
  }
  a.frames_per_seg = (a.T + segs - 1) / segs;
  a.segs = (a.T + a.frames_per_seg - 1) / a.frames_per_seg;
  if (a.segs < 1) a.segs = 1;
  kern<<<a.S * a.segs, FBF_WARPS * 32, smem, st>>>(a, tw.h32, tw.n32);
  DS_LAUNCH_CHECK();
  flip_parity_kernel<<<1, 1, 0, st>>>(a.state);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_stft_num_frames(const ds_stft_params *p) {
  if (!p || p->n_fft <= 0 || p->hop <= 0 || p->n_samples < 0) return DS_EINVAL;
  const int ov = p->n_fft - p->hop;
  switch (p->mode) {
    case DS_STFT_STREAMING:
      if (ov + p->n_samples < p->n_fft) return 0;
      return 1 + (ov + p->n_samples - p->n_fft) / p->hop;
    case DS_STFT_CENTER:
      return 1 + p->n_samples / p->hop;
    case DS_STFT_PLAIN:
      if (p->n_samples < p->n_fft) return 0;
      return 1 + (p->n_samples - p->n_fft) / p->hop;
  }
  return DS_EINVAL;
}

static int stft_run_impl(const ds_stft_params *p, const double *window, float *history, const float *x, const short *x16,
                         void *X, void *stream) {
  DS_CHECK_ARG(p && window && (x || x16) && X, "ds_stft_run: null argument");
  DS_CHECK_ARG(p->hop >= 1 && p->hop <= p->n_fft, "ds_stft_run: hop %d out of range", p->hop);
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_ch >= 1 && p->n_samples >= 1, "ds_stft_run: bad shape");
  DS_CHECK_ARG(p->mode != DS_STFT_STREAMING || history || p->hop == p->n_fft, "ds_stft_run: streaming mode needs a history buffer");
  DS_CHECK_ARG(p->mode != DS_STFT_CENTER || p->n_samples > p->n_fft / 2, "ds_stft_run: reflect padding needs n_samples > n_fft/2");
  TwiddleSet tw;
  int rc = get_twiddles(p->n_fft, &tw);
  if (rc != DS_OK) return rc;
  const int T = ds_stft_num_frames(p);
  if (T < 0) { set_error("ds_stft_run: bad mode"); return DS_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  StftArgs a;
  a.x = x; a.x16 = x16; a.history = (p->mode == DS_STFT_STREAMING) ? history : nullptr; a.X = X; a.window = window; a.out_c128 = p->out_c128;
  a.S = p->n_streams; a.C = p->n_ch; a.Ns = p->n_samples; a.T = T; a.hop = p->hop; a.mode = p->mode;
  if (T > 0) {
    rc = p->fft_fp64 ? dispatch_stft<double>(p->n_fft, a, tw, st) : dispatch_stft<float>(p->n_fft, a, tw, st);
    if (rc != DS_OK) return rc;
  }
  const int ov = p->n_fft - p->hop;
  if (p->mode == DS_STFT_STREAMING && ov > 0) {
    int threads = (ov + 7) / 8;
    threads = ((threads + 31) / 32) * 32;
    if (x16) stft_history_kernel<true><<<p->n_streams * p->n_ch, threads, 0, st>>>(nullptr, x16, history, p->n_samples, ov);
    else stft_history_kernel<false><<<p->n_streams * p->n_ch, threads, 0, st>>>(x, nullptr, history, p->n_samples, ov);
    DS_LAUNCH_CHECK();
  }
  return DS_OK;
}

int ds_stft_run(const ds_stft_params *p, const double *window, float *history, const float *x, void *X, void *stream) {
  DS_CHECK_ARG(x, "ds_stft_run: null argument");
  return stft_run_impl(p, window, history, x, nullptr, X, stream);
}

int ds_stft_pcm16_run(const ds_stft_params *p, const double *window, float *history, const int16_t *x_pcm, void *X, void *stream) {
  DS_CHECK_ARG(x_pcm, "ds_stft_pcm16_run: null argument");
  return stft_run_impl(p, window, history, nullptr, (const short *)x_pcm, X, stream);
}

static int istft_run_impl(const ds_istft_params *p, const double *window, float *tail, const void *Y, float *y, short *y16,
                          void *stream, long long frame_pitch = 0) {
  DS_CHECK_ARG(p && window && Y && (y || y16), "ds_istft_run: null argument");
  DS_CHECK_ARG(p->hop >= 1 && p->hop <= p->n_fft, "ds_istft_run: hop %d out of range", p->hop);
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_ch >= 1 && p->n_frames >= 1, "ds_istft_run: bad shape");
  DS_CHECK_ARG(p->mode == DS_STFT_STREAMING || p->mode == DS_STFT_PLAIN, "ds_istft_run: mode must be STREAMING or PLAIN");
  DS_CHECK_ARG(p->mode != DS_STFT_STREAMING || tail || p->hop == p->n_fft, "ds_istft_run: streaming mode needs a tail buffer");
  TwiddleSet tw;
  int rc = get_twiddles(p->n_fft, &tw);
  if (rc != DS_OK) return rc;
  IstftArgs a;
  a.Y = Y; a.in_c128 = p->in_c128; a.y = y; a.y16 = y16; a.tail = (p->mode == DS_STFT_STREAMING) ? tail : nullptr; a.window = window;
  a.S = p->n_streams; a.C = p->n_ch; a.T = p->n_frames; a.hop = p->hop; a.mode = p->mode;
  a.n_out = (p->mode == DS_STFT_STREAMING) ? p->n_frames * p->hop : p->n_fft + p->hop * (p->n_frames - 1);
  a.scale = p->scale;
  const long long K = p->n_fft / 2 + 1;
  if (frame_pitch > 0) {       // [S][C][K][frame_pitch]: frames innermost (the tensor-core multi-beam kernel's output)
    a.sK = frame_pitch; a.sT = 1; a.sC = K * frame_pitch; a.sS = (long long)a.C * K * frame_pitch;
  } else {                     // [S][T][C][K]
    a.sK = 1; a.sC = K; a.sT = (long long)a.C * K; a.sS = (long long)a.T * a.C * K;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return p->fft_fp64 ? dispatch_istft<double>(p->n_fft, a, tw, st) : dispatch_istft<float>(p->n_fft, a, tw, st);
}

int ds_istft_frames_inner_run(const ds_istft_params *p, const double *window, float *tail, const void *Y, long long frame_pitch,
                              float *y, void *stream) {
  DS_CHECK_ARG(y && p && frame_pitch >= p->n_frames, "ds_istft_frames_inner_run: bad argument");
  return istft_run_impl(p, window, tail, Y, y, nullptr, stream, frame_pitch);
}

int ds_istft_run(const ds_istft_params *p, const double *window, float *tail, const void *Y, float *y, void *stream) {
  DS_CHECK_ARG(y, "ds_istft_run: null argument");
  return istft_run_impl(p, window, tail, Y, y, nullptr, stream);
}

int ds_istft_pcm16_run(const ds_istft_params *p, const double *window, float *tail, const void *Y, int16_t *y_pcm, void *stream) {
  DS_CHECK_ARG(y_pcm, "ds_istft_pcm16_run: null argument");
  return istft_run_impl(p, window, tail, Y, nullptr, (short *)y_pcm, stream);
}

size_t ds_fixedbf_state_bytes(const ds_fixedbf_params *p) {
  if (!p) return 0;
  const int ov = p->n_fft - p->hop;
  return 16 + 2 * fbf_half_bytes(p->n_streams, p->n_mics, p->n_beams, ov);
}

int ds_fixedbf_run(const ds_fixedbf_params *p, const double *window, const void *W, void *state, const float *x,
                   float *y, void *stream) {
  DS_CHECK_ARG(p && window && W && state && x && y, "ds_fixedbf_run: null argument");
  DS_CHECK_ARG(p->hop * 2 == p->n_fft || p->hop * 4 == p->n_fft, "ds_fixedbf_run: hop must be n_fft/2 or n_fft/4");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_mics >= 1 && p->n_beams >= 1, "ds_fixedbf_run: bad shape");
  DS_CHECK_ARG(p->n_samples >= p->hop && p->n_samples % p->hop == 0, "ds_fixedbf_run: n_samples must be a positive multiple of hop");
  TwiddleSet tw;
  int rc = get_twiddles(p->n_fft, &tw);
  if (rc != DS_OK) return rc;
  FixedBfArgs a;
  a.x = x; a.y = y; a.W = (const float2 *)W; a.window = window; a.state = (unsigned char *)state;
  a.S = p->n_streams; a.M = p->n_mics; a.B = p->n_beams; a.Ns = p->n_samples; a.T = p->n_samples / p->hop;
  a.hop = p->hop; a.segs = 1; a.frames_per_seg = a.T; a.scale = p->scale;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->n_fft) {
    case 128: return launch_fixedbf<128>(a, tw, st);
    case 256: return launch_fixedbf<256>(a, tw, st);
    case 512: return launch_fixedbf<512>(a, tw, st);
    case 1024: return launch_fixedbf<1024>(a, tw, st);
    case 2048: return launch_fixedbf<2048>(a, tw, st);
  }
  set_error("ds_fixedbf_run: unsupported n_fft %d", p->n_fft);
  return DS_EUNSUPPORTED;
}

}  // extern "C"
