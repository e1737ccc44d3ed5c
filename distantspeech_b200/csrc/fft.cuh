// fft.cuh -- warp-cooperative real FFTs in shared memory.
//
// One warp transforms one frame.  A real frame of N samples is viewed as
// H = N/2 complex points z[n] = x[2n] + i x[2n+1] (which is simply the frame as
// it lies in memory), transformed by an in-place Stockham autosort FFT whose
// radix-8/4/2 butterflies live in registers, and then split into the N/2+1
// bins of the real transform.  Between passes the lanes exchange data through a
// padded shared-memory buffer (FPAD) so that both the strided and the
// contiguous side of every pass stay (almost) bank-conflict free.
//
// Twiddles come from exactly rounded tables (computed on the host in long
// double): tw_h[i] = exp(-2 pi i/H), tw_n[k] = exp(-2 pi k/N).
#pragma once
#include "common.cuh"

namespace ds {

// element index -> swizzled / padded element index.  Chosen by exhaustive search over
// every warp-level access of the transforms below (Stockham passes, real-FFT split,
// frame load, bin read-out):
//   8-byte elements (float2):  XOR swizzle of the 16-element group index  -> 1.10 x the
//                              conflict-free wavefront count (pad-per-8 was 1.78 x)
//   16-byte elements (double2): one pad element every 8                    -> 1.10 x
template <typename T> __host__ __device__ __forceinline__ constexpr int FPAD(int i) {
  return sizeof(T) == 4 ? (i ^ ((i >> 4) & 7) ^ (((i >> 6) & 1) << 3)) : (i + (i >> 3));
}
// number of V2 elements a frame buffer needs (H+1 bins, swizzled within 16-element groups / padded)
__host__ __device__ constexpr int fft_buf_elems(int n_fft) { return (n_fft / 2 + 1) + (n_fft / 2 + 1) / 8 + 17; }

// product rounded on its own (never contracted into an FMA with the first butterfly), so a frame gives
// the same bits whether its windowed samples went through shared memory or straight into registers
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
#ifndef DS_FFT_NO_F32X2
// sm_100a packed fp32: one FADD2 per complex add / subtract instead of two FADDs.  The transforms are bound by issue slots
// (STFT: issue 77 %): -16 % floating-point instructions in the STFT kernel, measured 22.26 -> 22.05 ms per config-4 step and
// 133.7 -> 130.0 ms for the FDGSC pipeline (A/B on the B200, profiles/ab_runs_r02.txt); same rounding as the scalar adds.
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return *reinterpret_cast<float2 *>(&r);
}
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return *reinterpret_cast<float2 *>(&r);
}
#endif
// multiply by -i : (x + iy)(-i) = y - ix
template <typename C> __device__ __forceinline__ C cmul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }

template <typename C> __device__ __forceinline__ void dft2(C *v) {
  C a = v[0], b = v[1];
  v[0] = cadd(a, b); v[1] = csub(a, b);
}
template <typename C> __device__ __forceinline__ void dft4(C &x0, C &x1, C &x2, C &x3) {
  C t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmul_mi(csub(x1, x3));
  x0 = cadd(t0, t2); x2 = csub(t0, t2); x1 = cadd(t1, t3); x3 = csub(t1, t3);
}
template <typename T, typename C> __device__ __forceinline__ void dft8(C *v) {
  // even / odd 4-point DFTs, then combine with W8^k
  C e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  C o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  const T h = (T)0.70710678118654752440;
  C w1; w1.x = (o1.x + o1.y) * h; w1.y = (o1.y - o1.x) * h;   // o1 * (1-i)/sqrt2
  C w2 = cmul_mi(o2);                                          // o2 * (-i)
  C w3; w3.x = (o3.y - o3.x) * h; w3.y = -(o3.x + o3.y) * h;  // o3 * (-1-i)/sqrt2
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
  v[2] = cadd(e2, w2); v[6] = csub(e2, w2);
  v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
}

template <int H, int NS, typename T> struct FftPass {
  typedef typename V2<T>::type C;
  static constexpr int REM = H / NS;
  static constexpr int R = (REM % 8 == 0) ? 8 : ((REM % 4 == 0) ? 4 : 2);
  static constexpr int NB = H / R;                  // butterflies in this pass
  static constexpr int PER = (NB + 31) / 32;        // per lane

  __device__ __forceinline__ static void run(C *buf, const C *__restrict__ tw, int lane) {
    C v[PER][R];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      if (NB % 32 == 0 || j < NB) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[i][r] = buf[FPAD<T>(j + r * NB)];
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      if (NB % 32 == 0 || j < NB) {
        const int k = j & (NS - 1);
        if (NS > 1) {
          constexpr int TSTEP = H / (NS * R);
#ifndef DS_FFT_TW_TABLE
          // fp32: one table load per butterfly, the other twiddles as powers of it.  The transforms are bound by
          // shared-memory wavefronts (l1tex 92 %), the FMA pipe has room: -6 % on the STFT kernel (A/B on the B200),
          // spectra still within 3e-6 of the frame peak of numpy's (chain SNR 113 dB).  fp64 keeps the exact table.
          if constexpr (sizeof(T) == 4 && R >= 4) {
            C w[R];
            w[1] = tw[k * TSTEP];
            w[2] = cmul(w[1], w[1]);
            w[3] = cmul(w[2], w[1]);
            if constexpr (R == 8) { w[4] = cmul(w[2], w[2]); w[5] = cmul(w[4], w[1]); w[6] = cmul(w[3], w[3]); w[7] = cmul(w[4], w[3]); }
#pragma unroll
            for (int r = 1; r < R; ++r) v[i][r] = cmul(v[i][r], w[r]);
          } else
#endif
          {
#pragma unroll
            for (int r = 1; r < R; ++r) v[i][r] = cmul(v[i][r], tw[r * k * TSTEP]);
          }
        }
        if (R == 8) dft8<T>(v[i]);
        else if (R == 4) dft4(v[i][0], v[i][1], v[i][2], v[i][3]);
        else dft2(v[i]);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) buf[FPAD<T>(j0 + r * NS)] = v[i][r];
      }
    }
    __syncwarp();
    if constexpr (NS * R < H) FftPass<H, NS * R, T>::run(buf, tw, lane);
  }
};

// First Stockham pass fed from registers: v[i][r] = z[(lane + 32 i) + r * NB].  Saves one full
// shared-memory round trip per transform when the caller can load its frame straight from
// global memory in this (coalesced) order.  Continues with the remaining passes in `buf`.
template <int H, typename T> struct FftFirst {
  typedef typename V2<T>::type C;
  static constexpr int R = FftPass<H, 1, T>::R;
  static constexpr int NB = H / R;
  static constexpr int PER = (NB + 31) / 32;
  __device__ __forceinline__ static void run(C (&v)[PER][R], C *buf, const C *__restrict__ tw, int lane) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      if (NB % 32 == 0 || j < NB) {
        if (R == 8) dft8<T>(v[i]);
        else if (R == 4) dft4(v[i][0], v[i][1], v[i][2], v[i][3]);
        else dft2(v[i]);
#pragma unroll
        for (int r = 0; r < R; ++r) buf[FPAD<T>(j * R + r)] = v[i][r];
      }
    }
    __syncwarp();
    if constexpr (R < H) FftPass<H, R, T>::run(buf, tw, lane);
  }
};

// Real-FFT split of the H-point complex spectrum in `buf`, written straight to `out[0..H]`
// (global memory) instead of back to shared memory.  OutC is float2 or double2.
template <int N, typename T, typename OutC>
__device__ __forceinline__ void warp_rfft_split_store(const typename V2<T>::type *buf, const typename V2<T>::type *__restrict__ tw_n,
                                                      OutC *__restrict__ out, int lane) {
  typedef typename V2<T>::type C;
  constexpr int H = N / 2;
  // fp32: tw_n[lane + 32 i] = tw_n[lane] * tw_n[32]^i -- two table loads per frame instead of one per bin pair
  // (same trade as the Stockham twiddles above: -0.08 ms on the STFT); fp64 keeps the exact table
  C wrun = tw_n[lane <= H / 2 ? lane : 0];
  const C wstep = tw_n[H / 2 >= 32 ? 32 : 0];
  for (int k = lane; k <= H / 2; k += 32) {
    if (k == 0) {
      if (sizeof(T) == 4) wrun = cmul(wrun, wstep);
      const C a = buf[FPAD<T>(0)];
      OutC o0, oh;
      o0.x = a.x + a.y; o0.y = 0;
      oh.x = a.x - a.y; oh.y = 0;
      out[0] = o0; out[H] = oh;
    } else {
      const C a = buf[FPAD<T>(k)], b = buf[FPAD<T>(H - k)];
      C w;
      if (sizeof(T) == 4) { w = wrun; wrun = cmul(wrun, wstep); } else w = tw_n[k];
      const T sx = (T)0.5 * (a.x + b.x), sy = (T)0.5 * (a.y - b.y);
      const T dx = (T)0.5 * (a.x - b.x), dy = (T)0.5 * (a.y + b.y);
      const T px = w.x * dx - w.y * dy, py = w.x * dy + w.y * dx;
      const T ex = -py, ey = px;
      OutC o1, o2;
      o1.x = sx - ex; o1.y = sy - ey;
      o2.x = sx + ex; o2.y = -(sy + ey);
      out[k] = o1; out[H - k] = o2;
    }
  }
}

// The real-FFT split and the inverse's merge for one bin pair (k, H - k), 1 <= k <= H/2, w = exp(-2 pi i k / N) -- the
// bodies of warp_rfft / warp_irfft_unscaled below, for kernels that fuse them with their own elementwise work instead of
// paying a shared-memory round trip for each (fdgsc2.cu).  Same operations, same rounding.
//   split: a = Z[k], b = Z[H-k] of the H-point transform  ->  X1 = X[k], X2 = X[H-k]
template <typename T, typename C> __device__ __forceinline__ void rfft_split_pair(C a, C b, C w, C &X1, C &X2) {
  const T sx = (T)0.5 * (a.x + b.x), sy = (T)0.5 * (a.y - b.y);
  const T dx = (T)0.5 * (a.x - b.x), dy = (T)0.5 * (a.y + b.y);
  const T px = w.x * dx - w.y * dy, py = w.x * dy + w.y * dx;
  const T ex = -py, ey = px;
  X1 = mk2<T>(sx - ex, sy - ey);
  X2 = mk2<T>(sx + ex, -(sy + ey));
}
//   merge: a = Y[k], b = Y[H-k]  ->  Z1 = conj(Z[k]), Z2 = conj(Z[H-k]) (input of the forward transform that inverts)
template <typename T, typename C> __device__ __forceinline__ void irfft_merge_pair(C a, C b, C w, C &Z1, C &Z2) {
  const T sx = a.x + b.x, sy = a.y - b.y;
  const T dx = a.x - b.x, dy = a.y + b.y;
  const T px = w.x * dx + w.y * dy, py = w.x * dy - w.y * dx;
  const T fx = -py, fy = px;
  Z1 = mk2<T>(sx + fx, -(sy + fy));
  Z2 = mk2<T>(sx - fx, (sy - fy));
}

// in-place forward complex FFT of H points held in buf[FPAD<T>(i)], one warp
template <int H, typename T>
__device__ __forceinline__ void warp_cfft(typename V2<T>::type *buf, const typename V2<T>::type *__restrict__ tw_h, int lane) {
  FftPass<H, 1, T>::run(buf, tw_h, lane);
}

// Forward real FFT of N = 2H samples.  On entry buf[FPAD<T>(n)] = (x[2n], x[2n+1]);
// on exit buf[FPAD<T>(k)] = X[k], k = 0..H.   (numpy.fft.rfft)
template <int N, typename T>
__device__ __forceinline__ void warp_rfft(typename V2<T>::type *buf, const typename V2<T>::type *__restrict__ tw_h,
                                          const typename V2<T>::type *__restrict__ tw_n, int lane) {
  typedef typename V2<T>::type C;
  constexpr int H = N / 2;
  warp_cfft<H, T>(buf, tw_h, lane);
  // split: pair (k, H-k).  s = a + conj(b), d = a - conj(b), e = (i/2) W^k d
  // X[k] = s/2 - e ; X[H-k] = conj(s/2 + e)
  for (int k = lane; k <= H / 2; k += 32) {
    if (k == 0) {
      C a = buf[FPAD<T>(0)];
      buf[FPAD<T>(0)] = mk2<T>(a.x + a.y, (T)0);
      buf[FPAD<T>(H)] = mk2<T>(a.x - a.y, (T)0);
    } else {
      C a = buf[FPAD<T>(k)], b = buf[FPAD<T>(H - k)];
      C w = tw_n[k];
      T sx = (T)0.5 * (a.x + b.x), sy = (T)0.5 * (a.y - b.y);
      T dx = (T)0.5 * (a.x - b.x), dy = (T)0.5 * (a.y + b.y);
      // e = i * (w * d)
      T px = w.x * dx - w.y * dy, py = w.x * dy + w.y * dx;
      T ex = -py, ey = px;
      buf[FPAD<T>(k)] = mk2<T>(sx - ex, sy - ey);
      buf[FPAD<T>(H - k)] = mk2<T>(sx + ex, -(sy + ey));
    }
  }
  __syncwarp();
}

// Inverse real FFT.  On entry buf[FPAD<T>(k)] = Y[k], k = 0..H (imaginary parts of
// Y[0], Y[H] are ignored like numpy.fft.irfft); on exit buf[FPAD<T>(n)] =
// (x[2n], x[2n+1]) * N, i.e. UNSCALED -- the caller folds 1/N into its window.
template <int N, typename T>
__device__ __forceinline__ void warp_irfft_unscaled(typename V2<T>::type *buf, const typename V2<T>::type *__restrict__ tw_h,
                                                    const typename V2<T>::type *__restrict__ tw_n, int lane) {
  typedef typename V2<T>::type C;
  constexpr int H = N / 2;
  // merge: Z[k] = s + f, Z[H-k] = conj(s - f), s = a + conj(b), d = a - conj(b),
  // f = i conj(W^k) d   (everything x2 relative to numpy; total scale N)
  // we store conj(Z) so that a forward FFT followed by a conjugate is the inverse.
#ifdef DS_FFT_MERGE_POW
  // next-round A/B (default off): merge twiddles by the same recurrence as warp_rfft_split_store
  C wrun = tw_n[lane <= H / 2 ? lane : 0];
  const C wstep = tw_n[H / 2 >= 32 ? 32 : 0];
#endif
  for (int k = lane; k <= H / 2; k += 32) {
    if (k == 0) {
#ifdef DS_FFT_MERGE_POW
      if (sizeof(T) == 4) wrun = cmul(wrun, wstep);
#endif
      T a = buf[FPAD<T>(0)].x, b = buf[FPAD<T>(H)].x;
      buf[FPAD<T>(0)] = mk2<T>(a + b, -(a - b));
    } else {
      C a = buf[FPAD<T>(k)], b = buf[FPAD<T>(H - k)];
#ifdef DS_FFT_MERGE_POW
      C w;
      if (sizeof(T) == 4) { w = wrun; wrun = cmul(wrun, wstep); } else w = tw_n[k];
#else
      C w = tw_n[k];
#endif
      T sx = a.x + b.x, sy = a.y - b.y;
      T dx = a.x - b.x, dy = a.y + b.y;
      // conj(w) * d
      T px = w.x * dx + w.y * dy, py = w.x * dy - w.y * dx;
      T fx = -py, fy = px;
      buf[FPAD<T>(k)] = mk2<T>(sx + fx, -(sy + fy));          // conj(Z[k])
      buf[FPAD<T>(H - k)] = mk2<T>(sx - fx, (sy - fy));       // conj(Z[H-k]) = s - f
    }
  }
  __syncwarp();
  warp_cfft<H, T>(buf, tw_h, lane);
  // result r[n] = conj(z[n]) * (2H): x[2n] = r.x, x[2n+1] = -r.y
  for (int n = lane; n < H; n += 32) {
    C r = buf[FPAD<T>(n)];
    r.y = -r.y;
    buf[FPAD<T>(n)] = r;
  }
  __syncwarp();
}

}  // namespace ds
