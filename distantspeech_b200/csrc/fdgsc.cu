// fdgsc.cu -- overlap-save frequency-domain GSC, FDGSC.process(postfilter=False)
// (beamformer/FDGSC.py:201-317): DC notch -> time-alignment FIR -> mean fixed
// beamformer -> M coefficient-constrained FDAF blocking filters -> M-channel
// norm-constrained FDAF interference canceller.
//
// Reference behaviour restated (file:line relative to the reference tree):
//   FilterDcNotch16.filter_dc_notch16     adaptivefilter/feature.py:37-49   (in place on x, FDGSC.py:211-213)
//   TimeAlignment.process / fir_filter    beamformer/fixedbeamformer.py:13-93
//   FDGSC.fixed_beamformer                FDGSC.py:123-138  (mean over mics)
//   Transform.stft + MCRA (L = 60)        FDGSC.py:239-253  (mic 0; p[:32] raised to 0.8 when mean(p[32:128]) > 0.8)
//   DelaySamples.delay                    beamformer/utils.py:241-274
//   FastFreqLms.compute_freq_conv/xcorr   adaptivefilter/FastFreqLms.py:138-200
//   AdaptiveBlockingMatrixFilter.update   beamformer/gsc_bm.py:61-122 (p = 1, tap clamp :92-111)
//   AdaptiveInterferenceCancellation.update  beamformer/gsc_aic.py:54-108 (p = 1 - mean(p), norm cap 0.003)
//
// One CTA owns one stream for all blocks; every length-512 FFT is done by one warp
// in shared memory; all filter state (2 M x 257 complex weights, power spectra,
// delay lines, FIR cache, MCRA) stays in shared memory for the whole utterance, so
// HBM traffic is the input once and the output once.
#include "common.cuh"
#include "fft.cuh"
#include "perbin.cuh"

namespace ds {

constexpr int FD_L = 256;            // frameLen (block length, filter length)
constexpr int FD_N = 512;            // FFT length
constexpr int FD_K = 257;            // bins
constexpr int FD_FLMAX = 128;        // max time-alignment FIR length
constexpr int FD_WARPS = 8;
constexpr int FD_NT = FD_WARPS * 32;

struct FdgscArgs {
  double *state;              // [S][elems] (see fdgsc_state_elems)
  const double *h;            // [M][FL] time-alignment filters
  const float *x;             // [S][M][N] (already DC-notched)
  float *y;                   // [S][N]
  float *bm_out;              // [S][M][N] or null
  float *fix_out;             // [S][N] or null
  double *p_out;              // [S][nblk][K] or null
  const double *window;       // [512] sqrt-hann (Transform of FDGSC.py:104)
  int S, M, Ns, FL, frm_cnt, ell;
  double mu_bm, mu_aic, alpha, maxnorm, delta;
  McraConst mc;
};

__host__ __device__ inline size_t fdgsc_state_elems(int M) {
  return (size_t)2 * M * FD_K * 2      // Wbm, Waic (complex)
         + 2 * FD_K                    // Pf, Pa
         + FD_L                        // fbf_prev
         + (size_t)M * FD_L            // bm_prev
         + FD_L                        // x0_prev
         + (size_t)M * (FD_FLMAX - 1)  // fir cache
         + (size_t)M * (FD_L / 2)      // delay line aligned
         + FD_L                        // delay line fbf
         + 5 * FD_K                    // mcra
         + 2 * (size_t)M;              // notch memories (used by the notch kernel)
}

// ---------------------------------------------------------------------------
// DC notch, in place, one thread per (stream, mic), bit-exact operation order
// ---------------------------------------------------------------------------
__global__ void dcnotch_kernel(float *x, double *state, int S, int M, int Ns, double radius, double den2) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= S * M) return;
  const int s = g / M, m = g % M;
  float *xs = x + (size_t)g * Ns;
  double *mem = state + (size_t)s * fdgsc_state_elems(M) + (fdgsc_state_elems(M) - 2 * (size_t)M) + 2 * m;
  double m0 = mem[0], m1 = mem[1];
  const bool vec = ((reinterpret_cast<size_t>(xs) & 15) == 0) && (Ns % 4 == 0);
  if (vec) {
    float4 *x4 = reinterpret_cast<float4 *>(xs);
    for (int i = 0; i < Ns / 4; ++i) {
      float4 v = x4[i];
      float *pv = reinterpret_cast<float *>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double vin = (double)pv[j];
        const double vout = __dadd_rn(m0, vin);
        m0 = __dadd_rn(m1, __dmul_rn(2.0, __dadd_rn(-vin, __dmul_rn(radius, vout))));
        m1 = __dsub_rn(vin, __dmul_rn(den2, vout));
        pv[j] = (float)__dmul_rn(radius, vout);
      }
      x4[i] = v;
    }
  } else {
    for (int i = 0; i < Ns; ++i) {
      const double vin = (double)xs[i];
      const double vout = __dadd_rn(m0, vin);
      m0 = __dadd_rn(m1, __dmul_rn(2.0, __dadd_rn(-vin, __dmul_rn(radius, vout))));
      m1 = __dsub_rn(vin, __dmul_rn(den2, vout));
      xs[i] = (float)__dmul_rn(radius, vout);
    }
  }
  mem[0] = m0; mem[1] = m1;
}

// The same filter, parallel along time.  One warp per (stream, mic) row walks it in tiles of 32 x 64 samples staged
// through shared memory (coalesced float4 loads / stores; the sequential kernel above reads 32 different rows per load
// instruction and spent 106 ms on config 3's 4096 x 6 x 160 000 samples).  The notch is the linear system
//   s' = A s + b v,  vout = s_0 + v,  A = [[2 r, 1], [-den2, 0]],  y = r vout,
// so each lane first runs its 64-sample segment from a zero state to get the segment's end state z_l, the lanes' true
// start states follow by s_l = A^64 s_{l-1} + z_{l-1}, and a second pass runs the recurrence from s_l in the reference's
// operation order.  s_l equals the sequentially computed state up to the last bits of a double, and the filter
// contracts (|eig A| = 0.98), so the float32 outputs are those of the sequential kernel (tests allow 1 ulp).
constexpr int NOTCH_SEG = 64, NOTCH_WARPS = 4;
__global__ void __launch_bounds__(NOTCH_WARPS * 32) dcnotch_scan_kernel(float *x, double *state, int S, int M, int Ns, double radius, double den2) {
  __shared__ float tile[NOTCH_WARPS][32 * (NOTCH_SEG + 1)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * NOTCH_WARPS + warp;
  if (row >= (long long)S * M) return;
  const int s = (int)(row / M), m = (int)(row % M);
  float *xs = x + (size_t)row * Ns;
  double *mem = state + (size_t)s * fdgsc_state_elems(M) + (fdgsc_state_elems(M) - 2 * (size_t)M) + 2 * m;
  float *tl = tile[warp];
  // T = A^SEG
  double t00 = 1.0, t01 = 0.0, t10 = 0.0, t11 = 1.0;
  for (int n = 0; n < NOTCH_SEG; ++n) {          // T <- A T
    const double n00 = fma(2.0 * radius, t00, t10), n01 = fma(2.0 * radius, t01, t11);
    const double n10 = -den2 * t00, n11 = -den2 * t01;
    t00 = n00; t01 = n01; t10 = n10; t11 = n11;
  }
  double c0 = mem[0], c1 = mem[1];                // state at the start of the tile (warp-uniform)
  constexpr int TILE = 32 * NOTCH_SEG;
  for (int base = 0; base < Ns; base += TILE) {
    const int nt = min(TILE, Ns - base);          // a multiple of NOTCH_SEG (host guarantees Ns % 64 == 0)
    for (int e = lane * 4; e < nt; e += 128) {
      const float4 v = *reinterpret_cast<const float4 *>(xs + base + e);
      float *d = tl + (e / NOTCH_SEG) * (NOTCH_SEG + 1) + (e % NOTCH_SEG);
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncwarp();
    const bool active = lane * NOTCH_SEG < nt;
    float *seg = tl + lane * (NOTCH_SEG + 1);
    double m0 = 0.0, m1 = 0.0;
    if (active) {
#pragma unroll 8
      for (int n = 0; n < NOTCH_SEG; ++n) {       // pass 1: end state of the segment's zero-state response
        const double vin = (double)seg[n];
        const double vout = m0 + vin;
        m0 = m1 + 2.0 * (radius * vout - vin);
        m1 = vin - den2 * vout;
      }
    }
    // start state of every lane: serial combine over the 32 segments of the tile
    double s0 = 0.0, s1 = 0.0;
    for (int l = 0; l < 32; ++l) {
      if (l * NOTCH_SEG >= nt) break;
      if (lane == l) { s0 = c0; s1 = c1; }
      const double z0 = __shfl_sync(0xffffffffu, m0, l), z1 = __shfl_sync(0xffffffffu, m1, l);
      const double n0 = fma(t00, c0, fma(t01, c1, z0)), n1 = fma(t10, c0, fma(t11, c1, z1));
      c0 = n0; c1 = n1;
    }
    // second pass from the true start state (sequential recurrence again: exact, no cancellation between a
    // zero-state and a homogeneous part)
    if (active) {
      m0 = s0; m1 = s1;
#pragma unroll 8
      for (int n = 0; n < NOTCH_SEG; ++n) {
        const double vin = (double)seg[n];
        const double vout = __dadd_rn(m0, vin);
        m0 = __dadd_rn(m1, __dmul_rn(2.0, __dadd_rn(-vin, __dmul_rn(radius, vout))));
        m1 = __dsub_rn(vin, __dmul_rn(den2, vout));
        seg[n] = (float)__dmul_rn(radius, vout);
      }
    }
    __syncwarp();
    for (int e = lane * 4; e < nt; e += 128) {
      const float *d = tl + (e / NOTCH_SEG) * (NOTCH_SEG + 1) + (e % NOTCH_SEG);
      *reinterpret_cast<float4 *>(xs + base + e) = make_float4(d[0], d[1], d[2], d[3]);
    }
    __syncwarp();
  }
  if (lane == 0) { mem[0] = c0; mem[1] = c1; }
}

int fdgsc_notch_launch(float *x, double *state, int S, int M, int Ns, double r, cudaStream_t st) {
  const double den2 = r * r + 0.7 * (1 - r) * (1 - r);
  const long long rows = (long long)S * M;
  if (Ns % NOTCH_SEG == 0 && ((reinterpret_cast<size_t>(x) & 15) == 0)) {
    dcnotch_scan_kernel<<<(unsigned)((rows + NOTCH_WARPS - 1) / NOTCH_WARPS), NOTCH_WARPS * 32, 0, st>>>(x, state, S, M, Ns, r, den2);
  } else {
    dcnotch_kernel<<<(unsigned)((rows + 63) / 64), 64, 0, st>>>(x, state, S, M, Ns, r, den2);
  }
  DS_LAUNCH_CHECK();
  return DS_OK;
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
template <typename T> struct FdSmem {
  typedef typename V2<T>::type C2;
  // element counts (in units of T unless noted)
  static __host__ __device__ size_t bytes(int M) {
    size_t c2 = (size_t)FD_N / 2 + (FD_N / 4 + 2)          // twiddles
                + (size_t)FD_WARPS * fft_buf_elems(FD_N)   // fft buffers
                + (size_t)2 * M * FD_K                      // Wbm, Waic
                + (size_t)M * FD_K                          // Xa
                + 2 * FD_K;                                 // Xf, E
    size_t t = (size_t)2 * FD_K                             // Pf, Pa
               + FD_N                                       // window
               + (size_t)M * FD_FLMAX                       // h
               + (size_t)M * (FD_FLMAX - 1 + FD_L)          // inp (cache + block)
               + (size_t)2 * M * FD_L                       // xa (aliased by bm_cur), xad
               + (size_t)M * FD_L                           // bm_prev
               + (size_t)M * (FD_L / 2)                     // dl_al
               + 5 * FD_L;                                  // fbf, fbf_prev, fbf_d, dl_fbf, x0_prev
    size_t d = (size_t)FD_K + 64;                           // P0 (K) + reduction scratch (doubles); MCRA state lives in registers
    return c2 * sizeof(C2) + t * sizeof(T) + d * sizeof(double) + 64;
  }
};

__device__ __forceinline__ double block_sum(double v, double *scratch) {
  // scratch: >= 32 doubles
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < FD_WARPS; ++w) r += scratch[w];
  return r;
}

template <typename T>
__global__ void __launch_bounds__(FD_NT, 2) fdgsc_kernel(FdgscArgs a, const typename V2<T>::type *__restrict__ tw_h_g,
                                                      const typename V2<T>::type *__restrict__ tw_n_g) {
  typedef typename V2<T>::type C2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int L = FD_L, N = FD_N, K = FD_K, H = N / 2, BE = fft_buf_elems(FD_N);
  const int M = a.M, FL = a.FL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x;

  // ---- carve shared memory -------------------------------------------------
  C2 *tw_h = reinterpret_cast<C2 *>(smem_raw);
  C2 *tw_n = tw_h + H;
  C2 *fftb = tw_n + (H / 2 + 2);
  C2 *Wbm = fftb + (size_t)FD_WARPS * BE;
  C2 *Waic = Wbm + (size_t)M * K;
  C2 *Xa = Waic + (size_t)M * K;
  C2 *Xf = Xa + (size_t)M * K;
  C2 *Ef = Xf + K;
  T *Pf = reinterpret_cast<T *>(Ef + K);
  T *Pa = Pf + K;
  T *win = Pa + K;
  T *hh = win + N;                               // [M][FD_FLMAX]
  T *inp = hh + (size_t)M * FD_FLMAX;            // [M][FLMAX-1+L]
  T *xa = inp + (size_t)M * (FD_FLMAX - 1 + L);  // [M][L]  (dead after the delay stage: aliased by bmc)
  T *xad = xa + (size_t)M * L;
  T *bmc = xa;
  T *bmp = xad + (size_t)M * L;
  T *dl_al = bmp + (size_t)M * L;                // [M][L/2]
  T *fbf = dl_al + (size_t)M * (L / 2);
  T *fbf_prev = fbf + L;
  T *fbf_d = fbf_prev + L;
  T *dl_fbf = fbf_d + L;
  T *x0_prev = dl_fbf + L;
  double *P0 = reinterpret_cast<double *>((reinterpret_cast<size_t>(x0_prev + L) + 15) & ~(size_t)15);     // [K]
  double *red = P0 + K;                          // 64 doubles scratch
  const int IS = FD_FLMAX - 1 + L;               // inp row stride

  // ---- load state ------------------------------------------------------------
  double *st = a.state + (size_t)s * fdgsc_state_elems(M);
  size_t o = 0;
  for (int i = tid; i < M * K; i += FD_NT) { Wbm[i] = mk2<T>((T)st[o + 2 * i], (T)st[o + 2 * i + 1]); }
  o += (size_t)2 * M * K;
  for (int i = tid; i < M * K; i += FD_NT) { Waic[i] = mk2<T>((T)st[o + 2 * i], (T)st[o + 2 * i + 1]); }
  o += (size_t)2 * M * K;
  for (int i = tid; i < K; i += FD_NT) { Pf[i] = (T)st[o + i]; Pa[i] = (T)st[o + K + i]; }
  o += 2 * K;
  for (int i = tid; i < L; i += FD_NT) fbf_prev[i] = (T)st[o + i];
  o += L;
  for (int i = tid; i < M * L; i += FD_NT) bmp[i] = (T)st[o + i];
  o += (size_t)M * L;
  for (int i = tid; i < L; i += FD_NT) x0_prev[i] = (T)st[o + i];
  o += L;
  for (int i = tid; i < M * (FD_FLMAX - 1); i += FD_NT) inp[(i / (FD_FLMAX - 1)) * IS + (i % (FD_FLMAX - 1))] = (T)st[o + i];
  o += (size_t)M * (FD_FLMAX - 1);
  for (int i = tid; i < M * (L / 2); i += FD_NT) dl_al[i] = (T)st[o + i];
  o += (size_t)M * (L / 2);
  for (int i = tid; i < L; i += FD_NT) dl_fbf[i] = (T)st[o + i];
  o += L;
  // MCRA state of bin tid (and of bin 256 for thread 0) stays in registers for the whole utterance
  double mst[2][5];
  const size_t mcra_off = o;
#pragma unroll
  for (int e = 0; e < 5; ++e) { mst[0][e] = st[o + (size_t)e * K + tid]; mst[1][e] = (tid == 0) ? st[o + (size_t)e * K + (K - 1)] : 0.0; }
  for (int i = tid; i < H; i += FD_NT) tw_h[i] = tw_h_g[i];
  for (int i = tid; i <= H / 2; i += FD_NT) tw_n[i] = tw_n_g[i];
  for (int i = tid; i < N; i += FD_NT) win[i] = (T)a.window[i];
  for (int i = tid; i < M * FD_FLMAX; i += FD_NT) {
    const int m = i / FD_FLMAX, k = i % FD_FLMAX;
    hh[i] = (k < FL) ? (T)a.h[m * FL + k] : (T)0;
  }
  __syncthreads();

  C2 *buf = fftb + (size_t)warp * BE;
  T *fb = reinterpret_cast<T *>(buf);
  const T invN = (T)1 / (T)N;
  const int nblk = a.Ns / L;
  int frm = a.frm_cnt, ell = a.ell % a.mc.L;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
#define FIDX(n) (2 * FPAD<T>((n) >> 1) + ((n) & 1))

  for (int blk = 0; blk < nblk; ++blk) {
    // ---- S1: load the (notched) block behind the FIR cache ---------------------
    for (int i = tid; i < M * L; i += FD_NT) {
      const int m = i / L, n = i % L;
      inp[m * IS + (FD_FLMAX - 1) + n] = (T)a.x[((size_t)s * M + m) * a.Ns + (size_t)blk * L + n];
    }
    __syncthreads();
    // ---- S2: time alignment FIR (fixedbeamformer.py:13-48) + mean beamformer --------
    {
      const int n = tid;                       // FD_NT == L
      T acc_mean = (T)0;
      for (int m = 0; m < M; ++m) {
        const T *row = inp + m * IS + (FD_FLMAX - 1) + n;
        const T *hm = hh + m * FD_FLMAX;
        T acc = (T)0;
        for (int k = 0; k < FL; ++k) acc += hm[k] * row[-k];
        xa[m * L + n] = acc;
        acc_mean += acc;
      }
      fbf[n] = acc_mean / (T)M;                // np.mean(x, axis=1)  (FDGSC.py:138)
    }
    __syncthreads();
    // cache <- last FLMAX-1 samples ; delay lines (utils.py:241-274)
    for (int i = tid; i < M * (FD_FLMAX - 1); i += FD_NT) {
      const int m = i / (FD_FLMAX - 1), k = i % (FD_FLMAX - 1);
      inp[m * IS + k] = inp[m * IS + L + k];    // FLMAX-1 <= L, so source and destination never overlap
    }
    for (int i = tid; i < M * L; i += FD_NT) {
      const int m = i / L, n = i % L;
      xad[i] = (n < L / 2) ? dl_al[m * (L / 2) + n] : xa[m * L + n - L / 2];
    }
    {
      const int n = tid;
      fbf_d[n] = dl_fbf[n];
    }
    __syncthreads();
    for (int i = tid; i < M * (L / 2); i += FD_NT) {
      const int m = i / (L / 2), j = i % (L / 2);
      dl_al[i] = xa[m * L + L / 2 + j];
    }
    dl_fbf[tid] = fbf[tid];

    // ---- S4: spectra of the BM reference (fbf) and of raw mic 0 (for the MCRA) -------
    if (warp == 0) {
      for (int n = lane; n < N; n += 32) fb[FIDX(n)] = (n < L) ? fbf_prev[n] : fbf[n - L];
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) Xf[k] = buf[FPAD<T>(k)];
    } else if (warp == 1) {
      const T *x0 = inp + (FD_FLMAX - 1);       // mic 0 block (still intact: only the cache part was rewritten)
      for (int n = lane; n < N; n += 32) fb[FIDX(n)] = ((n < L) ? x0_prev[n] : x0[n - L]) * win[n];
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) {
        const C2 v = buf[FPAD<T>(k)];
        const double re = (double)(float)v.x, im = (double)(float)v.y;    // complex64 rounding (transform.py:212)
        P0[k] = re * re + im * im;
      }
    }
    __syncthreads();
    for (int n = tid; n < L; n += FD_NT) { fbf_prev[n] = fbf[n]; x0_prev[n] = inp[(FD_FLMAX - 1) + n]; }

    // ---- S5: BM input power, MCRA on mic 0, adaptation-control heuristics ------------
    double p_loc[2] = {0.0, 0.0};
    {
      const bool reset = (frm > 0) && (ell == 0);
      int q = 0;
      for (int k = tid; k < K; k += FD_NT, ++q) {
        const C2 v = Xf[k];
        T pf = (T)a.alpha * Pf[k] + ((T)1 - (T)a.alpha) * (v.x * v.x + v.y * v.y);      // FastFreqLms.py:158
        Pf[k] = (pf < (T)1e-4) ? (T)1e-4 : pf;                                          // :189
        const double Ym1 = (k > 0) ? P0[k - 1] : 0.0, Yp1 = (k < K - 1) ? P0[k + 1] : 0.0;
        if (q == 0) mcra_step(mst[0][0], mst[0][1], mst[0][2], mst[0][3], mst[0][4], Ym1, P0[k], Yp1, k, K, frm, reset, a.mc);
        else mcra_step(mst[1][0], mst[1][1], mst[1][2], mst[1][3], mst[1][4], Ym1, P0[k], Yp1, k, K, frm, reset, a.mc);
        p_loc[q] = (q == 0) ? mst[0][3] : mst[1][3];
      }
      if (reset) ell = 0;
      ++ell; ++frm;
      if (ell == a.mc.L) ell = 0;
    }
    // mean(p[32:128]) > 0.8  ->  p[:32] = max(p[:32], 0.8)            (FDGSC.py:247-249)
    const double mid = block_sum((tid >= 32 && tid < 128) ? p_loc[0] : 0.0, red) / 96.0;
    if (mid > 0.8 && tid < 32 && p_loc[0] < 0.8) p_loc[0] = 0.8;
    const double pbar = block_sum(p_loc[0] + ((tid == 0) ? p_loc[1] : 0.0), red) / (double)K;
    if (a.p_out) {
      double *po = a.p_out + ((size_t)s * nblk + blk) * K;
      po[tid] = p_loc[0];
      if (tid == 0) po[K - 1] = p_loc[1];
    }
    const T step_aic = (T)((1.0 - pbar) * a.mu_aic);                   // p * mu   (gsc_aic.py:82, FDGSC.py:279)
    const T step_bm = (T)(1.0 * a.mu_bm);                              // p = 1.0  (gsc_bm.py:90, FDGSC.py:260)
    __syncthreads();

    // ---- S6: M blocking filters, one warp each (gsc_bm.py:61-122) -------------------------
    for (int m = warp; m < M; m += FD_WARPS) {
      C2 *W = Wbm + (size_t)m * K;
      for (int k = lane; k < K; k += 32) buf[FPAD<T>(k)] = cmul(Xf[k], W[k]);
      __syncwarp();
      warp_irfft_unscaled<N, T>(buf, tw_h, tw_n, lane);
      for (int n = lane; n < L; n += 32) {
        const T yv = fb[FIDX(L + n)] * invN;                            // last hop_len samples (:161)
        bmc[m * L + n] = xad[m * L + n] - yv;                           // e = d - y (:174)
      }
      __syncwarp();
      for (int n = lane; n < N; n += 32) fb[FIDX(n)] = (n < L) ? (T)0 : bmc[m * L + n - L];     // e_pad (:185)
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) {
        const C2 g = cmulc(buf[FPAD<T>(k)], Xf[k]);                     // conj(X) * E
        const T ip = (T)1 / Pf[k];
        C2 w = W[k];
        w.x += step_bm * (g.x * ip); w.y += step_bm * (g.y * ip);
        buf[FPAD<T>(k)] = w;
      }
      __syncwarp();
      warp_irfft_unscaled<N, T>(buf, tw_h, tw_n, lane);
      for (int n = lane; n < N; n += 32) {
        T v = fb[FIDX(n)] * invN;
        if (n >= L) {
          v = (T)0;                                                     // w[-hop_len:] = 0 (:94)
        } else {
          T ub = (T)a.delta;
          const int d = n - N / 4;
          if (d == 0) ub = (T)0.9; else if (d == 1 || d == -1) ub = (T)0.3; else if (d == 2 || d == -2) ub = (T)0.05;
          v = fmin(fmax(v, -(T)a.delta), ub);                           // tap bounds (:48-59, :96-108)
        }
        fb[FIDX(n)] = v;
      }
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) W[k] = buf[FPAD<T>(k)];
      __syncwarp();
    }
    __syncthreads();
    if (a.bm_out) {
      for (int i = tid; i < M * L; i += FD_NT) {
        const int m = i / L, n = i % L;
        a.bm_out[((size_t)s * M + m) * a.Ns + (size_t)blk * L + n] = (float)bmc[i];
      }
    }
    if (a.fix_out) a.fix_out[(size_t)s * a.Ns + (size_t)blk * L + tid] = (float)fbf[tid];

    // ---- S7: interference canceller (gsc_aic.py:54-108) ------------------------------------
    for (int m = warp; m < M; m += FD_WARPS) {
      for (int n = lane; n < N; n += 32) fb[FIDX(n)] = (n < L) ? bmp[m * L + n] : bmc[m * L + n - L];
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) Xa[(size_t)m * K + k] = buf[FPAD<T>(k)];
      __syncwarp();
    }
    __syncthreads();
    for (int i = tid; i < M * L; i += FD_NT) bmp[i] = bmc[i];
    for (int k = tid; k < K; k += FD_NT) {
      T pw = (T)0;
      C2 acc = mk2<T>((T)0, (T)0);
      for (int m = 0; m < M; ++m) {
        const C2 v = Xa[(size_t)m * K + k];
        pw += v.x * v.x + v.y * v.y;
        const C2 pr = cmul(v, Waic[(size_t)m * K + k]);
        acc.x += pr.x; acc.y += pr.y;
      }
      T pa = (T)a.alpha * Pa[k] + ((T)1 - (T)a.alpha) * pw;
      Pa[k] = (pa < (T)1e-4) ? (T)1e-4 : pa;
      Ef[k] = acc;                                                      // sum_ch X W  (:161)
    }
    __syncthreads();
    if (warp == 0) {
      for (int k = lane; k < K; k += 32) buf[FPAD<T>(k)] = Ef[k];
      __syncwarp();
      warp_irfft_unscaled<N, T>(buf, tw_h, tw_n, lane);
      T ev[L / 32];
#pragma unroll
      for (int i = 0; i < L / 32; ++i) {
        const int n = lane + 32 * i;
        ev[i] = fbf_d[n] - fb[FIDX(L + n)] * invN;                      // e = d - y
        a.y[(size_t)s * a.Ns + (size_t)blk * L + n] = (float)ev[i];
      }
      __syncwarp();
      for (int n = lane; n < L; n += 32) fb[FIDX(n)] = (T)0;
#pragma unroll
      for (int i = 0; i < L / 32; ++i) fb[FIDX(L + lane + 32 * i)] = ev[i];
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) Ef[k] = buf[FPAD<T>(k)];
    }
    __syncthreads();
    double nrm = 0.0;
    for (int i = tid; i < M * K; i += FD_NT) {
      const int k = i % K;
      const C2 g = cmulc(Ef[k], Xa[i]);                                 // conj(X) * E
      const T ip = (T)1 / Pa[k];
      C2 w = Waic[i];
      w.x += step_aic * (g.x * ip); w.y += step_aic * (g.y * ip);
      Waic[i] = w;
      nrm += (double)w.x * (double)w.x + (double)w.y * (double)w.y;
    }
    nrm = block_sum(nrm, red) / (double)N / (double)N;                  // :86
    const T sc = (nrm > a.maxnorm) ? (T)sqrt(a.maxnorm / nrm) : (T)1;
    for (int m = warp; m < M; m += FD_WARPS) {
      C2 *W = Waic + (size_t)m * K;
      for (int k = lane; k < K; k += 32) buf[FPAD<T>(k)] = W[k];
      __syncwarp();
      warp_irfft_unscaled<N, T>(buf, tw_h, tw_n, lane);
      for (int n = lane; n < N; n += 32) fb[FIDX(n)] = (n >= L) ? (T)0 : fb[FIDX(n)] * invN * sc;   // :95-96
      __syncwarp();
      warp_rfft<N, T>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) W[k] = buf[FPAD<T>(k)];
      __syncwarp();
    }
    __syncthreads();
  }
#undef FIDX

  // ---- save state ----------------------------------------------------------------
  o = 0;
  for (int i = tid; i < M * K; i += FD_NT) { st[o + 2 * i] = (double)Wbm[i].x; st[o + 2 * i + 1] = (double)Wbm[i].y; }
  o += (size_t)2 * M * K;
  for (int i = tid; i < M * K; i += FD_NT) { st[o + 2 * i] = (double)Waic[i].x; st[o + 2 * i + 1] = (double)Waic[i].y; }
  o += (size_t)2 * M * K;
  for (int i = tid; i < K; i += FD_NT) { st[o + i] = (double)Pf[i]; st[o + K + i] = (double)Pa[i]; }
  o += 2 * K;
  for (int i = tid; i < L; i += FD_NT) st[o + i] = (double)fbf_prev[i];
  o += L;
  for (int i = tid; i < M * L; i += FD_NT) st[o + i] = (double)bmp[i];
  o += (size_t)M * L;
  for (int i = tid; i < L; i += FD_NT) st[o + i] = (double)x0_prev[i];
  o += L;
  for (int i = tid; i < M * (FD_FLMAX - 1); i += FD_NT) st[o + i] = (double)inp[(i / (FD_FLMAX - 1)) * IS + (i % (FD_FLMAX - 1))];
  o += (size_t)M * (FD_FLMAX - 1);
  for (int i = tid; i < M * (L / 2); i += FD_NT) st[o + i] = (double)dl_al[i];
  o += (size_t)M * (L / 2);
  for (int i = tid; i < L; i += FD_NT) st[o + i] = (double)dl_fbf[i];
  o += L;
#pragma unroll
  for (int e = 0; e < 5; ++e) { st[mcra_off + (size_t)e * K + tid] = mst[0][e]; if (tid == 0) st[mcra_off + (size_t)e * K + (K - 1)] = mst[1][e]; }
}

template <typename T>
static int launch_fdgsc(const FdgscArgs &a, const TwiddleSet &tw, cudaStream_t st) {
  const size_t smem = FdSmem<T>::bytes(a.M);
  if (smem > 227 * 1024) { set_error("fdgsc: %d mics need %zu bytes of shared memory", a.M, smem); return DS_EUNSUPPORTED; }
  auto kern = fdgsc_kernel<T>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.S, FD_NT, smem, st>>>(a, TwSel<T>::h(tw), TwSel<T>::n(tw));
  DS_LAUNCH_CHECK();
  return DS_OK;
}

// streaming per-channel FIR (fir_filter, fixedbeamformer.py:13-48): y[n] = sum_k h[k] x[n-k]
__global__ void fir_kernel(const double *__restrict__ h, double *cache, const double *__restrict__ x, double *__restrict__ y,
                           int S, int M, int Ns, int FL, int pass) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int sm = (int)(g / Ns), n = (int)(g % Ns);
  if (sm >= S * M) return;
  const int m = sm % M;
  const double *xs = x + (size_t)sm * Ns;
  double *cs = cache + (size_t)sm * (FL - 1);
  if (pass == 0) {
    double acc = 0.0;
    for (int k = 0; k < FL; ++k) {
      const int i = n - k;
      acc += h[m * FL + k] * ((i >= 0) ? xs[i] : cs[FL - 1 + i]);
    }
    y[(size_t)sm * Ns + n] = acc;
  } else if (n < FL - 1) {        // cache <- last FL-1 samples of concat(cache, x), computed into y scratch rows first
    const int i = Ns - (FL - 1) + n;
    y[(size_t)sm * Ns + n] = (i >= 0) ? xs[i] : cs[FL - 1 + i];
  }
}
__global__ void fir_cache_commit_kernel(double *cache, const double *scratch, int SM, int Ns, int FL) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= SM * (FL - 1)) return;
  cache[g] = scratch[(size_t)(g / (FL - 1)) * Ns + (g % (FL - 1))];
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_fir_run(int n_streams, int n_ch, int n_samples, int filter_len, const double *h, double *cache, const double *x,
               double *y, double *scratch, void *stream) {
  DS_CHECK_ARG(h && cache && x && y && scratch, "ds_fir_run: null argument");
  DS_CHECK_ARG(n_streams >= 1 && n_ch >= 1 && n_samples >= 1 && filter_len >= 2 && n_samples >= filter_len - 1,
               "ds_fir_run: need n_samples >= filter_len - 1 >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n_streams * n_ch * n_samples;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  fir_kernel<<<blocks, 256, 0, st>>>(h, cache, x, y, n_streams, n_ch, n_samples, filter_len, 0);
  DS_LAUNCH_CHECK();
  fir_kernel<<<blocks, 256, 0, st>>>(h, cache, x, scratch, n_streams, n_ch, n_samples, filter_len, 1);
  DS_LAUNCH_CHECK();
  const int cn = n_streams * n_ch * (filter_len - 1);
  fir_cache_commit_kernel<<<(cn + 255) / 256, 256, 0, st>>>(cache, scratch, n_streams * n_ch, n_samples, filter_len);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

void ds_fdgsc_default_params(ds_fdgsc_params *p, int n_streams, int n_mics, int n_samples, int filter_len) {
  if (!p) return;
  p->frame_len = FD_L; p->n_streams = n_streams; p->n_mics = n_mics; p->n_samples = n_samples; p->filter_len = filter_len;
  p->frm_cnt = 0; p->ell = 1; p->mcra_L = 60; p->fp64 = 0; p->dc_notch = 1; p->reserved = 0; p->reserved2 = 0;
  p->mu_bm = 0.1; p->mu_aic = 0.1; p->alpha = 0.9; p->notch_radius = 0.98; p->maxnorm = 0.003; p->delta = 0.001;
  p->mcra_alpha_d = 0.95; p->mcra_alpha_s = 0.8; p->mcra_delta_s = 5.0; p->mcra_alpha_p = 0.2;
  p->mcra_p_min = 1e-3; p->mcra_p_max = 0.999;
}

size_t ds_fdgsc_state_bytes(const ds_fdgsc_params *p) {
  if (!p) return 0;
  return (size_t)p->n_streams * fdgsc_state_elems(p->n_mics) * sizeof(double);
}

int ds_fdgsc_run(const ds_fdgsc_params *p, const double *delay_filter, const double *window, void *state, float *x,
                 float *y, float *bm_out, float *fix_out, double *p_out, void *stream) {
  DS_CHECK_ARG(p && delay_filter && window && state && x && y, "ds_fdgsc_run: null argument");
  DS_CHECK_ARG(p->frame_len == FD_L, "ds_fdgsc_run: only frameLen = 256 is compiled");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_mics >= 2 && p->n_mics <= 8, "ds_fdgsc_run: n_mics must be 2..8");
  DS_CHECK_ARG(p->n_samples >= FD_L && p->n_samples % FD_L == 0, "ds_fdgsc_run: n_samples must be a positive multiple of 256");
  DS_CHECK_ARG(p->filter_len >= 1 && p->filter_len <= FD_FLMAX, "ds_fdgsc_run: alignment filter longer than %d taps", FD_FLMAX);
  TwiddleSet tw;
  int rc = get_twiddles(FD_N, &tw);
  if (rc != DS_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dc_notch) {
    rc = fdgsc_notch_launch(x, (double *)state, p->n_streams, p->n_mics, p->n_samples, p->notch_radius, st);
    if (rc != DS_OK) return rc;
  }
  FdgscArgs a;
  a.state = (double *)state; a.h = delay_filter; a.x = x; a.y = y; a.bm_out = bm_out; a.fix_out = fix_out; a.p_out = p_out;
  a.window = window;
  a.S = p->n_streams; a.M = p->n_mics; a.Ns = p->n_samples; a.FL = p->filter_len; a.frm_cnt = p->frm_cnt; a.ell = p->ell;
  a.mu_bm = p->mu_bm; a.mu_aic = p->mu_aic; a.alpha = p->alpha; a.maxnorm = p->maxnorm; a.delta = p->delta;
  a.mc.alpha_d = p->mcra_alpha_d; a.mc.alpha_s = p->mcra_alpha_s; a.mc.delta_s = p->mcra_delta_s;
  a.mc.alpha_p = p->mcra_alpha_p; a.mc.p_min = p->mcra_p_min; a.mc.p_max = p->mcra_p_max; a.mc.L = p->mcra_L;
  return p->fp64 ? launch_fdgsc<double>(a, tw, st) : launch_fdgsc<float>(a, tw, st);
}

int ds_fdgsc_notch_run(const ds_fdgsc_params *p, void *state, float *x, int n_samples, void *stream) {
  DS_CHECK_ARG(p && state && x && n_samples >= 1, "ds_fdgsc_notch_run: bad argument");
  return fdgsc_notch_launch(x, (double *)state, p->n_streams, p->n_mics, n_samples, p->notch_radius, (cudaStream_t)stream);
}

}  // extern "C"
