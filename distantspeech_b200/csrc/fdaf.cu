// fdaf.cu -- constrained frequency-domain adaptive filter, the base FastFreqLms.update
// (adaptivefilter/FastFreqLms.py:203-245) that TDGSC uses as interference canceller (beamformer/TDGSC.py:96-108).
// frame length 256 (n_fft 512), C input channels, overlap-save:
//   input buffer [old | new] per channel, X = rfft(buffer)                              :137-146
//   P = alpha P + (1 - alpha) sum_c |X_c|^2                                             :147
//   y = irfft(sum_c X_c W_c)[-256:];  d delayed by 128 samples when non_causal          :148, :156-157
//   e = d - y;  E = rfft([0 ... 0 | e]);  P[P < 1e-4] = 1e-4;  grad = conj(X) E / P     :160, :189-193
//   gradient constraint: irfft, zero the last 256 samples, rfft                         :196-200
//   W += p 2 mu grad  (p: per-bin gate or 1)                                            :231
//   fir_truncate: w = irfft(W)[:256], zero the first / last `fir_truncate` taps, W = rfft(w, 512)   :238-243
// One CTA per stream, one warp per input channel, blocks sequential; three CTA barriers per block.  Filter state and
// spectra live in shared memory for the whole call (fp32 arithmetic; the reference is float64 -- the parity tests
// hold it to the same 1e-4 / 60 dB bar as the other pipelines).
#include "common.cuh"
#include "fft.cuh"

namespace ds {

namespace fdaf {
constexpr int N = 512, H = 256, K = 257, L = 256, BE = fft_buf_elems(N);
constexpr int MAXC = 7;

struct Args {
  float *state;             // [S][state_elems(C)] float32
  const float *x;           // [S][C][Ns]
  const float *d;           // [S][Ns]
  const double *p;          // [S][T][K] or null
  float *e;                 // [S][Ns]
  const float2 *tw_h, *tw_n;
  int S, C, Ns, fir_truncate, non_causal, one_minus_p;
  float mu, alpha;
};
// state per stream: W [C][K] float2, P [K], input old halves [C][L], delay tail [L/2]
__host__ __device__ inline size_t state_elems(int C) { return (size_t)C * K * 2 + K + (size_t)C * L + L / 2; }

__global__ void __launch_bounds__(32 * MAXC) fdaf_kernel(Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = a.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x, nthr = blockDim.x;
  float2 *tw_h = reinterpret_cast<float2 *>(smem_raw);              // [H]
  float2 *tw_n = tw_h + H;                                          // [H/2 + 1] (+1 pad)
  float2 *bufs = tw_n + H / 2 + 2;                                  // [C + 1][BE]   per-channel FFT buffers + one shared
  float2 *Ws = bufs + (size_t)(C + 1) * BE;                         // [C][K]
  float2 *Xs = Ws + (size_t)C * K;                                  // [C][K]
  float2 *Es = Xs + (size_t)C * K;                                  // [K]
  float *Ps = reinterpret_cast<float *>(Es + K);                    // [K]
  float *PW = Ps + K;                                               // [C][K]  |X_c|^2
  float *inb = PW + (size_t)C * K;                                  // [C][N]  [old | new]
  float *dl = inb + (size_t)C * N;                                  // [L + L/2] delay line of the desired signal
  float *es = dl + L + L / 2;                                       // [L] error block
  const int s = blockIdx.x;
  float *st = a.state + (size_t)s * state_elems(C);
  for (int i = tid; i < H; i += nthr) tw_h[i] = a.tw_h[i];
  for (int i = tid; i <= H / 2; i += nthr) tw_n[i] = a.tw_n[i];
  for (int i = tid; i < C * K; i += nthr) Ws[i] = make_float2(st[2 * i], st[2 * i + 1]);
  for (int i = tid; i < K; i += nthr) Ps[i] = st[(size_t)C * K * 2 + i];
  for (int i = tid; i < C * L; i += nthr) inb[(i / L) * N + L + (i % L)] = st[(size_t)C * K * 2 + K + i];   // becomes "old" below
  for (int i = tid; i < L / 2; i += nthr) dl[i] = st[(size_t)C * K * 2 + K + (size_t)C * L + i];
  __syncthreads();

  float2 *buf = bufs + (size_t)warp * BE;
  float *fb = reinterpret_cast<float *>(buf);
  float2 *bufY = bufs + (size_t)C * BE;
  float *fbY = reinterpret_cast<float *>(bufY);
  const float inv_n = 1.0f / (float)N;
  const int nblk = a.Ns / L;
  const float *xs = a.x + ((size_t)s * C + warp) * a.Ns;
  float2 *Wc = Ws + (size_t)warp * K, *Xc = Xs + (size_t)warp * K;
  float *inc = inb + (size_t)warp * N;

  for (int n = 0; n < nblk; ++n) {
    // ---- S1 (channel warps): X_c = rfft([old | new]), |X_c|^2, product X_c W_c ----------------------------
    for (int i = lane; i < L; i += 32) { inc[i] = inc[L + i]; }
    __syncwarp();
    for (int i = lane; i < L; i += 32) inc[L + i] = xs[(size_t)n * L + i];
    __syncwarp();
    for (int i = lane; i < H; i += 32) buf[FPAD<float>(i)] = make_float2(inc[2 * i], inc[2 * i + 1]);
    __syncwarp();
    warp_rfft<N, float>(buf, tw_h, tw_n, lane);
    for (int k = lane; k < K; k += 32) {
      const float2 X = buf[FPAD<float>(k)], W = Wc[k];
      Xc[k] = X;
      PW[warp * K + k] = X.x * X.x + X.y * X.y;
      buf[FPAD<float>(k)] = make_float2(X.x * W.x - X.y * W.y, X.x * W.y + X.y * W.x);
    }
    __syncthreads();
    // ---- S2 (all threads): power recursion and the summed output spectrum ------------------------------------
    for (int k = tid; k < K; k += nthr) {
      float pw = 0.f, yr = 0.f, yi = 0.f;
      for (int c = 0; c < C; ++c) {
        pw += PW[c * K + k];
        const float2 v = bufs[(size_t)c * BE + FPAD<float>(k)];
        yr += v.x; yi += v.y;
      }
      Ps[k] = a.alpha * Ps[k] + (1.0f - a.alpha) * pw;
      bufY[FPAD<float>(k)] = make_float2(yr, yi);
    }
    __syncthreads();
    // ---- S3 (warp 0): y, delayed desired signal, error block and its spectrum --------------------------------
    if (warp == 0) {
      warp_irfft_unscaled<N, float>(bufY, tw_h, tw_n, lane);
      const float *ds_ = a.d + (size_t)s * a.Ns + (size_t)n * L;
      if (a.non_causal) {
        for (int i = lane; i < L; i += 32) dl[L / 2 + i] = ds_[i];          // buffer[-L:] = d
        __syncwarp();
      }
      for (int i = lane; i < L; i += 32) {
        const int idx = L + i;                                              // last L samples of the inverse transform
        const float y = fbY[2 * FPAD<float>(idx >> 1) + (idx & 1)] * inv_n;
        const float dv = a.non_causal ? dl[i] : ds_[i];
        const float e = dv - y;
        es[i] = e;
        a.e[(size_t)s * a.Ns + (size_t)n * L + i] = e;
      }
      __syncwarp();
      if (a.non_causal) {
        float keep[(L / 2 + 31) / 32];
        for (int i = lane, j = 0; i < L / 2; i += 32, ++j) keep[j] = dl[L + i];   // buffer[:L/2] = buffer[-L/2:]
        __syncwarp();
        for (int i = lane, j = 0; i < L / 2; i += 32, ++j) dl[i] = keep[j];
        __syncwarp();
      }
      // E = rfft([0 ... 0 | e])
      for (int i = lane; i < H; i += 32) {
        const int n0 = 2 * i;
        bufY[FPAD<float>(i)] = (n0 >= L) ? make_float2(es[n0 - L], es[n0 + 1 - L]) : make_float2(0.f, 0.f);
      }
      __syncwarp();
      warp_rfft<N, float>(bufY, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) {
        Es[k] = bufY[FPAD<float>(k)];
        if (Ps[k] < 1e-4f) Ps[k] = 1e-4f;                                   // :189
      }
    }
    __syncthreads();
    // ---- S4 (channel warps): constrained gradient, gated update, tap truncation ------------------------------
    for (int k = lane; k < K; k += 32) {
      const float2 X = Xc[k], E = Es[k];
      const float ip = 1.0f / Ps[k];
      buf[FPAD<float>(k)] = make_float2((X.x * E.x + X.y * E.y) * ip, (X.x * E.y - X.y * E.x) * ip);     // conj(X) E / P
    }
    __syncwarp();
    warp_irfft_unscaled<N, float>(buf, tw_h, tw_n, lane);
    for (int i = lane; i < H; i += 32) {                                    // keep the first L samples, scaled by 1/N
      float2 v = buf[FPAD<float>(i)];
      v = (2 * i < L) ? make_float2(v.x * inv_n, v.y * inv_n) : make_float2(0.f, 0.f);
      buf[FPAD<float>(i)] = v;
    }
    __syncwarp();
    warp_rfft<N, float>(buf, tw_h, tw_n, lane);
    const double *pp = a.p ? a.p + ((size_t)s * nblk + n) * K : nullptr;
    for (int k = lane; k < K; k += 32) {
      float gate = pp ? (float)pp[k] : 1.0f;
      if (a.one_minus_p) gate = 1.0f - gate;
      const float step = gate * 2.0f * a.mu;
      const float2 gsp = buf[FPAD<float>(k)];
      float2 W = Wc[k];
      W.x += step * gsp.x; W.y += step * gsp.y;
      Wc[k] = W;
      buf[FPAD<float>(k)] = W;
    }
    __syncwarp();
    if (a.fir_truncate >= 0) {
      warp_irfft_unscaled<N, float>(buf, tw_h, tw_n, lane);
      const int lo = a.fir_truncate, hi = L - a.fir_truncate;
      for (int i = lane; i < H; i += 32) {
        float2 v = buf[FPAD<float>(i)];
        const int n0 = 2 * i;
        v.x = (n0 >= lo && n0 < hi) ? v.x * inv_n : 0.f;
        v.y = (n0 + 1 >= lo && n0 + 1 < hi) ? v.y * inv_n : 0.f;
        buf[FPAD<float>(i)] = v;
      }
      __syncwarp();
      warp_rfft<N, float>(buf, tw_h, tw_n, lane);
      for (int k = lane; k < K; k += 32) Wc[k] = buf[FPAD<float>(k)];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < C * K; i += nthr) { st[2 * i] = Ws[i].x; st[2 * i + 1] = Ws[i].y; }
  for (int i = tid; i < K; i += nthr) st[(size_t)C * K * 2 + i] = Ps[i];
  for (int i = tid; i < C * L; i += nthr) st[(size_t)C * K * 2 + K + i] = inb[(i / L) * N + L + (i % L)];
  for (int i = tid; i < L / 2; i += nthr) st[(size_t)C * K * 2 + K + (size_t)C * L + i] = dl[i];
}

}  // namespace fdaf
}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_fdaf_state_bytes(int n_streams, int n_ch) {
  return (size_t)n_streams * fdaf::state_elems(n_ch) * sizeof(float);
}

int ds_fdaf_run(const ds_fdaf_params *p, void *state, const float *x, const float *d, const double *prob, float *e, void *stream) {
  DS_CHECK_ARG(p && state && x && d && e, "ds_fdaf_run: null argument");
  if (p->frame_len != fdaf::L) { set_error("ds_fdaf_run: compiled for frame_len 256 (n_fft 512)"); return DS_EUNSUPPORTED; }
  DS_CHECK_ARG(p->n_ch >= 1 && p->n_ch <= fdaf::MAXC, "ds_fdaf_run: n_ch must be 1..%d", fdaf::MAXC);
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_samples >= fdaf::L && p->n_samples % fdaf::L == 0,
               "ds_fdaf_run: n_samples must be a positive multiple of frame_len");
  DS_CHECK_ARG(p->fir_truncate < fdaf::L / 2, "ds_fdaf_run: fir_truncate too large");
  TwiddleSet tw;
  int rc = get_twiddles(fdaf::N, &tw);
  if (rc != DS_OK) return rc;
  fdaf::Args a;
  a.state = (float *)state; a.x = x; a.d = d; a.p = prob; a.e = e; a.tw_h = tw.h32; a.tw_n = tw.n32;
  a.S = p->n_streams; a.C = p->n_ch; a.Ns = p->n_samples; a.fir_truncate = p->fir_truncate; a.non_causal = p->non_causal;
  a.one_minus_p = p->one_minus_p; a.mu = (float)p->mu; a.alpha = (float)p->alpha;
  const int C = a.C;
  const size_t smem = (size_t)(fdaf::H + fdaf::H / 2 + 2) * sizeof(float2) + (size_t)(C + 1) * fdaf::BE * sizeof(float2) +
                      (size_t)(2 * C + 1) * fdaf::K * sizeof(float2) + (size_t)(1 + C) * fdaf::K * sizeof(float) +
                      (size_t)C * fdaf::N * sizeof(float) + (size_t)(fdaf::L + fdaf::L / 2 + fdaf::L) * sizeof(float);
  DS_CUDA(cudaFuncSetAttribute(fdaf::fdaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fdaf::fdaf_kernel<<<a.S, 32 * C, smem, (cudaStream_t)stream>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"

namespace ds {
// fixed blocking matrix of TDGSC (TDGSC.py:77-81): out[c] = x[c] - x[c+1], float64 in, float32 out
__global__ void adjacent_diff_kernel(const double *__restrict__ x, float *__restrict__ out, int S, int C, long long N) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * (C - 1) * N) return;
  const long long n = g % N, c = (g / N) % (C - 1), s = g / (N * (C - 1));
  const double *xs = x + (s * C + c) * N + n;
  out[g] = (float)(xs[0] - xs[N]);
}
}  // namespace ds

extern "C" int ds_adjacent_diff_run(int n_streams, int n_ch, long long n_samples, const double *x, float *out, void *stream) {
  DS_CHECK_ARG(x && out && n_streams >= 1 && n_ch >= 2 && n_samples >= 1, "ds_adjacent_diff_run: bad argument");
  const long long items = (long long)n_streams * (n_ch - 1) * n_samples;
  ds::adjacent_diff_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, n_streams, n_ch, n_samples);
  DS_LAUNCH_CHECK();
  return DS_OK;
}
