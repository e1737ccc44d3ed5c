// srp.cu -- SRP-PHAT steered response power over a direction grid
// (doa/srp.py:17-53, MicArray.steering_vector MicArray.py:74-94):
//
//   P[d, t] = sum_k | sum_m conj(a[d,k,m]) * y[k,t,m] / (|y[k,t,m]| + 1e-6) |,   a = exp(-j w_k tau[d,m])
//
// (|a| = 1, so the PHAT normaliser of the reference, applied to conj(a) y, factors
// out of the steering: SURVEY.md a18.)  Per bin this is a [D x M] . [M x T] complex
// product followed by |.| and a sum over bins -- the one dense contraction of the path.
//
// Kernels:
//   phat_kernel        X[T][M][K] c64 -> Yhat[K][T][M] c64 (normalised, bin-major for the contraction)
//   srp_simt_kernel    CUDA-core reference implementation of the contraction (any shape)
//   srp_tc_kernel      tcgen05 / TMEM implementation (srp_tc.cu), used when the shape allows
#include "common.cuh"

namespace ds {

__global__ void phat_kernel(const float2 *__restrict__ X, float2 *__restrict__ Yhat, int T, int M, int K, int phat) {
  // one block per (t, m) row of K bins; transposes to [K][T][M]
  const int tm = blockIdx.x;
  const int t = tm / M, m = tm % M;
  const float2 *src = X + (size_t)tm * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float2 v = src[k];
    if (phat) {
      // y / (|y| + 1e-6)   (srp.py:49-50), evaluated in double like the reference
      const double re = (double)v.x, im = (double)v.y;
      const double inv = 1.0 / (sqrt(re * re + im * im) + 1e-6);
      v.x = (float)(re * inv); v.y = (float)(im * inv);
    }
    Yhat[((size_t)k * T + t) * M + m] = v;
  }
}

constexpr int SRP_DT = 64, SRP_TT = 64, SRP_NT = 256, SRP_MMAX = 16;

// CTA tile: 64 directions x 64 frames, each thread 4 x 4 outputs, loop over bins.
__global__ void __launch_bounds__(SRP_NT) srp_simt_kernel(const float *__restrict__ tau, const float2 *__restrict__ Yhat,
                                                          float *__restrict__ P, int D, int T, int M, int K, float two_f0) {
  __shared__ float2 As[SRP_DT][SRP_MMAX + 1];
  __shared__ float2 Ys[SRP_TT][SRP_MMAX + 1];
  const int d0 = blockIdx.x * SRP_DT, t0 = blockIdx.y * SRP_TT;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;      // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < K; ++k) {
    // steering for this bin: a = exp(-j 2 pi f_k tau) = cospi(2 f_k tau) - j sinpi(2 f_k tau)
    for (int i = tid; i < SRP_DT * M; i += SRP_NT) {
      const int dd = i / M, m = i % M;
      const int d = d0 + dd;
      float s = 0.f, c = 1.f;
      if (d < D) sincospif(two_f0 * (float)k * tau[(size_t)d * M + m], &s, &c);
      As[dd][m] = make_float2(c, -s);
    }
    for (int i = tid; i < SRP_TT * M; i += SRP_NT) {
      const int tt = i / M, m = i % M;
      const int t = t0 + tt;
      Ys[tt][m] = (t < T) ? Yhat[((size_t)k * T + t) * M + m] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float zr[4][4], zi[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { zr[i][j] = 0.f; zi[i][j] = 0.f; }
    for (int m = 0; m < M; ++m) {
      float2 av[4], yv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[ty * 4 + i][m];
#pragma unroll
      for (int j = 0; j < 4; ++j) yv[j] = Ys[tx * 4 + j][m];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // conj(a) * y
          zr[i][j] = fmaf(av[i].x, yv[j].x, fmaf(av[i].y, yv[j].y, zr[i][j]));
          zi[i][j] = fmaf(av[i].x, yv[j].y, fmaf(-av[i].y, yv[j].x, zi[i][j]));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += sqrtf(fmaf(zr[i][j], zr[i][j], zi[i][j] * zi[i][j]));
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = d0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (d < D && t < T) P[(size_t)d * T + t] = acc[i][j];
    }
  }
}

int srp_tc_launch(const float *tau, const float2 *Yhat, void *workspace, float *P, int D, int T, int M, int K, float two_f0,
                  cudaStream_t st);
size_t srp_tc_workspace_bytes(int T, int M, int K);
bool srp_tc_supported(int D, int T, int M, int K);

}  // namespace ds

using namespace ds;

extern "C" {

int ds_phat_run(int n_frames, int n_mics, int n_bins, int phat, const void *X, void *Yhat, void *stream) {
  DS_CHECK_ARG(X && Yhat, "ds_phat_run: null argument");
  DS_CHECK_ARG(n_frames >= 1 && n_mics >= 1 && n_bins >= 1, "ds_phat_run: bad shape");
  phat_kernel<<<n_frames * n_mics, 128, 0, (cudaStream_t)stream>>>((const float2 *)X, (float2 *)Yhat, n_frames, n_mics, n_bins, phat);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

size_t ds_srp_workspace_bytes(int n_frames, int n_mics, int n_bins, int use_tensor_cores) {
  if (!use_tensor_cores || !srp_tc_supported(1, n_frames, n_mics, n_bins)) return 0;
  return srp_tc_workspace_bytes(n_frames, n_mics, n_bins);
}

int ds_srp_run(int n_dirs, int n_frames, int n_mics, int n_bins, double fs, int n_fft, const float *tau, const void *Yhat,
               void *workspace, float *P, int use_tensor_cores, void *stream) {
  DS_CHECK_ARG(tau && Yhat && P, "ds_srp_run: null argument");
  DS_CHECK_ARG(n_dirs >= 1 && n_frames >= 1 && n_bins >= 1 && n_mics >= 1 && n_mics <= SRP_MMAX,
               "ds_srp_run: n_mics must be 1..%d", SRP_MMAX);
  const float two_f0 = (float)(2.0 * fs / (double)n_fft);       // 2 f_k = two_f0 * k
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tensor_cores) {
    if (!srp_tc_supported(n_dirs, n_frames, n_mics, n_bins)) {
      set_error("ds_srp_run: tensor-core path needs n_mics in {4, 8, 16}");
      return DS_EUNSUPPORTED;
    }
    DS_CHECK_ARG(workspace && (reinterpret_cast<size_t>(workspace) & 127) == 0, "ds_srp_run: the tensor-core path needs a 128-byte aligned workspace");
    return srp_tc_launch(tau, (const float2 *)Yhat, workspace, P, n_dirs, n_frames, n_mics, n_bins, two_f0, st);
  }
  dim3 grid((n_dirs + SRP_DT - 1) / SRP_DT, (n_frames + SRP_TT - 1) / SRP_TT);
  srp_simt_kernel<<<grid, SRP_NT, 0, st>>>(tau, (const float2 *)Yhat, P, n_dirs, n_frames, n_mics, n_bins, two_f0);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
