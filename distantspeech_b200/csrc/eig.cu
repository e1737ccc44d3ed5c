// eig.cu -- data-driven steering and mask-based / GEV beamformer weights:
//   ds_masked_cov_run        example/mvdr.ipynb cell 6 (and the frame-range averages of cell 2)
//   ds_steering_run          steering()                     beamformer/beamformer.py:10-31
//   ds_gev_run               get_gev_vector()               beamformer/beamformer.py:77-97
//   ds_phase_correction_run  phase_correction()             beamformer/beamformer.py:64-74
//   ds_ban_run               blind_analytic_normalization() beamformer/beamformer.py:34-61
//
// One (stream, bin) matrix per thread; the eigen-solvers keep their matrices in shared memory,
// interleaved over the threads of the CTA (eig_core.cuh).  These run once per utterance (or once
// per frame for the online variant of cell 4), not per sample: 2.6e5 8x8 problems for 1024 streams,
// a few tens of kFLOP each -- latency-bound fp64, far off every roofline; nothing to tile.
#include "common.cuh"
#include "eig_core.cuh"

namespace ds {

constexpr int EIG_NT = 32;   // threads (= matrices) per CTA: 6 M^2 doubles each -> 96 KB at M = 8

static __device__ __forceinline__ void load_cmat(const CMatRef &A, const double2 *src) {
  for (int i = 0; i < A.M; ++i)
    for (int j = 0; j < A.M; ++j) { const double2 v = src[i * A.M + j]; A.r(i, j) = v.x; A.c(i, j) = v.y; }
}
static __device__ __forceinline__ void set_identity(const CMatRef &V) {
  for (int i = 0; i < V.M; ++i)
    for (int j = 0; j < V.M; ++j) { V.r(i, j) = i == j ? 1.0 : 0.0; V.c(i, j) = 0.0; }
}
// column `col` of V times conj(e^{i angle(V[0][col])}): component 0 becomes real and non-negative
static __device__ __forceinline__ void reference_phase(const CMatRef &V, int col) {
  const double v0r = V.r(0, col), v0i = V.c(0, col), n = sqrt(v0r * v0r + v0i * v0i);
  const double er = n > 0.0 ? v0r / n : 1.0, ei = n > 0.0 ? v0i / n : 0.0;   // np.angle(0) = 0
  for (int i = 0; i < V.M; ++i) {
    const double xr = V.r(i, col) * er + V.c(i, col) * ei, xi = V.c(i, col) * er - V.r(i, col) * ei;
    V.r(i, col) = xr; V.c(i, col) = xi;
  }
}

__global__ void __launch_bounds__(EIG_NT) steering_kernel(long long n, int M, const double2 *XXs, double2 *out) {
  extern __shared__ double sm[];
  const long long g = (long long)blockIdx.x * EIG_NT + threadIdx.x;
  if (g >= n) return;
  const int MM = M * M;
  double *base = sm + threadIdx.x;
  const CMatRef A{base, base + MM * EIG_NT, M, EIG_NT}, V{base + 2 * MM * EIG_NT, base + 3 * MM * EIG_NT, M, EIG_NT};
  load_cmat(A, XXs + g * MM);
  herm_from_lower(A);
  set_identity(V);
  jacobi_hermitian(A, V);
  const int col = argmax_diag(A);
  reference_phase(V, col);                                                   // beamformer.py:27-29
  for (int i = 0; i < M; ++i) out[g * M + i] = make_double2(V.r(i, col), V.c(i, col));
}

__global__ void __launch_bounds__(EIG_NT) gev_kernel(long long n, int M, const double2 *target, const double2 *noise,
                                                     double2 *out) {
  extern __shared__ double sm[];
  const long long g = (long long)blockIdx.x * EIG_NT + threadIdx.x;
  if (g >= n) return;
  const int MM = M * M;
  double *base = sm + threadIdx.x;
  const CMatRef A{base, base + MM * EIG_NT, M, EIG_NT}, B{base + 2 * MM * EIG_NT, base + 3 * MM * EIG_NT, M, EIG_NT},
      V{base + 4 * MM * EIG_NT, base + 5 * MM * EIG_NT, M, EIG_NT};
  load_cmat(A, target + g * MM);
  load_cmat(B, noise + g * MM);
  herm_from_lower(A);
  if (!cholesky_lower(B)) {
    // the reference's LinAlgError branch (beamformer.py:94-96): ones / trace(noise) * sensors
    double tr = 0.0, ti = 0.0;
    for (int i = 0; i < M; ++i) { const double2 v = noise[g * MM + i * M + i]; tr += v.x; ti += v.y; }
    const double d = tr * tr + ti * ti;
    for (int i = 0; i < M; ++i) out[g * M + i] = make_double2(M * tr / d, -M * ti / d);
    return;
  }
  reduce_to_standard(A, B);
  herm_from_lower(A);
  set_identity(V);
  jacobi_hermitian(A, V);
  const int col = argmax_diag(A);
  reference_phase(V, col);          // LAPACK leaves component 0 of the standard-form vector real; sign fixed to +
  if (M == 1) { out[g] = make_double2(V.r(0, 0) / B.r(0, 0), V.c(0, 0) / B.r(0, 0)); return; }
  // w = L^-H x, back-substituted into another (no longer needed) column of V
  const int dst = col == 0 ? 1 : 0;
  for (int i = M - 1; i >= 0; --i) {
    double sr = V.r(i, col), si = V.c(i, col);
    for (int k = i + 1; k < M; ++k) {          // - conj(L_ki) w_k
      sr -= B.r(k, i) * V.r(k, dst) + B.c(k, i) * V.c(k, dst);
      si -= B.r(k, i) * V.c(k, dst) - B.c(k, i) * V.r(k, dst);
    }
    V.r(i, dst) = sr / B.r(i, i); V.c(i, dst) = si / B.r(i, i);
  }
  for (int i = 0; i < M; ++i) out[g * M + i] = make_double2(V.r(i, dst), V.c(i, dst));
}

// w[f, :] *= exp(-j angle(sum_m w[f, m] conj(w[f-1, m]))), f = 1..F-1, each bin against the already
// corrected previous one: sequential over bins, one thread per stream.
__global__ void phase_correction_kernel(int S, int F, int M, double2 *W) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double2 *w = W + (long long)s * F * M;
  for (int f = 1; f < F; ++f) {
    double dr = 0.0, di = 0.0;
    for (int m = 0; m < M; ++m) {
      const double2 a = w[f * M + m], b = w[(f - 1) * M + m];
      dr += a.x * b.x + a.y * b.y;
      di += a.y * b.x - a.x * b.y;
    }
    const double n = sqrt(dr * dr + di * di);
    const double er = n > 0.0 ? dr / n : 1.0, ei = n > 0.0 ? di / n : 0.0;
    for (int m = 0; m < M; ++m) {              // times conj(e)
      const double2 a = w[f * M + m];
      w[f * M + m] = make_double2(a.x * er + a.y * ei, a.y * er - a.x * ei);
    }
  }
}

// vector * |sqrt(v^H N N v)| / (|v^H N v| + eps)
__global__ void ban_kernel(long long n, int M, const double2 *vec, const double2 *noise, double eps, double2 *out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const double2 *v = vec + g * M, *N = noise + g * M * M;
  double nr = 0.0, ni = 0.0, dr = 0.0, di = 0.0;
  for (int b = 0; b < M; ++b) {
    double rr = 0.0, ri = 0.0, tr = 0.0, ti = 0.0;   // r_b = sum_a conj(v_a) N_ab ; t_b = sum_c N_bc v_c
    for (int a = 0; a < M; ++a) {
      const double2 va = v[a], Nab = N[a * M + b], Nba = N[b * M + a];
      rr += va.x * Nab.x + va.y * Nab.y; ri += va.x * Nab.y - va.y * Nab.x;
      tr += Nba.x * va.x - Nba.y * va.y; ti += Nba.x * va.y + Nba.y * va.x;
    }
    nr += rr * tr - ri * ti; ni += rr * ti + ri * tr;
    dr += v[b].x * tr + v[b].y * ti; di += v[b].x * ti - v[b].y * tr;
  }
  const double scale = sqrt(sqrt(nr * nr + ni * ni)) / (sqrt(dr * dr + di * di) + eps);
  for (int m = 0; m < M; ++m) out[g * M + m] = make_double2(v[m].x * scale, v[m].y * scale);
}

// Phi_xx[s,k,i,:] += scale * sum_t p y_i conj(y_:) ; Phi_vv += scale * sum_t (1 - p) y_i conj(y_:).
// One thread per (stream, row i, bin k) -- bin innermost so that the spectrum loads coalesce;
// frames in order, each term rounded like the reference's einsum(...) * p before it is added.
template <typename CT, int M>
__global__ void masked_cov_kernel(int S, int T, int K, int t0, int t1, const CT *X, const double *p, double scale,
                                  double2 *Pxx, double2 *Pvv) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * M * K) return;
  const int k = (int)(g % K), i = (int)((g / K) % M), s = (int)(g / ((long long)K * M));
  double xr[M], xi[M], vr[M], vi[M];
#pragma unroll
  for (int j = 0; j < M; ++j) xr[j] = xi[j] = vr[j] = vi[j] = 0.0;
  for (int t = t0; t < t1; ++t) {
    const CT *y = X + ((long long)s * T + t) * M * K + k;
    const double pt = p ? p[((long long)s * T + t) * K + k] : 1.0;
    const double qt = 1.0 - pt;
    const double ar = (double)y[(long long)i * K].x, ai = (double)y[(long long)i * K].y;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double br = (double)y[(long long)j * K].x, bi = (double)y[(long long)j * K].y;
      const double cr = ar * br + ai * bi, ci = ai * br - ar * bi;            // y_i conj(y_j)
      xr[j] += cr * pt; xi[j] += ci * pt;
      if (Pvv) { vr[j] += cr * qt; vi[j] += ci * qt; }
    }
  }
  const long long o = (((long long)s * K + k) * M + i) * M;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double2 a = Pxx[o + j];
    a.x += scale * xr[j]; a.y += scale * xi[j];
    Pxx[o + j] = a;
    if (Pvv) { double2 b = Pvv[o + j]; b.x += scale * vr[j]; b.y += scale * vi[j]; Pvv[o + j] = b; }
  }
}

// w = R^-1 a / (a^H R^-1 a) straight from the covariance (Cholesky + two triangular solves) -- what
// compute_mvdr_weight(steer, np.linalg.inv(Phi_vv)) evaluates in example/mvdr.ipynb cell 6, without
// forming the inverse.  A matrix that is not positive definite yields NaN weights.
__global__ void __launch_bounds__(EIG_NT) mvdr_from_cov_kernel(long long n, int M, const double2 *steer, const double2 *R,
                                                               double2 *out) {
  extern __shared__ double sm[];
  const long long g = (long long)blockIdx.x * EIG_NT + threadIdx.x;
  if (g >= n) return;
  const int MM = M * M;
  double *base = sm + threadIdx.x;
  const CMatRef B{base, base + MM * EIG_NT, M, EIG_NT};
  double *zr = base + 2 * MM * EIG_NT, *zi = zr + M * EIG_NT;      // [i * EIG_NT]
  load_cmat(B, R + g * MM);
  if (!cholesky_lower(B)) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < M; ++i) out[g * M + i] = make_double2(nan, nan);
    return;
  }
  for (int i = 0; i < M; ++i) {                 // z1 = L^-1 a
    double sr = steer[g * M + i].x, si = steer[g * M + i].y;
    for (int k = 0; k < i; ++k) {
      sr -= B.r(i, k) * zr[k * EIG_NT] - B.c(i, k) * zi[k * EIG_NT];
      si -= B.r(i, k) * zi[k * EIG_NT] + B.c(i, k) * zr[k * EIG_NT];
    }
    zr[i * EIG_NT] = sr / B.r(i, i); zi[i * EIG_NT] = si / B.r(i, i);
  }
  for (int i = M - 1; i >= 0; --i) {            // z = L^-H z1
    double sr = zr[i * EIG_NT], si = zi[i * EIG_NT];
    for (int k = i + 1; k < M; ++k) {
      sr -= B.r(k, i) * zr[k * EIG_NT] + B.c(k, i) * zi[k * EIG_NT];
      si -= B.r(k, i) * zi[k * EIG_NT] - B.c(k, i) * zr[k * EIG_NT];
    }
    zr[i * EIG_NT] = sr / B.r(i, i); zi[i * EIG_NT] = si / B.r(i, i);
  }
  double dr = 0.0, di = 0.0;                    // a^H z
  for (int i = 0; i < M; ++i) {
    const double2 a = steer[g * M + i];
    dr += a.x * zr[i * EIG_NT] + a.y * zi[i * EIG_NT];
    di += a.x * zi[i * EIG_NT] - a.y * zr[i * EIG_NT];
  }
  const double d2 = dr * dr + di * di;
  for (int i = 0; i < M; ++i) {                 // z / (a^H z)
    const double xr = zr[i * EIG_NT], xi = zi[i * EIG_NT];
    out[g * M + i] = make_double2((xr * dr + xi * di) / d2, (xi * dr - xr * di) / d2);
  }
}

// Y[s,t,k] = sum_m conj(W[s,k,m]) X[s,t,m,k]: per-stream weights (the mask-based beamformers design one
// weight set per utterance); einsum('inj,ij->in', D, w.conj()) of mvdr.ipynb cells 6 / 8.
template <typename XT>
__global__ void apply_stream_weights_kernel(int S, int T, int M, int K, const XT *__restrict__ X,
                                            const double2 *__restrict__ W, double2 *__restrict__ Y) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * T * K) return;
  const long long st = g / K;
  const int k = (int)(g % K), s = (int)(st / T);
  double yr = 0.0, yi = 0.0;
  for (int m = 0; m < M; ++m) {
    const XT x = X[(st * M + m) * K + k];
    const double2 w = W[((long long)s * K + k) * M + m];
    yr += w.x * (double)x.x + w.y * (double)x.y;
    yi += w.x * (double)x.y - w.y * (double)x.x;
  }
  Y[g] = make_double2(yr, yi);
}

template <typename CT>
static int launch_masked_cov(int S, int T, int M, int K, int t0, int t1, const void *X, const double *p, double scale,
                             void *Pxx, void *Pvv, cudaStream_t st) {
  const long long n = (long long)S * M * K;
  const unsigned grid = (unsigned)((n + 127) / 128);
#define DS_MC(MM)                                                                                              \
  case MM: masked_cov_kernel<CT, MM><<<grid, 128, 0, st>>>(S, T, K, t0, t1, (const CT *)X, p, scale, (double2 *)Pxx, \
                                                            (double2 *)Pvv); break;
  switch (M) { DS_MC(2) DS_MC(3) DS_MC(4) DS_MC(5) DS_MC(6) DS_MC(7) DS_MC(8)
    default: set_error("masked_cov: n_mics %d outside the compiled range 2..8", M); return DS_EUNSUPPORTED; }
#undef DS_MC
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // namespace ds

using namespace ds;

extern "C" int ds_masked_cov_run(int n_streams, int n_frames, int n_mics, int n_bins, int t0, int t1, const void *X,
                                 int x_is_c128, const double *p, double scale, void *Phi_xx, void *Phi_vv,
                                 void *stream) {
  DS_CHECK_ARG(X && Phi_xx, "masked_cov: null pointer");
  DS_CHECK_ARG(n_streams > 0 && n_frames > 0 && n_bins > 0, "masked_cov: empty shape");
  DS_CHECK_ARG(0 <= t0 && t0 <= t1 && t1 <= n_frames, "masked_cov: frame range [%d, %d) outside [0, %d)", t0, t1, n_frames);
  DS_CHECK_ARG(p || !Phi_vv, "masked_cov: Phi_vv needs a mask");
  cudaStream_t st = (cudaStream_t)stream;
  return x_is_c128 ? launch_masked_cov<double2>(n_streams, n_frames, n_mics, n_bins, t0, t1, X, p, scale, Phi_xx, Phi_vv, st)
                   : launch_masked_cov<float2>(n_streams, n_frames, n_mics, n_bins, t0, t1, X, p, scale, Phi_xx, Phi_vv, st);
}

static int eig_smem(int M, int mats, size_t *bytes) {
  if (M < 1 || M > 8) { set_error("eig: n_mics %d outside 1..8", M); return DS_EUNSUPPORTED; }
  *bytes = (size_t)mats * 2 * M * M * EIG_NT * sizeof(double);
  return DS_OK;
}

extern "C" int ds_steering_run(long long n, int n_mics, const void *XXs, void *out, void *stream) {
  DS_CHECK_ARG(XXs && out, "steering: null pointer");
  DS_CHECK_ARG(n > 0, "steering: empty batch");
  size_t smem;
  int rc = eig_smem(n_mics, 2, &smem);
  if (rc) return rc;
  DS_CUDA(cudaFuncSetAttribute(steering_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  steering_kernel<<<(unsigned)((n + EIG_NT - 1) / EIG_NT), EIG_NT, smem, (cudaStream_t)stream>>>(
      n, n_mics, (const double2 *)XXs, (double2 *)out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_gev_run(long long n, int n_mics, const void *target, const void *noise, void *out, void *stream) {
  DS_CHECK_ARG(target && noise && out, "gev: null pointer");
  DS_CHECK_ARG(n > 0, "gev: empty batch");
  size_t smem;
  int rc = eig_smem(n_mics, 3, &smem);
  if (rc) return rc;
  DS_CUDA(cudaFuncSetAttribute(gev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gev_kernel<<<(unsigned)((n + EIG_NT - 1) / EIG_NT), EIG_NT, smem, (cudaStream_t)stream>>>(
      n, n_mics, (const double2 *)target, (const double2 *)noise, (double2 *)out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_phase_correction_run(int n_streams, int n_bins, int n_mics, void *W, void *stream) {
  DS_CHECK_ARG(W, "phase_correction: null pointer");
  DS_CHECK_ARG(n_streams > 0 && n_bins > 0 && n_mics > 0, "phase_correction: empty shape");
  phase_correction_kernel<<<(n_streams + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n_streams, n_bins, n_mics, (double2 *)W);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_ban_run(long long n, int n_mics, const void *vector, const void *noise, double eps, void *out,
                          void *stream) {
  DS_CHECK_ARG(vector && noise && out, "ban: null pointer");
  DS_CHECK_ARG(n > 0 && n_mics > 0, "ban: empty shape");
  ban_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, n_mics, (const double2 *)vector,
                                                                           (const double2 *)noise, eps, (double2 *)out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_mvdr_from_cov_run(long long n, int n_mics, const void *steer, const void *Rvv, void *w_out,
                                    void *stream) {
  DS_CHECK_ARG(steer && Rvv && w_out, "mvdr_from_cov: null pointer");
  DS_CHECK_ARG(n > 0, "mvdr_from_cov: empty batch");
  if (n_mics < 1 || n_mics > 8) { set_error("mvdr_from_cov: n_mics %d outside 1..8", n_mics); return DS_EUNSUPPORTED; }
  const size_t smem = (size_t)(2 * n_mics * n_mics + 2 * n_mics) * EIG_NT * sizeof(double);
  DS_CUDA(cudaFuncSetAttribute(mvdr_from_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mvdr_from_cov_kernel<<<(unsigned)((n + EIG_NT - 1) / EIG_NT), EIG_NT, smem, (cudaStream_t)stream>>>(
      n, n_mics, (const double2 *)steer, (const double2 *)Rvv, (double2 *)w_out);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

extern "C" int ds_apply_stream_weights_run(int n_streams, int n_frames, int n_mics, int n_bins, const void *X,
                                           int x_is_c128, const void *W, void *Y, void *stream) {
  DS_CHECK_ARG(X && W && Y, "apply_stream_weights: null pointer");
  DS_CHECK_ARG(n_streams > 0 && n_frames > 0 && n_mics > 0 && n_bins > 0, "apply_stream_weights: empty shape");
  const long long n = (long long)n_streams * n_frames * n_bins;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_is_c128)
    apply_stream_weights_kernel<double2><<<blocks, 256, 0, st>>>(n_streams, n_frames, n_mics, n_bins, (const double2 *)X,
                                                                 (const double2 *)W, (double2 *)Y);
  else
    apply_stream_weights_kernel<float2><<<blocks, 256, 0, st>>>(n_streams, n_frames, n_mics, n_bins, (const float2 *)X,
                                                                (const double2 *)W, (double2 *)Y);
  DS_LAUNCH_CHECK();
  return DS_OK;
}
