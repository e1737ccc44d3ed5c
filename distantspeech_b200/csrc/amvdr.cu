// amvdr.cu -- online MVDR with an MCRA-VAD-gated noise covariance: the
// run_MVDRbeamformer.py path, adaptivebeamfomer.process
// (beamformer/adaptivebeamformer.py:44-128), one thread per (stream, bin),
// frames sequential, all recursive state float64 like the reference.
//
// Per frame and bin (file:line of the reference):
//   mcra.estimation(|Z0|^2)                                   :81  -> mcra.py:27-77
//   Ryy = 0.8 Ryy + 0.2 z z^H                                 :86
//   if mcra.p[k] < 0.4:                                       :94
//       Rvv = 0.9998 Rvv + 0.0002 z z^H                       :95-98
//       Rvv_inv = inv(Rvv + 1e-6 I)                           :101-104   (stale elsewhere, quirk 9)
//   H = getweights(a, 'MVDR'|'TFGSC'|..)                      :105 -> beamformer.py:306-336
//   Y = sum_m conj(H_m) z_m                                   :119-120
#include <type_traits>
#include "common.cuh"
#include "perbin.cuh"
#include "herm.cuh"

namespace ds {

struct AmvdrArgs {
  double *state;            // [S][NE][K]
  const double2 *a;         // [M][K] propagation vectors exp(-j w_k tau_m)
  const float2 *X;          // [S][T][M][K]
  float2 *Yout;             // [S][T][K]
  double2 *H_last;          // [S][K][M] weights of the last frame (or null)
  double *p_out;            // [S][T][K] MCRA speech presence (or null)
  int S, K, T, frm_cnt, ell, method;
  double alpha_y, alpha_v, diag, vad_thr;
  McraConst mc;
};

// state element order (doubles): Rvv[M*M] Rinv[M*M] Ryy[M*M] mcra[5]
// Hermitian packing of an MxM matrix into M*M doubles: diag[M], then (re, im) of the
// strictly upper entries in qidx order.
template <int M> __host__ __device__ constexpr int amvdr_state_elems() { return 3 * M * M + 5; }

// resident CTAs per SM the register allocation aims for (the state lives in shared memory: 3 M^2 doubles per thread)
#ifndef AMVDR_MINB4
#define AMVDR_MINB4 3          // A/B on the B200 (config 1, ms per step): 2 -> 37.5 (255 registers), 3 -> 33.8 (168), 4 -> 35.5 (128, spills)
#endif
template <int M> __host__ __device__ constexpr int amvdr_minb() { return M <= 4 ? AMVDR_MINB4 : (M <= 6 ? 3 : 2); }

template <int M, int NT>
__global__ void __launch_bounds__(NT, amvdr_minb<M>()) amvdr_kernel(AmvdrArgs a) {
  constexpr int NQ = M * (M - 1) / 2, MM = M * M;
  constexpr int NE = amvdr_state_elems<M>();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  const int K = a.K;
  const long long g = (long long)blockIdx.x * NT + tid;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  double *blob = a.state + (long long)s * NE * K + k;
  // shared-memory views, element e of thread tid at base[e * NT]
  double *sv = sm + tid;                  // Rvv
  double *si = sm + MM * NT + tid;        // Rvv_inv
  double *sy = sm + 2 * MM * NT + tid;    // Ryy
  for (int e = 0; e < 3 * MM; ++e) sm[e * NT + tid] = blob[(long long)e * K];
  double mS = blob[(long long)(3 * MM + 0) * K], mSmin = blob[(long long)(3 * MM + 1) * K],
         mStmp = blob[(long long)(3 * MM + 2) * K], mp = blob[(long long)(3 * MM + 3) * K],
         mlam = blob[(long long)(3 * MM + 4) * K];
  double ar[M], ai[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { const double2 v = a.a[(long long)m * K + k]; ar[m] = v.x; ai[m] = v.y; }
  int frm = a.frm_cnt, ell = a.ell % a.mc.L;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
  const double ay = a.alpha_y, one_m_ay = 1.0 - a.alpha_y, av = a.alpha_v, one_m_av = 1.0 - a.alpha_v;

  startup_dephase(8000);
  // the spectrum of frame t + 1 is loaded while frame t is processed, frames further ahead are pulled into L2
  // (ncu before: 1.6 long-scoreboard stalls per issue on these loads)
  const float2 *Xp = a.X + (long long)s * a.T * M * K + k;
  float2 zn[M], znb0 = make_float2(0.f, 0.f), znb1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < M; ++m) zn[m] = Xp[(long long)m * K];
  if (k > 0) znb0 = Xp[-1];
  if (k < K - 1) znb1 = Xp[1];
  for (int t = 0; t < a.T; ++t) {
    double zr[M], zi[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { zr[m] = (double)zn[m].x; zi[m] = (double)zn[m].y; }
    const double Ym1 = (k > 0) ? power_c((double)znb0.x, (double)znb0.y) : 0.0;
    const double Yp1 = (k < K - 1) ? power_c((double)znb1.x, (double)znb1.y) : 0.0;
    Xp += (long long)M * K;
    if (t + 1 < a.T) {
#pragma unroll
      for (int m = 0; m < M; ++m) zn[m] = Xp[(long long)m * K];
      if (k > 0) znb0 = Xp[-1];
      if (k < K - 1) znb1 = Xp[1];
      if (t + 5 < a.T) {
#pragma unroll
        for (int m = 0; m < M; ++m) asm volatile("prefetch.global.L2 [%0];" ::"l"(Xp + (long long)4 * M * K + (long long)m * K));
      }
    }
    const double Y0 = power_c(zr[0], zi[0]);
    const bool reset = (frm > 0) && (ell == 0);
    mcra_step(mS, mSmin, mStmp, mp, mlam, Ym1, Y0, Yp1, k, K, frm, reset, a.mc);
    if (reset) ell = 0;
    ++ell; ++frm;
    if (ell == a.mc.L) ell = 0;
    if (a.p_out) a.p_out[((long long)s * a.T + t) * K + k] = mp;

    // ---- Ryy = alpha_y Ryy + (1 - alpha_y) z z^H                     :86
#pragma unroll
    for (int i = 0; i < M; ++i) {
      sy[i * NT] = fma(ay, sy[i * NT], one_m_ay * fma(zi[i], zi[i], zr[i] * zr[i]));
#pragma unroll
      for (int j = i + 1; j < M; ++j) {
        const int e = M + 2 * qidx<M>(i, j);
        // z_i conj(z_j)
        sy[e * NT] = fma(ay, sy[e * NT], one_m_ay * fma(zi[i], zi[j], zr[i] * zr[j]));
        sy[(e + 1) * NT] = fma(ay, sy[(e + 1) * NT], one_m_ay * fma(zi[i], zr[j], -zr[i] * zi[j]));
      }
    }

    // ---- VAD-gated noise covariance and its inverse                  :94-104
    if (mp < a.vad_thr) {
      Herm<M> h;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const double v = fma(av, sv[i * NT], one_m_av * fma(zi[i], zi[i], zr[i] * zr[i]));
        sv[i * NT] = v;
        h.d[i] = v + a.diag;
#pragma unroll
        for (int j = i + 1; j < M; ++j) {
          const int q = qidx<M>(i, j), e = M + 2 * q;
          const double vr = fma(av, sv[e * NT], one_m_av * fma(zi[i], zi[j], zr[i] * zr[j]));
          const double vi = fma(av, sv[(e + 1) * NT], one_m_av * fma(zi[i], zr[j], -zr[i] * zi[j]));
          sv[e * NT] = vr; sv[(e + 1) * NT] = vi;
          h.ur[q] = vr; h.ui[q] = vi;
        }
      }
      herm_inverse<M>(h);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        si[i * NT] = h.d[i];
#pragma unroll
        for (int j = i + 1; j < M; ++j) {
          const int q = qidx<M>(i, j), e = M + 2 * q;
          si[e * NT] = h.ur[q]; si[(e + 1) * NT] = h.ui[q];
        }
      }
    }

    // ---- weights                                                      beamformer.py:318-333
    double wr[M], wi[M];
    if (a.method == 2 || a.method == 3) {
      // b = Rinv v ; v = a (MVDR) or Ryy[:,0] (TFGSC)
      double vr[M], vi[M];
      if (a.method == 2) {
#pragma unroll
        for (int m = 0; m < M; ++m) { vr[m] = ar[m]; vi[m] = ai[m]; }
      } else {
        vr[0] = sy[0]; vi[0] = 0.0;
#pragma unroll
        for (int m = 1; m < M; ++m) {         // Ryy[m][0] = conj(Ryy[0][m])
          const int e = M + 2 * qidx<M>(0, m);
          vr[m] = sy[e * NT]; vi[m] = -sy[(e + 1) * NT];
        }
      }
      double br[M], bi[M];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double sr = si[i * NT] * vr[i], sim = si[i * NT] * vi[i];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          if (j == i) continue;
          const int lo = i < j ? i : j, hi = i < j ? j : i;
          const int e = M + 2 * qidx<M>(lo, hi);
          const double rr = si[e * NT];
          const double ri = (i < j) ? si[(e + 1) * NT] : -si[(e + 1) * NT];
          sr = fma(rr, vr[j], fma(-ri, vi[j], sr));
          sim = fma(rr, vi[j], fma(ri, vr[j], sim));
        }
        br[i] = sr; bi[i] = sim;
      }
      if (a.method == 2) {
        // w = b / (a^H b)
        double dr = 0.0, di = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
          dr = fma(ar[i], br[i], fma(ai[i], bi[i], dr));
          di = fma(ar[i], bi[i], fma(-ai[i], br[i], di));
        }
        const double dn = 1.0 / (dr * dr + di * di);
#pragma unroll
        for (int i = 0; i < M; ++i) {
          wr[i] = (br[i] * dr + bi[i] * di) * dn;
          wi[i] = (bi[i] * dr - br[i] * di) * dn;
        }
      } else {
        // TFGSC: w = ((Rinv Ryy) - I) u / (trace(Rinv Ryy) - M)        beamformer.py:327-333
        double trr = 0.0, tri = 0.0;      // trace(Rinv Ryy) = sum_ij Rinv_ij Ryy_ji
#pragma unroll
        for (int i = 0; i < M; ++i) {
          trr = fma(si[i * NT], sy[i * NT], trr);
#pragma unroll
          for (int j = i + 1; j < M; ++j) {
            const int e = M + 2 * qidx<M>(i, j);
            // Rinv_ij Ryy_ji + Rinv_ji Ryy_ij = 2 Re(Rinv_ij conj(Ryy_ij))
            trr = fma(2.0, fma(si[e * NT], sy[e * NT], si[(e + 1) * NT] * sy[(e + 1) * NT]), trr);
          }
        }
        const double dr = trr - (double)M, di = tri;
        const double dn = 1.0 / (dr * dr + di * di);
        br[0] -= 1.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
          wr[i] = (br[i] * dr + bi[i] * di) * dn;
          wi[i] = (bi[i] * dr - br[i] * di) * dn;
        }
      }
    } else if (a.method == 1) {            // DS: a / M
#pragma unroll
      for (int m = 0; m < M; ++m) { wr[m] = ar[m] / (double)M; wi[m] = ai[m] / (double)M; }
    } else {                               // src: a with weights[1:] = 0
#pragma unroll
      for (int m = 0; m < M; ++m) { wr[m] = (m == 0) ? ar[0] : 0.0; wi[m] = (m == 0) ? ai[0] : 0.0; }
    }

    // ---- Y = sum_m conj(H_m) z_m                                      :119-120
    double yr = 0.0, yi = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      yr = fma(wr[m], zr[m], fma(wi[m], zi[m], yr));
      yi = fma(wr[m], zi[m], fma(-wi[m], zr[m], yi));
    }
    a.Yout[((long long)s * a.T + t) * K + k] = make_float2((float)yr, (float)yi);
    if (a.H_last && t == a.T - 1) {
#pragma unroll
      for (int m = 0; m < M; ++m) a.H_last[((long long)s * K + k) * M + m] = make_double2(wr[m], wi[m]);
    }
  }
  for (int e = 0; e < 3 * MM; ++e) blob[(long long)e * K] = sm[e * NT + tid];
  blob[(long long)(3 * MM + 0) * K] = mS; blob[(long long)(3 * MM + 1) * K] = mSmin; blob[(long long)(3 * MM + 2) * K] = mStmp;
  blob[(long long)(3 * MM + 3) * K] = mp; blob[(long long)(3 * MM + 4) * K] = mlam;
}

template <int M>
static int launch_amvdr_m(const AmvdrArgs &a, cudaStream_t st) {
  constexpr int NT = (M <= 4) ? 128 : 64;
  const size_t smem = (size_t)3 * M * M * NT * sizeof(double);
  auto kern = amvdr_kernel<M, NT>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)a.S * a.K;
  kern<<<(unsigned)((items + NT - 1) / NT), NT, smem, st>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

// dense export of one Hermitian field: 0 Rvv, 1 Rvv_inv, 2 Ryy -> [S][K][M][M] c128
__global__ void amvdr_export_kernel(const double *state, double2 *out, int S, int K, int M, int field) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * K * M * M;
  if (g >= total) return;
  const int j = (int)(g % M), i = (int)((g / M) % M), k = (int)((g / ((long long)M * M)) % K), s = (int)(g / ((long long)M * M * K));
  const int NE = 3 * M * M + 5;
  const double *b = state + (long long)s * NE * K + k + (long long)field * M * M * K;
  double re, im = 0.0;
  if (i == j) {
    re = b[(long long)i * K];
  } else {
    const int lo = min(i, j), hi = max(i, j);
    const int e = M + 2 * (lo * (M - 1) - (lo * (lo - 1)) / 2 + (hi - lo - 1));
    re = b[(long long)e * K];
    im = b[(long long)(e + 1) * K];
    if (i > j) im = -im;
  }
  out[g] = make_double2(re, im);
}

}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_amvdr_state_bytes(const ds_amvdr_params *p) {
  if (!p) return 0;
  const int M = p->n_mics, K = p->n_fft / 2 + 1;
  return (size_t)p->n_streams * (3 * M * M + 5) * K * sizeof(double);
}

void ds_amvdr_default_params(ds_amvdr_params *p, int n_fft, int n_streams, int n_mics, int n_frames) {
  if (!p) return;
  p->n_fft = n_fft; p->n_streams = n_streams; p->n_mics = n_mics; p->n_frames = n_frames;
  p->frm_cnt = 0; p->ell = 1; p->mcra_L = 15; p->method = 2;
  p->alpha_y = 0.8; p->alpha_v = 0.9998; p->diag = 1e-6; p->vad_thr = 0.4;
  p->mcra_alpha_d = 0.95; p->mcra_alpha_s = 0.8; p->mcra_delta_s = 5.0; p->mcra_alpha_p = 0.2;
  p->mcra_p_min = 1e-3; p->mcra_p_max = 0.999;
}

int ds_amvdr_run(const ds_amvdr_params *p, void *state, const void *a, const void *X, void *Yout, void *H_last,
                 double *p_out, void *stream) {
  DS_CHECK_ARG(p && state && a && X && Yout, "ds_amvdr_run: null argument");
  DS_CHECK_ARG(p->n_streams >= 1 && p->n_frames >= 1 && p->n_fft >= 4 && p->mcra_L >= 1, "ds_amvdr_run: bad shape");
  DS_CHECK_ARG(p->method >= 0 && p->method <= 3, "ds_amvdr_run: method must be 0..3 (src, DS, MVDR, TFGSC)");
  AmvdrArgs g;
  g.state = (double *)state; g.a = (const double2 *)a; g.X = (const float2 *)X; g.Yout = (float2 *)Yout;
  g.H_last = (double2 *)H_last; g.p_out = p_out;
  g.S = p->n_streams; g.K = p->n_fft / 2 + 1; g.T = p->n_frames; g.frm_cnt = p->frm_cnt; g.ell = p->ell; g.method = p->method;
  g.alpha_y = p->alpha_y; g.alpha_v = p->alpha_v; g.diag = p->diag; g.vad_thr = p->vad_thr;
  g.mc.alpha_d = p->mcra_alpha_d; g.mc.alpha_s = p->mcra_alpha_s; g.mc.delta_s = p->mcra_delta_s;
  g.mc.alpha_p = p->mcra_alpha_p; g.mc.p_min = p->mcra_p_min; g.mc.p_max = p->mcra_p_max; g.mc.L = p->mcra_L;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->n_mics) {
    case 2: return launch_amvdr_m<2>(g, st);
    case 3: return launch_amvdr_m<3>(g, st);
    case 4: return launch_amvdr_m<4>(g, st);
    case 5: return launch_amvdr_m<5>(g, st);
    case 6: return launch_amvdr_m<6>(g, st);
    case 7: return launch_amvdr_m<7>(g, st);
    case 8: return launch_amvdr_m<8>(g, st);
  }
  set_error("ds_amvdr_run: n_mics %d outside the compiled range 2..8", p->n_mics);
  return DS_EUNSUPPORTED;
}

int ds_amvdr_export(const ds_amvdr_params *p, const void *state, int field, void *out, void *stream) {
  DS_CHECK_ARG(p && state && out, "ds_amvdr_export: null argument");
  DS_CHECK_ARG(field >= 0 && field <= 2, "ds_amvdr_export: field must be 0 (Rvv), 1 (Rvv_inv) or 2 (Ryy)");
  const int M = p->n_mics, K = p->n_fft / 2 + 1, S = p->n_streams;
  const long long total = (long long)S * K * M * M;
  amvdr_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const double *)state, (double2 *)out, S, K, M, field);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
