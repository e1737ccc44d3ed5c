// postfilter.cu -- postfilter gains, one thread per (stream, bin), frames sequential.
//
// Reference behaviour restated (file:line relative to the reference tree):
//   NsOmlsaMulti.estimation      noise_estimation/omlsa_multi.py:73-156   (Cohen/Gannot multichannel OMLSA)
//   PostFilter.update_CSD_PSD    postfilter/postfilter.py:19-43
//   PostFilter.getweights        postfilter/postfilter.py:45-84           (Zelinski / McCowan)
#include "common.cuh"
#include "perbin.cuh"

namespace ds {

constexpr int PF_MAXM = 8;

// ---------------------------------------------------------------------------
// NsOmlsaMulti
// ---------------------------------------------------------------------------
struct OmlsaArgs {
  double *state;       // [S][NE][K]
  const double *y;     // [S][T][K]      beamformer output power
  const double *u;     // [S][T][M-1][K] reference (blocking matrix) powers, or [S][M-1][K] when u_const
  int u_const;
  double *G_out, *lam_out, *p_out;   // [S][T][K] or null
  int S, K, T, M, first_frame, frm_cnt, ell, cal_weights;
  double alpha_d, alpha_s, alpha_xi, beta, Gmin, q_min, q_max;
  McraConst mc;
};
// state elements: mcra[5] x M (fixed first, then M-1 refs), zeta_Y, zeta_U[M-1], lambda_d, gamma, G_H1, p, G, q_hat, xi_hat
__host__ __device__ inline int omlsa_state_elems(int M) { return 5 * M + 1 + (M - 1) + 7; }

__global__ void omlsa_multi_kernel(OmlsaArgs a) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * a.K) return;
  const int s = (int)(g / a.K), k = (int)(g % a.K), K = a.K, M = a.M;
  const int NE = omlsa_state_elems(M);
  double *st = a.state + (long long)s * NE * K + k;
#define ST(e) st[(long long)(e) * K]
  double mc[PF_MAXM][5];
  for (int c = 0; c < M; ++c)
    for (int e = 0; e < 5; ++e) mc[c][e] = ST(5 * c + e);
  int o = 5 * M;
  double zeta_Y = ST(o);
  double zeta_U[PF_MAXM];
  for (int c = 0; c < M - 1; ++c) zeta_U[c] = ST(o + 1 + c);
  o += M;
  double lambda_d = ST(o), gamma = ST(o + 1), G_H1 = ST(o + 2), p = ST(o + 3), G = ST(o + 4), q_hat = ST(o + 5), xi_hat = ST(o + 6);
  int frm = a.frm_cnt, ell = a.ell % a.mc.L, first = a.first_frame;   // ell is kept modulo L: no integer division per frame (mcra.py:52-56)
  const double w0 = 0.25, w1 = 0.5, w2 = 0.25;
  for (int t = 0; t < a.T; ++t) {
    const double *yt = a.y + ((long long)s * a.T + t) * K;
    const double y0 = yt[k], ym = (k > 0) ? yt[k - 1] : 0.0, yp = (k < K - 1) ? yt[k + 1] : 0.0;
    const bool reset = (frm > 0) && (ell == 0);
    mcra_step(mc[0][0], mc[0][1], mc[0][2], mc[0][3], mc[0][4], ym, y0, yp, k, K, frm, reset, a.mc);   // :82
    const double MU_Y = mc[0][4];
    double u0[PF_MAXM], um[PF_MAXM], up[PF_MAXM], MU_U[PF_MAXM];
    for (int c = 0; c < M - 1; ++c) {
      const double *ut = a.u + (((long long)s * (a.u_const ? 1 : a.T) + (a.u_const ? 0 : t)) * (M - 1) + c) * K;
      u0[c] = ut[k]; um[c] = (k > 0) ? ut[k - 1] : 0.0; up[c] = (k < K - 1) ? ut[k + 1] : 0.0;
      mcra_step(mc[c + 1][0], mc[c + 1][1], mc[c + 1][2], mc[c + 1][3], mc[c + 1][4], um[c], u0[c], up[c], k, K, frm, reset, a.mc);
      MU_U[c] = mc[c + 1][4];
    }
    if (reset) ell = 0;
    ++ell; ++frm;
    if (ell == a.mc.L) ell = 0;
    if (first) {                                                          // :87-93
      first = 0;
      lambda_d = y0; zeta_Y = y0;
      for (int c = 0; c < M - 1; ++c) zeta_U[c] = u0[c];
    } else {
      zeta_Y = a.alpha_s * zeta_Y + (1 - a.alpha_s) * (ym * w2 + y0 * w1 + yp * w0);          // Eq 21 (:98)
      double maxdiff = -1e300;
      for (int c = 0; c < M - 1; ++c) {
        zeta_U[c] = a.alpha_s * zeta_U[c] + (1 - a.alpha_s) * (um[c] * w2 + u0[c] * w1 + up[c] * w0);
        maxdiff = fmax(maxdiff, zeta_U[c] - MU_U[c]);
      }
      const double eps = 0.01;
      double Omega = fmax(zeta_Y - MU_Y, 1e-6) / (fmax(maxdiff, eps * MU_Y) + 1e-6);            // Eq 6 (:107-111)
      Omega = fmin(fmax(Omega, 0.1), 100.0);
      const double Bmin = 1.66;
      const double gamma_s = fmin(y0 / (MU_Y * Bmin + 1e-6), 100.0);                            // Eq 27 (:115)
      const double gamma_high = 0.1 * 100.0, gamma_low = 1.0, Omega_high = 3.0, Omega_low = 0.3;
      if (gamma_s < gamma_low || Omega < Omega_low) q_hat = 1.0;                                 // Eq 29 (:122-130)
      else q_hat = fmax((gamma_high - gamma_s) / (gamma_high - gamma_low), (Omega_high - Omega) / (Omega_high - Omega_low));
      q_hat = fmin(fmax(q_hat, a.q_min), a.q_max);
      const double gamma_pre = gamma;
      gamma = y0 / fmax(lambda_d, 1e-10);                                                        // :134
      xi_hat = a.alpha_xi * (G_H1 * G_H1) * gamma_pre + (1 - a.alpha_xi) * fmax(gamma - 1.0, 0.0);   // Eq 30 (:137)
      const double nu = gamma * xi_hat / (1 + xi_hat);
      G_H1 = xi_hat / (1 + xi_hat);                                                              // :144
      p = 1.0 / (1.0 + q_hat / (1.0 - q_hat) * (1.0 + xi_hat) * exp(-1.0 * nu));                 // Eq 28 (:147)
      const double at = a.alpha_d + (1 - a.alpha_d) * p;                                         // :149, Base :56-60
      lambda_d = at * lambda_d + a.beta * (1 - at) * y0;
      if (a.cal_weights) {                                                                       // Eq 35 (:152-154)
        G = pow(G_H1, p) * pow(a.Gmin, 1.0 - p);
        G = fmax(fmin(G, 1.0), a.Gmin);
      }
    }
    const long long oo = ((long long)s * a.T + t) * K + k;
    if (a.G_out) a.G_out[oo] = G;
    if (a.lam_out) a.lam_out[oo] = lambda_d;
    if (a.p_out) a.p_out[oo] = p;
  }
  for (int c = 0; c < M; ++c)
    for (int e = 0; e < 5; ++e) ST(5 * c + e) = mc[c][e];
  o = 5 * M;
  ST(o) = zeta_Y;
  for (int c = 0; c < M - 1; ++c) ST(o + 1 + c) = zeta_U[c];
  o += M;
  ST(o) = lambda_d; ST(o + 1) = gamma; ST(o + 2) = G_H1; ST(o + 3) = p; ST(o + 4) = G; ST(o + 5) = q_hat; ST(o + 6) = xi_hat;
#undef ST
}

// ---------------------------------------------------------------------------
// Zelinski / McCowan postfilter
// ---------------------------------------------------------------------------
__global__ void zelinski_kernel(double *state, const double2 *__restrict__ Z, const double *__restrict__ Fvv, double *W,
                                int S, int K, int T, int M, double alpha, double fmax_coh) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  const int NS = M * (M - 1) / 2, NE = M + 2 * NS;
  double *st = state + (long long)s * NE * K + k;
  for (int t = 0; t < T; ++t) {
    const double2 *z = Z + ((long long)s * T + t) * M * K + k;
    double psum = 0.0;
    for (int i = 0; i < M; ++i) {                                          // :30-34
      const double2 zi = z[(long long)i * K];
      const double v = alpha * st[(long long)i * K] + (1 - alpha) * (zi.x * zi.x + zi.y * zi.y);
      st[(long long)i * K] = v;
      psum += v;
    }
    double acc = 0.0;
    int tt = 0;
    for (int i = 0; i < M - 1; ++i) {
      const double2 zi = z[(long long)i * K];
      for (int j = i + 1; j < M; ++j, ++tt) {
        const double2 zj = z[(long long)j * K];
        double *pr = st + (long long)(M + 2 * tt) * K, *pi = pr + K;
        const double cr = zi.x * zj.x + zi.y * zj.y, ci = zi.y * zj.x - zi.x * zj.y;      // z_i conj(z_j)
        *pr = alpha * (*pr) + (1 - alpha) * cr;                                             // :37-43
        *pi = alpha * (*pi) + (1 - alpha) * ci;
        const double F = fmin(Fvv[((long long)k * M + i) * M + j], fmax_coh);              // :67-68
        acc += (*pr - 0.5 * F * (st[(long long)i * K] + st[(long long)j * K])) / (1.0 - F); // Eq 22 (:69-71)
      }
    }
    const double Pss = (NS > 1) ? acc * 2.0 / (double)(M * M - M) : acc;                   // Eq 23 (:77-80)
    W[((long long)s * T + t) * K + k] = Pss / (psum / (double)M);                          // :82
  }
}

}  // namespace ds

using namespace ds;

extern "C" {

void ds_omlsa_multi_default_params(ds_omlsa_multi_params *p, int n_bins, int n_streams, int n_frames, int n_mics) {
  if (!p) return;
  p->n_bins = n_bins; p->n_streams = n_streams; p->n_frames = n_frames; p->n_mics = n_mics;
  p->first_frame = 1; p->frm_cnt = 0; p->ell = 1; p->mcra_L = 15; p->cal_weights = 0; p->u_const = 0;
  p->alpha_d = 0.85; p->alpha_s = 0.8; p->alpha_xi = 0.921; p->beta = 1.47; p->Gmin = pow(10.0, -12.0 / 10.0);
  p->q_min = 1e-6; p->q_max = 0.9999998;
  p->mcra_alpha_d = 0.95; p->mcra_alpha_s = 0.8; p->mcra_delta_s = 5.0; p->mcra_alpha_p = 0.2;
  p->mcra_p_min = 1e-3; p->mcra_p_max = 0.999;
}

size_t ds_omlsa_multi_state_bytes(const ds_omlsa_multi_params *p) {
  if (!p) return 0;
  return (size_t)p->n_streams * omlsa_state_elems(p->n_mics) * p->n_bins * sizeof(double);
}

int ds_omlsa_multi_run(const ds_omlsa_multi_params *p, void *state, const double *y, const double *u, double *G_out,
                       double *lambda_out, double *p_out, void *stream) {
  DS_CHECK_ARG(p && state && y && u, "ds_omlsa_multi_run: null argument");
  DS_CHECK_ARG(p->n_bins >= 3 && p->n_streams >= 1 && p->n_frames >= 1, "ds_omlsa_multi_run: bad shape");
  DS_CHECK_ARG(p->n_mics >= 2 && p->n_mics <= PF_MAXM, "ds_omlsa_multi_run: M must be 2..%d", PF_MAXM);
  OmlsaArgs a;
  a.state = (double *)state; a.y = y; a.u = u; a.G_out = G_out; a.lam_out = lambda_out; a.p_out = p_out;
  a.S = p->n_streams; a.K = p->n_bins; a.T = p->n_frames; a.M = p->n_mics; a.first_frame = p->first_frame;
  a.frm_cnt = p->frm_cnt; a.ell = p->ell; a.cal_weights = p->cal_weights; a.u_const = p->u_const;
  a.alpha_d = p->alpha_d; a.alpha_s = p->alpha_s; a.alpha_xi = p->alpha_xi; a.beta = p->beta; a.Gmin = p->Gmin;
  a.q_min = p->q_min; a.q_max = p->q_max;
  a.mc.alpha_d = p->mcra_alpha_d; a.mc.alpha_s = p->mcra_alpha_s; a.mc.delta_s = p->mcra_delta_s;
  a.mc.alpha_p = p->mcra_alpha_p; a.mc.p_min = p->mcra_p_min; a.mc.p_max = p->mcra_p_max; a.mc.L = p->mcra_L;
  const long long items = (long long)a.S * a.K;
  omlsa_multi_kernel<<<(unsigned)((items + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

size_t ds_zelinski_state_bytes(int n_streams, int n_mics, int n_bins) {
  return (size_t)n_streams * (n_mics + n_mics * (n_mics - 1)) * n_bins * sizeof(double);
}

int ds_zelinski_run(int n_streams, int n_frames, int n_mics, int n_bins, double alpha, double coh_max, void *state,
                    const void *Z, const double *Fvv, double *W, void *stream) {
  DS_CHECK_ARG(state && Z && Fvv && W, "ds_zelinski_run: null argument");
  DS_CHECK_ARG(n_streams >= 1 && n_frames >= 1 && n_mics >= 2 && n_bins >= 1, "ds_zelinski_run: bad shape");
  const long long items = (long long)n_streams * n_bins;
  zelinski_kernel<<<(unsigned)((items + 127) / 128), 128, 0, (cudaStream_t)stream>>>((double *)state, (const double2 *)Z, Fvv, W,
                                                                                      n_streams, n_bins, n_frames, n_mics, alpha, coh_max);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
