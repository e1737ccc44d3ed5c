// wpe.cu -- adaptive (RLS) weighted-prediction-error dereverberation, Wpe.update
// (dereverberation/awpe.py:128-191), per frequency bin:
//   X   = [x_c(t - D - l)]  c-major, l = 0..L-1                   buffer_input :77-100, reshape :154
//   err = d - W^H X                                                 :156-159
//   var = .98 var + .02 mean_c |d_c|^2                              :161-163
//   kn  = P X / (lambda var + X^H P X)   (complex denominator)      :172-178
//   P   = (P - kn (X^H P)) / lambda                                 :181-183
//   W_c += conj(err_c) kn                                           :186-187
// One thread per (stream, bin), frames sequential; the recursive state (W, P, the last D + L - 1 input frames, var)
// lives in a float64 blob [S][NE][K] (bin index innermost: coalesced) and is updated in place -- with C L up to 16 taps
// the matrices do not fit the register file, so P and W are walked through L1/L2 every frame; the frame's vectors
// (X, P X, kn, X^H P) stay in registers / local memory.  A delay of D hops in the time domain (DelaySamples :72,148)
// is a delay of D frames of the streaming STFT, so the kernel is fed the undelayed spectrum only.
#include "common.cuh"

namespace ds {

constexpr int WPE_MAXCL = 16;

struct WpeArgs {
  double *state;          // [S][NE][K]
  const void *X;          // [S][T][C][K] c64 / c128
  double2 *Err;           // [S][T][C][K] c128 prior error (the dereverberated spectrum)
  int S, K, T, C, L, D, x_c128;
  double lambda, alpha_var;
};

__host__ __device__ inline int wpe_hist_frames(int D, int L) { return D + L - 1; }
// element offsets: W re[C][CL] im[C][CL] | P re[CL][CL] im[CL][CL] | hist re[H][C] im[H][C] (frame t-1 first) | var
__host__ __device__ inline int wpe_state_elems(int C, int L, int D) {
  const int CL = C * L;
  return 2 * C * CL + 2 * CL * CL + 2 * wpe_hist_frames(D, L) * C + 1;
}

__global__ void __launch_bounds__(64) wpe_kernel(WpeArgs a) {
  const int K = a.K, C = a.C, L = a.L, D = a.D, CL = C * L, H = wpe_hist_frames(D, L);
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.S * K) return;
  const int s = (int)(g / K), k = (int)(g % K);
  const int NE = wpe_state_elems(C, L, D);
  double *st = a.state + (long long)s * NE * K + k;
  double *Wr = st, *Wi = st + (long long)C * CL * K;
  double *Pr = st + (long long)2 * C * CL * K, *Pi = Pr + (long long)CL * CL * K;
  double *Hr = st + (long long)(2 * C * CL + 2 * CL * CL) * K, *Hi = Hr + (long long)H * C * K;
  double *varp = st + (long long)(NE - 1) * K;
  double var = *varp;
  const double inv_lambda = 1.0 / a.lambda;
#define E(p, idx) (p)[(long long)(idx) * K]
  // spectrum of frame tt of this call (tt >= 0) or of the carried history (tt < 0: frame -1 is history slot 0)
  auto frame = [&](int tt, int c, double &re, double &im) {
    if (tt >= 0) {
      const long long o = (((long long)s * a.T + tt) * C + c) * K + k;
      if (a.x_c128) { const double2 v = reinterpret_cast<const double2 *>(a.X)[o]; re = v.x; im = v.y; }
      else { const float2 v = reinterpret_cast<const float2 *>(a.X)[o]; re = (double)v.x; im = (double)v.y; }
    } else {
      const int slot = -tt - 1;
      re = E(Hr, slot * C + c); im = E(Hi, slot * C + c);
    }
  };
  for (int t = 0; t < a.T; ++t) {
    double xr[WPE_MAXCL], xi[WPE_MAXCL], nr[WPE_MAXCL], ni[WPE_MAXCL], kr[WPE_MAXCL], ki[WPE_MAXCL];
    for (int c = 0; c < C; ++c)
      for (int l = 0; l < L; ++l) frame(t - D - l, c, xr[c * L + l], xi[c * L + l]);
    // prior error and variance of the current frame
    double pw = 0.0;
    double er[8], ei[8];
    for (int m = 0; m < C; ++m) {
      double dr, di;
      frame(t, m, dr, di);
      pw += dr * dr + di * di;
      double outr = 0.0, outi = 0.0;
      for (int i = 0; i < CL; ++i) {                         // conj(W[m][i]) X[i]
        const double wr = E(Wr, m * CL + i), wi = E(Wi, m * CL + i);
        outr = fma(wr, xr[i], fma(wi, xi[i], outr));
        outi = fma(wr, xi[i], fma(-wi, xr[i], outi));
      }
      er[m] = dr - outr; ei[m] = di - outi;
      a.Err[(((long long)s * a.T + t) * C + m) * K + k] = make_double2(er[m], ei[m]);
    }
    var = a.alpha_var * var + (1.0 - a.alpha_var) * (pw / (double)C);
    // num = P X ; den = lambda var + X^H num
    double denr = a.lambda * var, deni = 0.0;
    for (int i = 0; i < CL; ++i) {
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < CL; ++j) {
        const double pr = E(Pr, i * CL + j), pi = E(Pi, i * CL + j);
        sr = fma(pr, xr[j], fma(-pi, xi[j], sr));
        si = fma(pr, xi[j], fma(pi, xr[j], si));
      }
      nr[i] = sr; ni[i] = si;
      denr = fma(xr[i], sr, fma(xi[i], si, denr));           // conj(X_i) num_i
      deni = fma(xr[i], si, fma(-xi[i], sr, deni));
    }
    const double d2 = 1.0 / fma(denr, denr, deni * deni);
    for (int i = 0; i < CL; ++i) {                           // kn = num / den
      kr[i] = fma(nr[i], denr, ni[i] * deni) * d2;
      ki[i] = fma(ni[i], denr, -nr[i] * deni) * d2;
    }
    // v = X^H P (row vector), reusing nr / ni ;  P = (P - kn v) / lambda
    for (int j = 0; j < CL; ++j) {
      double sr = 0.0, si = 0.0;
      for (int i = 0; i < CL; ++i) {
        const double pr = E(Pr, i * CL + j), pi = E(Pi, i * CL + j);
        sr = fma(xr[i], pr, fma(xi[i], pi, sr));
        si = fma(xr[i], pi, fma(-xi[i], pr, si));
      }
      nr[j] = sr; ni[j] = si;
    }
    for (int i = 0; i < CL; ++i)
      for (int j = 0; j < CL; ++j) {
        E(Pr, i * CL + j) = (E(Pr, i * CL + j) - fma(kr[i], nr[j], -ki[i] * ni[j])) * inv_lambda;
        E(Pi, i * CL + j) = (E(Pi, i * CL + j) - fma(kr[i], ni[j], ki[i] * nr[j])) * inv_lambda;
      }
    for (int m = 0; m < C; ++m)                              // W[m] += conj(err_m) kn
      for (int i = 0; i < CL; ++i) {
        E(Wr, m * CL + i) = fma(er[m], kr[i], fma(ei[m], ki[i], E(Wr, m * CL + i)));
        E(Wi, m * CL + i) = fma(er[m], ki[i], fma(-ei[m], kr[i], E(Wi, m * CL + i)));
      }
  }
  // carry the last H frames (frame T-1 -> slot 0, ...); older slots shift when the call was shorter than H frames
  if (H > 0) {
    double nr_[WPE_MAXCL * 2], ni_[WPE_MAXCL * 2];           // H * C <= (D + L - 1) * C values, staged to allow in-place shift
    for (int slot = 0; slot < H; ++slot)
      for (int c = 0; c < C; ++c) {
        double re, im;
        frame(a.T - 1 - slot, c, re, im);
        if (slot * C + c < WPE_MAXCL * 2) { nr_[slot * C + c] = re; ni_[slot * C + c] = im; }
      }
    for (int e = 0; e < H * C && e < WPE_MAXCL * 2; ++e) { E(Hr, e) = nr_[e]; E(Hi, e) = ni_[e]; }
  }
  *varp = var;
#undef E
}

}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_wpe_state_bytes(int n_streams, int n_bins, int n_ch, int filter_len, int delay) {
  if (n_streams < 1 || n_bins < 1 || n_ch < 1 || filter_len < 1 || delay < 0) return 0;
  return (size_t)n_streams * wpe_state_elems(n_ch, filter_len, delay) * n_bins * sizeof(double);
}

int ds_wpe_run(int n_streams, int n_bins, int n_frames, int n_ch, int filter_len, int delay, double forgetting_factor,
               double alpha_var, void *state, const void *X, int x_is_c128, void *Err, void *stream) {
  DS_CHECK_ARG(state && X && Err, "ds_wpe_run: null argument");
  DS_CHECK_ARG(n_streams >= 1 && n_bins >= 1 && n_frames >= 1, "ds_wpe_run: bad shape");
  DS_CHECK_ARG(n_ch >= 1 && n_ch <= 8 && filter_len >= 1 && n_ch * filter_len <= WPE_MAXCL,
               "ds_wpe_run: channels (<= 8) x filter_len must not exceed %d taps", WPE_MAXCL);
  DS_CHECK_ARG(delay >= 0 && (delay + filter_len - 1) * n_ch <= 2 * WPE_MAXCL, "ds_wpe_run: (delay + filter_len - 1) x channels must not exceed %d", 2 * WPE_MAXCL);
  DS_CHECK_ARG(forgetting_factor > 0.0 && forgetting_factor <= 1.0, "ds_wpe_run: forgetting factor outside (0, 1]");
  WpeArgs a;
  a.state = (double *)state; a.X = X; a.Err = (double2 *)Err; a.S = n_streams; a.K = n_bins; a.T = n_frames; a.C = n_ch;
  a.L = filter_len; a.D = delay; a.x_c128 = x_is_c128; a.lambda = forgetting_factor; a.alpha_var = alpha_var;
  const long long items = (long long)n_streams * n_bins;
  wpe_kernel<<<(unsigned)((items + 63) / 64), 64, 0, (cudaStream_t)stream>>>(a);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // extern "C"
