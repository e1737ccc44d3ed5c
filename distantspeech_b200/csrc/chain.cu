// chain.cu -- config-4 composition: Transform.stft -> McSppBase.estimation ->
// compute_mvdr_weight -> compute_omlsa_weight -> (w^H y) G -> Transform.istft
// (example/mcsppbase.ipynb cell 3, example/mvdr.ipynb cell 4).
//
// Three kernels per call: analysis writes the complex64 spectrum X to the
// caller's workspace, the per-bin estimator/beamformer reads it once and writes
// the single-channel spectrum Y, synthesis overlap-adds Y to the waveform.
#include "common.cuh"

namespace ds {
int mcspp_run_impl(const ds_mcspp_params *p, void *state, const void *a0, const void *X, int x_is_c128, void *Yout,
                   int apply_gain, const ds_mcspp_taps *taps, cudaStream_t st);

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct ChainLayout {
  size_t est_off, hist_off, tail_off, total;      // state
  size_t X_off, Y_off, ws_total;                  // workspace
  int T, K, ov;
};

static ChainLayout chain_layout(const ds_chain_params *p) {
  ChainLayout L;
  const int S = p->est.n_streams, M = p->est.n_mics, N = p->est.n_fft;
  L.K = N / 2 + 1; L.ov = N - p->hop; L.T = p->n_samples / p->hop;
  L.est_off = 0;
  L.hist_off = align256(ds_mcspp_state_bytes(&p->est));
  L.tail_off = L.hist_off + align256((size_t)S * M * L.ov * sizeof(float));
  L.total = L.tail_off + align256((size_t)S * L.ov * sizeof(float));
  L.X_off = 0;
  L.Y_off = align256((size_t)S * L.T * M * L.K * sizeof(float2));
  L.ws_total = L.Y_off + align256((size_t)S * L.T * L.K * sizeof(float2));
  return L;
}
}  // namespace ds

using namespace ds;

extern "C" {

size_t ds_chain_state_bytes(const ds_chain_params *p) { return p ? chain_layout(p).total : 0; }
size_t ds_chain_workspace_bytes(const ds_chain_params *p) { return p ? chain_layout(p).ws_total : 0; }

static int chain_run_impl(const ds_chain_params *p, const double *window, const void *a0, void *state, void *workspace,
                          const void *x, int x_pcm16, void *y, int y_pcm16, void *stream, cudaEvent_t *ev);

int ds_chain_run(const ds_chain_params *p, const double *window, const void *a0, void *state, void *workspace,
                 const float *x, float *y, void *stream) {
  return chain_run_impl(p, window, a0, state, workspace, x, 0, y, 0, stream, nullptr);
}

int ds_chain_run_io(const ds_chain_params *p, const double *window, const void *a0, void *state, void *workspace,
                    const void *x, int x_is_pcm16, void *y, int y_is_pcm16, void *stream) {
  return chain_run_impl(p, window, a0, state, workspace, x, x_is_pcm16, y, y_is_pcm16, stream, nullptr);
}

int ds_chain_run_profiled(const ds_chain_params *p, const double *window, const void *a0, void *state, void *workspace,
                          const float *x, float *y, void *stream, float *phase_ms_h) {
  DS_CHECK_ARG(phase_ms_h, "ds_chain_run_profiled: null argument");
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) DS_CUDA(cudaEventCreate(&ev[i]));
  int rc = chain_run_impl(p, window, a0, state, workspace, x, 0, y, 0, stream, ev);
  if (rc == DS_OK) {
    DS_CUDA(cudaEventSynchronize(ev[3]));
    for (int i = 0; i < 3; ++i) DS_CUDA(cudaEventElapsedTime(&phase_ms_h[i], ev[i], ev[i + 1]));
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

static int chain_run_impl(const ds_chain_params *p, const double *window, const void *a0, void *state, void *workspace,
                          const void *x, int x_pcm16, void *y, int y_pcm16, void *stream, cudaEvent_t *ev) {
  DS_CHECK_ARG(p && window && a0 && state && workspace && x && y, "ds_chain_run: null argument");
  DS_CHECK_ARG(p->hop >= 1 && p->hop <= p->est.n_fft, "ds_chain_run: bad hop");
  DS_CHECK_ARG(p->n_samples >= p->hop && p->n_samples % p->hop == 0, "ds_chain_run: n_samples must be a positive multiple of hop");
  const ChainLayout L = chain_layout(p);
  unsigned char *st8 = (unsigned char *)state, *ws8 = (unsigned char *)workspace;
  ds_stft_params sp;
  sp.n_fft = p->est.n_fft; sp.hop = p->hop; sp.n_streams = p->est.n_streams; sp.n_ch = p->est.n_mics;
  sp.n_samples = p->n_samples; sp.mode = DS_STFT_STREAMING; sp.fft_fp64 = p->fft_fp64; sp.out_c128 = 0;
  cudaStream_t cst = (cudaStream_t)stream;
  if (ev) DS_CUDA(cudaEventRecord(ev[0], cst));
  // int16 PCM in / out: load_audio's and save_audio's scaling (beamformer/utils.py:184-185, :193) are fused into the
  // analysis kernel's first register-fed pass and the synthesis kernel's store -- no staging buffer, half the bytes
  int rc = x_pcm16 ? ds_stft_pcm16_run(&sp, window, (float *)(st8 + L.hist_off), (const int16_t *)x, ws8 + L.X_off, stream)
                   : ds_stft_run(&sp, window, (float *)(st8 + L.hist_off), (const float *)x, ws8 + L.X_off, stream);
  if (rc != DS_OK) return rc;
  if (ev) DS_CUDA(cudaEventRecord(ev[1], cst));
  ds_mcspp_params ep = p->est;
  ep.n_frames = L.T;
  rc = mcspp_run_impl(&ep, st8 + L.est_off, a0, ws8 + L.X_off, 0, ws8 + L.Y_off, p->apply_gain, nullptr, (cudaStream_t)stream);
  if (rc != DS_OK) return rc;
  if (ev) DS_CUDA(cudaEventRecord(ev[2], cst));
  ds_istft_params ip;
  ip.n_fft = p->est.n_fft; ip.hop = p->hop; ip.n_streams = p->est.n_streams; ip.n_ch = 1; ip.n_frames = L.T;
  ip.mode = DS_STFT_STREAMING; ip.fft_fp64 = p->fft_fp64; ip.in_c128 = 0; ip.scale = p->scale;
  rc = y_pcm16 ? ds_istft_pcm16_run(&ip, window, (float *)(st8 + L.tail_off), ws8 + L.Y_off, (int16_t *)y, stream)
               : ds_istft_run(&ip, window, (float *)(st8 + L.tail_off), ws8 + L.Y_off, (float *)y, stream);
  if (rc != DS_OK) return rc;
  if (ev) DS_CUDA(cudaEventRecord(ev[3], cst));
  return DS_OK;
}

}  // extern "C"
