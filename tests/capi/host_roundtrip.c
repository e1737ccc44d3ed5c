/* A host program in plain C that uses libds_b200.so the way a non-Python maintainer would:
 * dlopen, cudaMalloc'd device buffers, POD parameter structs, a CUDA stream -- no torch anywhere.
 *
 *   x (2 streams x 3 channels x 8192 samples) -> ds_stft_run (streaming) -> ds_istft_run (streaming)
 *
 * With the sqrt-Hann window at hop = n_fft/2 the analysis/synthesis pair is the identity delayed by
 * n_fft - hop samples (Transform.stft / Transform.istft, transform.py:430-481), which is what it checks,
 * together with the error path (a bad n_fft must return DS_EUNSUPPORTED and set ds_last_error()).
 * Exit code 0 = pass.  Build: gcc host_roundtrip.c -I<repo>/include -I$CUDA/include -L$CUDA/lib64 -lcudart -ldl -lm */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ds_b200.h"

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } \
  } while (0)

typedef int (*stft_fn)(const ds_stft_params *, const double *, float *, const float *, void *, void *);
typedef int (*istft_fn)(const ds_istft_params *, const double *, float *, const void *, float *, void *);
typedef int (*nframes_fn)(const ds_stft_params *);
typedef const char *(*err_fn)(void);
typedef int (*init_fn)(void);

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s /path/to/libds_b200.so\n", argv[0]); return 2; }
  void *h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  init_fn ds_init_p = (init_fn)dlsym(h, "ds_init");
  stft_fn stft = (stft_fn)dlsym(h, "ds_stft_run");
  istft_fn istft = (istft_fn)dlsym(h, "ds_istft_run");
  nframes_fn nframes = (nframes_fn)dlsym(h, "ds_stft_num_frames");
  err_fn last_error = (err_fn)dlsym(h, "ds_last_error");
  if (!ds_init_p || !stft || !istft || !nframes || !last_error) { fprintf(stderr, "missing symbol\n"); return 2; }
  if (ds_init_p() != DS_OK) { fprintf(stderr, "ds_init: %s\n", last_error()); return 2; }

  enum { S = 2, M = 3, N = 8192, NFFT = 512, HOP = 256, K = NFFT / 2 + 1, OV = NFFT - HOP };
  const double PI = 3.14159265358979323846;
  double window[NFFT], w0 = 0.0;
  for (int n = 0; n < NFFT; ++n) { window[n] = sqrt(0.5 - 0.5 * cos(2.0 * PI * n / NFFT)); w0 += window[n] * window[n]; }
  float *x = (float *)malloc(sizeof(float) * S * M * N), *y = (float *)malloc(sizeof(float) * S * M * N);
  unsigned r = 12345u;
  for (int i = 0; i < S * M * N; ++i) { r = r * 1664525u + 1013904223u; x[i] = ((float)(r >> 8) / 8388608.0f - 1.0f) * 0.25f; }

  ds_stft_params sp = {NFFT, HOP, S, M, N, DS_STFT_STREAMING, 0, 0};
  const int T = nframes(&sp);
  if (T != N / HOP) { fprintf(stderr, "unexpected frame count %d\n", T); return 1; }
  double *d_win; float *d_x, *d_y, *d_hist, *d_tail; void *d_X;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(cudaMalloc((void **)&d_win, sizeof(window)));
  CK(cudaMalloc((void **)&d_x, sizeof(float) * S * M * N));
  CK(cudaMalloc((void **)&d_y, sizeof(float) * S * M * N));
  CK(cudaMalloc((void **)&d_hist, sizeof(float) * S * M * OV));
  CK(cudaMalloc((void **)&d_tail, sizeof(float) * S * M * OV));
  CK(cudaMalloc(&d_X, sizeof(float) * 2 * (size_t)S * T * M * K));
  CK(cudaMemcpy(d_win, window, sizeof(window), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, x, sizeof(float) * S * M * N, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_hist, 0, sizeof(float) * S * M * OV));      /* zero state = fresh Transform object */
  CK(cudaMemset(d_tail, 0, sizeof(float) * S * M * OV));

  int rc = stft(&sp, d_win, d_hist, d_x, d_X, (void *)st);
  if (rc != DS_OK) { fprintf(stderr, "ds_stft_run: %d %s\n", rc, last_error()); return 1; }
  ds_istft_params ip = {NFFT, HOP, S, M, T, DS_STFT_STREAMING, 0, 0, (double)HOP / w0};
  rc = istft(&ip, d_win, d_tail, d_X, d_y, (void *)st);
  if (rc != DS_OK) { fprintf(stderr, "ds_istft_run: %d %s\n", rc, last_error()); return 1; }
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpy(y, d_y, sizeof(float) * S * M * N, cudaMemcpyDeviceToHost));

  double worst = 0.0;
  for (int sc = 0; sc < S * M; ++sc)
    for (int n = OV; n < N; ++n) {
      const double d = fabs((double)y[sc * N + n] - (double)x[sc * N + n - OV]);
      if (d > worst) worst = d;
    }
  printf("round trip max |y[n] - x[n - %d]| = %.3e\n", OV, worst);
  if (!(worst < 3e-6)) return 1;

  sp.n_fft = 500;                                              /* not a power of two */
  rc = stft(&sp, d_win, d_hist, d_x, d_X, (void *)st);
  printf("bad n_fft -> rc %d, message: %s\n", rc, last_error());
  if (rc != DS_EUNSUPPORTED || strlen(last_error()) == 0) return 1;
  printf("PASS\n");
  return 0;
}
