// fp64_probe.cu -- how many warps per scheduler / independent chains per thread the B200 fp64 pipe needs to reach its peak.
// One CTA per SM, W warps per scheduler (4 W warps), ILP independent DFMA chains per thread.  Prints TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(int iters, double seed, double *out) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x + i;
  const double m = 1.0 - 1e-9, c = 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 12345.678) out[0] = r;
}
template <int ILP> void run(int W, double *d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000 / ILP, threads = 128 * W;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<sms, threads>>>(iters, 1.0, d);
  cudaEventRecord(e0); k<ILP><<<sms, threads>>>(iters, 1.0, d); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 16 * ILP * (double)iters * sms * threads;
  printf("W=%2d ILP=%d  %.2f TFLOP/s\n", W, ILP, flop / ms / 1e9);
}
int main() {
  double *d; cudaMalloc(&d, 8);
  for (int W : {1, 2, 3, 4, 6, 8}) { run<1>(W, d); run<2>(W, d); run<4>(W, d); run<8>(W, d); }
  return 0;
}
