// Test-only host shim around distantspeech_b200/csrc/eig_core.cuh: lets the CPU test-suite check the
// Jacobi / Cholesky / reduction routines that the eig.cu kernels run per thread against LAPACK.
// Not part of the product (the product calls these routines from CUDA kernels only).
#include <vector>
#include <string.h>
#include "../../distantspeech_b200/csrc/eig_core.cuh"

using namespace ds;

// in: M x M complex128 (re, im interleaved, row-major); out: M complex128 = principal eigenvector
// with the phase of component 0 removed (steering(), beamformer.py:10-31).  Returns the sweep count.
extern "C" int eig_host_steering(int M, const double *in, double *out) {
  std::vector<double> buf(4 * M * M);
  CMatRef A{buf.data(), buf.data() + M * M, M, 1}, V{buf.data() + 2 * M * M, buf.data() + 3 * M * M, M, 1};
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      A.r(i, j) = in[2 * (i * M + j)]; A.c(i, j) = in[2 * (i * M + j) + 1];
      V.r(i, j) = i == j ? 1.0 : 0.0; V.c(i, j) = 0.0;
    }
  herm_from_lower(A);
  const int sw = jacobi_hermitian(A, V);
  const int col = argmax_diag(A);
  const double v0r = V.r(0, col), v0i = V.c(0, col), n = sqrt(v0r * v0r + v0i * v0i);
  const double er = n > 0 ? v0r / n : 1.0, ei = n > 0 ? v0i / n : 0.0;
  for (int i = 0; i < M; ++i) {              // v / e^{i angle(v0)} = v conj(e)
    out[2 * i] = V.r(i, col) * er + V.c(i, col) * ei;
    out[2 * i + 1] = V.c(i, col) * er - V.r(i, col) * ei;
  }
  return sw;
}

// a, b: M x M complex128; out: last generalised eigenvector (b-normalised, first component of the
// standard-form eigenvector real and non-negative).  Returns 0, or 1 if b is not positive definite.
extern "C" int eig_host_gev(int M, const double *a, const double *b, double *out) {
  std::vector<double> buf(6 * M * M);
  CMatRef A{buf.data(), buf.data() + M * M, M, 1}, B{buf.data() + 2 * M * M, buf.data() + 3 * M * M, M, 1},
      V{buf.data() + 4 * M * M, buf.data() + 5 * M * M, M, 1};
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      A.r(i, j) = a[2 * (i * M + j)]; A.c(i, j) = a[2 * (i * M + j) + 1];
      B.r(i, j) = b[2 * (i * M + j)]; B.c(i, j) = b[2 * (i * M + j) + 1];
      V.r(i, j) = i == j ? 1.0 : 0.0; V.c(i, j) = 0.0;
    }
  herm_from_lower(A);
  if (!cholesky_lower(B)) return 1;
  reduce_to_standard(A, B);
  herm_from_lower(A);
  jacobi_hermitian(A, V);
  const int col = argmax_diag(A);
  const double v0r = V.r(0, col), v0i = V.c(0, col), n = sqrt(v0r * v0r + v0i * v0i);
  const double er = n > 0 ? v0r / n : 1.0, ei = n > 0 ? v0i / n : 0.0;
  for (int i = 0; i < M; ++i) {
    const double xr = V.r(i, col) * er + V.c(i, col) * ei, xi = V.c(i, col) * er - V.r(i, col) * ei;
    V.r(i, col) = xr; V.c(i, col) = xi;
  }
  back_substitute_LH(B, V, col, out, out + 1, 2);
  return 0;
}
