// eig_core.cuh -- small complex Hermitian eigenproblems, one matrix per thread.
//
// Plain C++ underneath (DS_HD expands to __host__ __device__ only under nvcc), so the same
// routines are unit-tested on the host against LAPACK (tests/test_eig_core_host.py); the
// product only ever calls them from the kernels in eig.cu.
//
// Reference behaviour being replaced:
//   steering()        beamformer/beamformer.py:10-31   np.linalg.eigh(XXs)[1][:, :, -1], phase referenced to sensor 0
//   get_gev_vector()  beamformer/beamformer.py:77-97   scipy.linalg.eigh(target, noise), last eigenvector
// Both LAPACK drivers read the LOWER triangle only; so do these routines.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define DS_HD __host__ __device__ __forceinline__
#else
#define DS_HD inline
#endif

namespace ds {

// M x M complex matrix of one thread: element (i, j) at re/im[(i * M + j) * st].  st is the
// number of threads that interleave their matrices in the same buffer (shared memory on the
// device: consecutive threads -> consecutive 8-byte words, conflict-free; 1 on the host).
struct CMatRef {
  double *re, *im;
  int M, st;
  DS_HD double &r(int i, int j) const { return re[(i * M + j) * st]; }
  DS_HD double &c(int i, int j) const { return im[(i * M + j) * st]; }
};

// Mirror the lower triangle into the upper one and drop the imaginary part of the diagonal:
// what zheevd / zhegvd (UPLO = 'L') take the input to be.
DS_HD void herm_from_lower(const CMatRef &A) {
  for (int i = 0; i < A.M; ++i) {
    A.c(i, i) = 0.0;
    for (int j = i + 1; j < A.M; ++j) { A.r(i, j) = A.r(j, i); A.c(i, j) = -A.c(j, i); }
  }
}

// Cyclic complex Jacobi: A <- G^H A G until the off-diagonal mass is below 1e-34 of the
// diagonal's, V <- V G (V must come in as the identity).  On return the diagonal of A holds
// the eigenvalues and the columns of V the orthonormal eigenvectors.  Each rotation first
// removes the phase of the pivot (a_pq = g e^{i phi}) and then applies the real symmetric
// Jacobi rotation with t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = (a_qq - a_pp) / 2g.
DS_HD int jacobi_hermitian(const CMatRef &A, const CMatRef &V, int max_sweeps = 16) {
  const int M = A.M;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int p = 0; p < M; ++p) {
      dg += A.r(p, p) * A.r(p, p);
      for (int q = p + 1; q < M; ++q) off += A.r(p, q) * A.r(p, q) + A.c(p, q) * A.c(p, q);
    }
    if (!(off > 1e-34 * dg)) break;            // also leaves on NaN
    for (int p = 0; p < M - 1; ++p)
      for (int q = p + 1; q < M; ++q) {
        const double ar = A.r(p, q), ai = A.c(p, q);
        const double g2 = ar * ar + ai * ai;
        if (g2 == 0.0) continue;
        const double g = sqrt(g2);
        const double er = ar / g, ei = ai / g;                       // e = e^{i phi}
        const double tau = (A.r(q, q) - A.r(p, p)) / (2.0 * g);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        const double app = A.r(p, p) - t * g, aqq = A.r(q, q) + t * g;
        // columns p, q of A and V:  x_p' = c x_p - s conj(e) x_q ;  x_q' = s x_p + c conj(e) x_q
        for (int pass = 0; pass < 2; ++pass) {
          const CMatRef &X = pass ? V : A;
          for (int i = 0; i < M; ++i) {
            const double xpr = X.r(i, p), xpi = X.c(i, p), xqr = X.r(i, q), xqi = X.c(i, q);
            const double wr = er * xqr + ei * xqi, wi = er * xqi - ei * xqr;   // conj(e) x_q
            X.r(i, p) = c * xpr - s * wr; X.c(i, p) = c * xpi - s * wi;
            X.r(i, q) = s * xpr + c * wr; X.c(i, q) = s * xpi + c * wi;
          }
        }
        // rows p, q of A:  y_p' = c y_p - s e y_q ;  y_q' = s y_p + c e y_q
        for (int j = 0; j < M; ++j) {
          const double ypr = A.r(p, j), ypi = A.c(p, j), yqr = A.r(q, j), yqi = A.c(q, j);
          const double wr = er * yqr - ei * yqi, wi = er * yqi + ei * yqr;     // e y_q
          A.r(p, j) = c * ypr - s * wr; A.c(p, j) = c * ypi - s * wi;
          A.r(q, j) = s * ypr + c * wr; A.c(q, j) = s * ypi + c * wi;
        }
        A.r(p, q) = 0.0; A.c(p, q) = 0.0; A.r(q, p) = 0.0; A.c(q, p) = 0.0;
        A.r(p, p) = app; A.c(p, p) = 0.0; A.r(q, q) = aqq; A.c(q, q) = 0.0;
      }
  }
  return sweep;
}

DS_HD int argmax_diag(const CMatRef &A) {
  int best = 0;
  for (int i = 1; i < A.M; ++i) if (A.r(i, i) > A.r(best, best)) best = i;
  return best;
}

// In-place Cholesky B = L L^H of the Hermitian matrix given by its lower triangle (L replaces
// it; the strict upper triangle is not touched).  False if a pivot is not positive (zpotrf's
// "not positive definite", which scipy turns into LinAlgError).
DS_HD bool cholesky_lower(const CMatRef &B) {
  const int M = B.M;
  for (int j = 0; j < M; ++j) {
    double d = B.r(j, j);
    for (int k = 0; k < j; ++k) d -= B.r(j, k) * B.r(j, k) + B.c(j, k) * B.c(j, k);
    if (!(d > 0.0)) return false;
    const double l = sqrt(d);
    B.r(j, j) = l; B.c(j, j) = 0.0;
    for (int i = j + 1; i < M; ++i) {
      double sr = B.r(i, j), si = B.c(i, j);
      for (int k = 0; k < j; ++k) {            // - L_ik conj(L_jk)
        sr -= B.r(i, k) * B.r(j, k) + B.c(i, k) * B.c(j, k);
        si -= B.c(i, k) * B.r(j, k) - B.r(i, k) * B.c(j, k);
      }
      B.r(i, j) = sr / l; B.c(i, j) = si / l;
    }
  }
  return true;
}

// A <- L^-1 A L^-H (the reduction of A v = lambda B v to standard form, zhegst itype 1); A full.
DS_HD void reduce_to_standard(const CMatRef &A, const CMatRef &L) {
  const int M = A.M;
  for (int j = 0; j < M; ++j)                  // columns: forward substitution with L
    for (int i = 0; i < M; ++i) {
      double sr = A.r(i, j), si = A.c(i, j);
      for (int k = 0; k < i; ++k) {
        sr -= L.r(i, k) * A.r(k, j) - L.c(i, k) * A.c(k, j);
        si -= L.r(i, k) * A.c(k, j) + L.c(i, k) * A.r(k, j);
      }
      A.r(i, j) = sr / L.r(i, i); A.c(i, j) = si / L.r(i, i);
    }
  for (int i = 0; i < M; ++i)                  // rows: X L^H = A_row
    for (int j = 0; j < M; ++j) {
      double sr = A.r(i, j), si = A.c(i, j);
      for (int k = 0; k < j; ++k) {            // - X_k conj(L_jk)
        sr -= A.r(i, k) * L.r(j, k) + A.c(i, k) * L.c(j, k);
        si -= A.c(i, k) * L.r(j, k) - A.r(i, k) * L.c(j, k);
      }
      A.r(i, j) = sr / L.r(j, j); A.c(i, j) = si / L.r(j, j);
    }
}

// w <- L^-H x (column `col` of V), written to wr/wi[i * st]
DS_HD void back_substitute_LH(const CMatRef &L, const CMatRef &V, int col, double *wr, double *wi, int st) {
  const int M = L.M;
  for (int i = M - 1; i >= 0; --i) {
    double sr = V.r(i, col), si = V.c(i, col);
    for (int k = i + 1; k < M; ++k) {          // - conj(L_ki) w_k
      sr -= L.r(k, i) * wr[k * st] + L.c(k, i) * wi[k * st];
      si -= L.r(k, i) * wi[k * st] - L.c(k, i) * wr[k * st];
    }
    wr[i * st] = sr / L.r(i, i); wi[i * st] = si / L.r(i, i);
  }
}

}  // namespace ds
