// herm.cuh -- small Hermitian matrices held in registers (packed: real diagonal + strictly upper part)
#pragma once
#include <type_traits>
#include "perbin.cuh"

namespace ds {

template <int B, int E, typename F> __device__ __forceinline__ void sfor2(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    sfor2<B + 1, E>(f);
  }
}
#define SIDX2(ic) (decltype(ic)::value)

// Hermitian matrix in registers: d[M] real diagonal, ur/ui[NQ] strictly upper part.
template <int M> struct Herm {
  static constexpr int NQ = M * (M - 1) / 2;
  double d[M];
  double ur[NQ > 0 ? NQ : 1], ui[NQ > 0 ? NQ : 1];
};

// In-place inverse of a Hermitian positive-definite matrix by Hermitian sweeps
// (A_ij -= A_ik conj(A_jk) / A_kk; after all pivots the array holds -A^-1).
// IEEE_DIV: pivots are inverted with a true division (any sign, inf/NaN propagate) instead of the
// call-free positive-only reciprocal -- for callers that may hand in a matrix that is not
// numerically positive definite.
template <int M, bool IEEE_DIV = false> __device__ __forceinline__ void herm_inverse(Herm<M> &h) {
  sfor2<0, M>([&](auto kc) {
    constexpr int k = SIDX2(kc);
    const double r = IEEE_DIV ? 1.0 / h.d[k] : rcp_pos(h.d[k]);
    double cr[M], ci[M];       // column k: c_i = A_ik
    sfor2<0, M>([&](auto ic) {
      constexpr int i = SIDX2(ic);
      if constexpr (i < k) { cr[i] = h.ur[qidx<M>(i, k)]; ci[i] = h.ui[qidx<M>(i, k)]; }
      else if constexpr (i > k) { cr[i] = h.ur[qidx<M>(k, i)]; ci[i] = -h.ui[qidx<M>(k, i)]; }
    });
    sfor2<0, M>([&](auto ic) {
      constexpr int i = SIDX2(ic);
      if constexpr (i != k) {
        const double tr = cr[i] * r, ti = ci[i] * r;
        // diagonal: A_ii -= |c_i|^2 r
        h.d[i] = fma(-tr, cr[i], fma(-ti, ci[i], h.d[i]));
        sfor2<i + 1, M>([&](auto jc) {
          constexpr int j = SIDX2(jc);
          if constexpr (j != k) {
            // A_ij -= t_i conj(c_j)
            h.ur[qidx<M>(i, j)] = fma(-tr, cr[j], fma(-ti, ci[j], h.ur[qidx<M>(i, j)]));
            h.ui[qidx<M>(i, j)] = fma(-ti, cr[j], fma(tr, ci[j], h.ui[qidx<M>(i, j)]));
          }
        });
      }
    });
    sfor2<0, M>([&](auto ic) {
      constexpr int i = SIDX2(ic);
      if constexpr (i < k) { h.ur[qidx<M>(i, k)] = cr[i] * r; h.ui[qidx<M>(i, k)] = ci[i] * r; }
      else if constexpr (i > k) { h.ur[qidx<M>(k, i)] = cr[i] * r; h.ui[qidx<M>(k, i)] = -ci[i] * r; }
    });
    h.d[k] = -r;
  });
  sfor2<0, M>([&](auto ic) { constexpr int i = SIDX2(ic); h.d[i] = -h.d[i]; });
  sfor2<0, Herm<M>::NQ>([&](auto ec) { constexpr int e = SIDX2(ec); h.ur[e] = -h.ur[e]; h.ui[e] = -h.ui[e]; });
}

}  // namespace ds
