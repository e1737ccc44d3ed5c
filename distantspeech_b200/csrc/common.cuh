// common.cuh -- shared helpers for libds_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ds_b200.h"

namespace ds {

void set_error(const char *fmt, ...);

#define DS_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ds::set_error(__VA_ARGS__);      \
      return DS_EINVAL;                \
    }                                  \
  } while (0)

#define DS_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ds::set_error("%s:%d CUDA error %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return DS_ECUDA;                                                             \
    }                                                                              \
  } while (0)

#define DS_LAUNCH_CHECK() DS_CUDA(cudaGetLastError())

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// ---- scalar <-> vector-2 traits -------------------------------------------
template <typename T> struct V2;
template <> struct V2<float> { typedef float2 type; };
template <> struct V2<double> { typedef double2 type; };

template <typename T> __host__ __device__ __forceinline__ typename V2<T>::type mk2(T a, T b) {
  typename V2<T>::type r; r.x = a; r.y = b; return r;
}

// complex multiply a*b
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
  C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}

// twiddle tables: tw_h[i] = exp(-2 pi i / H) i<H ; tw_n[k] = exp(-2 pi k / N) k<=H/2 (N = 2H)
struct TwiddleSet {
  const float2 *h32; const float2 *n32;
  const double2 *h64; const double2 *n64;
};
int get_twiddles(int n_fft, TwiddleSet *out);   // DS_OK / DS_EUNSUPPORTED / DS_ECUDA

template <typename T> struct TwSel;
template <> struct TwSel<float> {
  static const float2 *h(const TwiddleSet &t) { return t.h32; }
  static const float2 *n(const TwiddleSet &t) { return t.n32; }
};
template <> struct TwSel<double> {
  static const double2 *h(const TwiddleSet &t) { return t.h64; }
  static const double2 *n(const TwiddleSet &t) { return t.n64; }
};

}  // namespace ds
