// fft16.cuh -- 16-point DFT in registers, the building block of the "square" real FFT of the STFT kernel for
// n_fft = 512 (stft_sq_kernel, transform.cu): H = 256 = 16 x 16 complex points per frame, a HALF-warp per frame,
// two radix-16 passes with one 16 x 16 transpose through shared memory in between, and a real-FFT split whose
// partner elements Z[H - k] come from the mirror lane by warp shuffles.  Against the radix-8/8/4 Stockham transform of
// fft.cuh this removes two of the three shared-memory round trips and every per-frame twiddle / window load (all
// lane-invariant constants live in registers for the life of the persistent warp).
#pragma once
#include "fft.cuh"

namespace ds {

// register that holds output q of dft16 (compile-time permutation, never materialised)
__host__ __device__ constexpr int P16(int q) { return 4 * (q & 3) + (q >> 2); }

// a * exp(-2 pi i m / 16) for the m that occur between the two radix-4 stages
template <int MM> __device__ __forceinline__ float2 mul_w16(float2 a) {
  constexpr float h = 0.70710678118654752440f, c = 0.92387953251128675613f, s = 0.38268343236508977173f;
  if constexpr (MM == 0) return a;
  else if constexpr (MM == 4) return make_float2(a.y, -a.x);
  else if constexpr (MM == 2) return make_float2((a.x + a.y) * h, (a.y - a.x) * h);
  else if constexpr (MM == 6) return make_float2((a.y - a.x) * h, -(a.x + a.y) * h);
  else if constexpr (MM == 1) return make_float2(fmaf(a.y, s, a.x * c), fmaf(-a.x, s, a.y * c));
  else if constexpr (MM == 3) return make_float2(fmaf(a.y, c, a.x * s), fmaf(-a.x, c, a.y * s));
  else { static_assert(MM == 9, "unused twiddle"); return make_float2(-fmaf(a.y, s, a.x * c), fmaf(a.x, s, -(a.y * c))); }
}

// forward 16-point DFT: on exit V[q] = sum_r v_in[r] exp(-2 pi i r q / 16) sits in v[P16(q)]
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  // r = r0 + 4 r1: 4-point DFTs over r1; v[r0 + 4 qa] <- U_r0[qa]
  dft4(v[0], v[4], v[8], v[12]);
  dft4(v[1], v[5], v[9], v[13]);
  dft4(v[2], v[6], v[10], v[14]);
  dft4(v[3], v[7], v[11], v[15]);
  v[5] = mul_w16<1>(v[5]);   v[9] = mul_w16<2>(v[9]);    v[13] = mul_w16<3>(v[13]);
  v[6] = mul_w16<2>(v[6]);   v[10] = mul_w16<4>(v[10]);  v[14] = mul_w16<6>(v[14]);
  v[7] = mul_w16<3>(v[7]);   v[11] = mul_w16<6>(v[11]);  v[15] = mul_w16<9>(v[15]);
  // q = qa + 4 qb: 4-point DFTs over r0; v[4 qa + qb] <- V[qa + 4 qb]
  dft4(v[0], v[1], v[2], v[3]);
  dft4(v[4], v[5], v[6], v[7]);
  dft4(v[8], v[9], v[10], v[11]);
  dft4(v[12], v[13], v[14], v[15]);
}

// row stride (complex elements) of a half-warp's 16 x 16 transpose buffer: 16-byte aligned rows, conflict-free STS.64
// columns and LDS.128 rows
constexpr int SQ_RS = 18;

// 256-point complex forward transform by one half-warp: on entry v[r] = z[j + 16 r] (lane j of the half-warp), on exit
// u[P16(q)] = Z[j + 16 q].  xb: the half-warp's 16 x SQ_RS buffer; tw2[r] = exp(-2 pi i r j / 256).  Every lane of the
// warp must call it (two __syncwarp).
__device__ __forceinline__ void sq_cfft256(float2 (&v)[16], float2 (&u)[16], float2 *xb, const float2 (&tw2)[16], int j) {
  dft16(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) xb[q * SQ_RS + j] = v[P16(q)];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w4 = *reinterpret_cast<const float4 *>(xb + j * SQ_RS + 2 * i);
    u[2 * i] = make_float2(w4.x, w4.y); u[2 * i + 1] = make_float2(w4.z, w4.w);
  }
  __syncwarp();
#pragma unroll
  for (int r = 1; r < 16; ++r) u[r] = cmul(u[r], tw2[r]);
  dft16(u);
}

}  // namespace ds
