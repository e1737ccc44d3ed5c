// multibeam_tc.cu -- fixed / superdirective multi-beam weighting on the 5th-generation tensor cores.
//
//   Y[b, k, t] = sum_m conj(W[b, k, m]) X[k, t, m]        FixedBeamformer.process_freframe, einsum 'ij,ij->i'
//                                                         (beamformer/fixedbeamformer.py:147-165) for B look directions
//                                                         at once (weights: compute_weights :109-145)
//
// Per frequency bin this is a real GEMM [Re Y | Im Y] (128 beams x 2*64 frames) = A' (128 x KD) . B'^T (KD x 128):
//   conj(w) x = (wr xr + wi xi) + j (wr xi - wi xr)
//   A'[b]      = [ wr_m ... | wi_m ... ]
//   B'[t]      = [ xr_m ... |  xi_m ... ]       -> Re Y
//   B'[64 + t] = [ xi_m ... | -xr_m ... ]       -> Im Y
// The tensor cores multiply tf32 operands; one tf32 rounding of weights and spectrum (2^-11) would leave only ~70 dB
// against the float64 reference and breach the 1e-4 max-abs contract, so both operands are split into a tf32 head and a
// tf32 tail (x = hi + lo) and the three significant cross terms are evaluated: the contraction depth becomes
// KD = 3 * 2M with A'' = [hi | hi | lo], B'' = [hi | lo | hi] (error ~2^-21, the fp32 accumulator's own level).
// The arithmetic is trivial next to the data movement -- per bin and tile 2 x 24 KB of operands produce 64 KB of output --
// so the kernel is built around the stores: operand tiles are pre-packed once (weights once per look-direction set) in the
// canonical no-swizzle K-major layout and streamed with cp.async.bulk, accumulators are double-buffered in TMEM, and the
// eight epilogue warps write each beam's 64 frames of a bin as contiguous 512-byte runs into Y[s][b][k][t] (frame index
// innermost), which ds_istft_frames_inner_run overlap-adds per beam.
//   warps 0-7  epilogue: tcgen05.ld -> interleave (re, im) -> 16-byte stores
//   warp  8    one elected thread issues tcgen05.mma (kind::tf32, M128 N128 K8, KD/8 per bin)
//   warp  9    one elected thread streams the A and B tiles of each bin into shared memory (2 stages)
#include "common.cuh"
#include "tc_common.cuh"

namespace ds {
namespace mbtc {

using namespace tc;

constexpr int TILE_B = 128, TILE_T = 64, UMMA_N = 2 * TILE_T;
constexpr int STAGES = 2, ACC_STAGES = 2, EPI_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS, COPY_WARP = EPI_WARPS + 1, NTHREADS = (EPI_WARPS + 2) * 32;
constexpr int EPI_COLS = TILE_T / (EPI_WARPS / 4);      // frames per epilogue thread (32)

__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = to_tf32(v);
  lo = to_tf32(v - hi);
}

// Weights -> A'' tiles [K][beam tile][128 x KD] (canonical layout), once per set of look directions.
// W [NB][K][MM] complex64; one thread per 16-byte chunk of 4 consecutive contraction indices.
template <int MM>
__global__ void mb_pack_weights_kernel(const float2 *__restrict__ W, unsigned char *__restrict__ Ap, int NB, int K, int n_btiles) {
  constexpr int KD = 6 * MM, CHUNKS = KD / 4;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K * n_btiles * TILE_B * CHUNKS;
  if (g >= total) return;
  const int q = (int)(g % CHUNKS), r = (int)((g / CHUNKS) % TILE_B);
  const long long kt = g / ((long long)CHUNKS * TILE_B);          // k * n_btiles + tile
  const int tile = (int)(kt % n_btiles), k = (int)(kt / n_btiles);
  const int b = tile * TILE_B + r;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (b < NB) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e, part = c / (2 * MM), cc = c % (2 * MM);          // part 0, 1: head; 2: tail
      const float2 w = W[((size_t)b * K + k) * MM + (cc < MM ? cc : cc - MM)];
      float hi, lo;
      split_tf32(cc < MM ? w.x : w.y, hi, lo);
      v[e] = (part < 2) ? hi : lo;
    }
  }
  *reinterpret_cast<float4 *>(Ap + (size_t)kt * (TILE_B * KD * 4) + tile_off<TILE_B>(r, 4 * q)) = make_float4(v[0], v[1], v[2], v[3]);
}

// Spectrum -> B'' tiles [S][K][frame tile][128 x KD].  X [S][T][MM][K] complex64 (the STFT kernel's layout).  The lanes of a
// warp take 32 consecutive bins of the same (frame, microphone) -- coalesced 256-byte reads; their 16-byte chunks land in 32
// different tiles and are merged in L2 with the neighbouring rows written by other warps.
template <int MM>
__global__ void mb_pack_spectrum_kernel(const float2 *__restrict__ X, unsigned char *__restrict__ Bp, int S, int T, int K, int n_ttiles) {
  constexpr int KD = 6 * MM, CHUNKS = KD / 4;
  const int kgroups = (K + 31) / 32;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * kgroups * n_ttiles * UMMA_N * CHUNKS * 32;
  if (g >= total) return;
  const int kl = (int)(g % 32);
  long long rest = g / 32;
  const int q = (int)(rest % CHUNKS); rest /= CHUNKS;
  const int r = (int)(rest % UMMA_N); rest /= UMMA_N;
  const int tile = (int)(rest % n_ttiles); rest /= n_ttiles;
  const int k = (int)(rest % kgroups) * 32 + kl;
  const int s = (int)(rest / kgroups);
  if (k >= K) return;
  const bool im_row = r >= TILE_T;
  const int t = tile * TILE_T + (im_row ? r - TILE_T : r);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (t < T) {
    const float2 *src = X + (((size_t)s * T + t) * MM) * K + k;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e, part = c / (2 * MM), cc = c % (2 * MM);          // part 0, 2: head; 1: tail
      const float2 y = src[(size_t)(cc < MM ? cc : cc - MM) * K];
      const float val = (cc < MM) ? (im_row ? y.y : y.x) : (im_row ? -y.x : y.y);
      float hi, lo;
      split_tf32(val, hi, lo);
      v[e] = (part == 1) ? lo : hi;
    }
  }
  const size_t skt = ((size_t)s * K + k) * n_ttiles + tile;
  *reinterpret_cast<float4 *>(Bp + skt * (UMMA_N * KD * 4) + tile_off<UMMA_N>(r, 4 * q)) = make_float4(v[0], v[1], v[2], v[3]);
}

template <int MM>
__global__ void __launch_bounds__(NTHREADS, 1) mb_tc_kernel(const unsigned char *__restrict__ Ap, const unsigned char *__restrict__ Bp,
                                                            float2 *__restrict__ Y, int K, int NB, int n_btiles, int n_ttiles, int Tp) {
  constexpr int KD = 6 * MM;
  constexpr int A_BYTES = TILE_B * KD * 4, B_BYTES = UMMA_N * KD * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *tiles = smem;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *acc_full = empty_bar + STAGES;
  uint64_t *acc_empty = acc_full + ACC_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + ACC_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bt = blockIdx.x, tt = blockIdx.y, s = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(ACC_STAGES * UMMA_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARPS) {
    // ===================== epilogue: accumulator -> Y[s][b][k][t] ============================================
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int b = bt * TILE_B + quad * 32 + lane;                     // beam row of this thread; rows beyond NB are padding
    const bool live = b < NB;
    float2 *out = Y + (((size_t)s * NB + (live ? b : 0)) * K) * Tp + (size_t)tt * TILE_T + half * EPI_COLS;
    for (int k = 0; k < K; ++k) {
      const int as = k % ACC_STAGES;
      mbar_wait(&acc_full[as], (k / ACC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tcol = tmem_base + lane_addr + as * UMMA_N + half * EPI_COLS;
      float2 *row = out + (size_t)k * Tp;
#pragma unroll
      for (int h = 0; h < EPI_COLS / 16; ++h) {
        uint32_t re[16], im[16];
        tmem_ld16(tcol + h * 16, re);
        tmem_ld16(tcol + TILE_T + h * 16, im);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (live) {
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            *reinterpret_cast<float4 *>(row + h * 16 + j) =
                make_float4(__uint_as_float(re[j]), __uint_as_float(im[j]), __uint_as_float(re[j + 1]), __uint_as_float(im[j + 1]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc_empty[as]);
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer ===========================================================================
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UMMA_N >> 3) << 17) | ((uint32_t)(TILE_B >> 4) << 24);
    for (int k = 0; k < K; ++k) {
      const int st = k % STAGES, as = k % ACC_STAGES;
      if (k >= ACC_STAGES) mbar_wait(&acc_empty[as], ((k / ACC_STAGES) - 1) & 1);
      mbar_wait(&full_bar[st], (k / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(tiles + st * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < KD / 8; ++kk) {
          const uint64_t ad = make_desc(a_addr + kk * 2 * (TILE_B * 16), TILE_B * 16, 128);
          const uint64_t bd = make_desc(b_addr + kk * 2 * (UMMA_N * 16), UMMA_N * 16, 128);
          umma_tf32(tmem_base + as * UMMA_N, ad, bd, IDESC, kk > 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[st]);
        umma_commit(&acc_full[as]);
      }
      __syncwarp();
    }
  } else if (warp == COPY_WARP && lane == 0) {
    // ===================== operand tiles: two bulk copies per bin ==================================================
    const unsigned char *asrc = Ap + (size_t)bt * A_BYTES;
    const unsigned char *bsrc = Bp + ((size_t)s * K * n_ttiles + tt) * B_BYTES;
    for (int k = 0; k < K; ++k) {
      const int st = k % STAGES;
      if (k >= STAGES) mbar_wait(&empty_bar[st], ((k / STAGES) - 1) & 1);
      const uint32_t bar = smem_u32(&full_bar[st]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)STAGE_BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(tiles + st * STAGE_BYTES)),
                   "l"(asrc + (size_t)k * n_btiles * A_BYTES), "r"((uint32_t)A_BYTES), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(tiles + st * STAGE_BYTES + A_BYTES)),
                   "l"(bsrc + (size_t)k * n_ttiles * B_BYTES), "r"((uint32_t)B_BYTES), "r"(bar)
                   : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC_STAGES * UMMA_N));
  }
}

struct Layout {
  int n_btiles, n_ttiles, Tp;
  size_t Ap_bytes, Bp_bytes, Y_elems;
};
static Layout layout(int S, int T, int M, int K, int NB) {
  Layout L;
  L.n_btiles = (NB + TILE_B - 1) / TILE_B;
  L.n_ttiles = (T + TILE_T - 1) / TILE_T;
  L.Tp = L.n_ttiles * TILE_T;
  const size_t KD = 6 * (size_t)M;
  L.Ap_bytes = (size_t)K * L.n_btiles * TILE_B * KD * 4;
  L.Bp_bytes = (size_t)S * K * L.n_ttiles * UMMA_N * KD * 4;
  L.Y_elems = (size_t)S * NB * K * L.Tp;
  return L;
}

template <int MM>
static int run(int S, int T, int K, int NB, const float2 *W, const float2 *X, unsigned char *ws, float2 *Y, cudaStream_t st) {
  const Layout L = layout(S, T, MM, K, NB);
  constexpr int KD = 6 * MM;
  unsigned char *Ap = ws, *Bp = ws + ((L.Ap_bytes + 255) & ~(size_t)255);
  {
    const long long chunks = (long long)K * L.n_btiles * TILE_B * (KD / 4);
    mb_pack_weights_kernel<MM><<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(W, Ap, NB, K, L.n_btiles);
    DS_LAUNCH_CHECK();
  }
  {
    const long long chunks = (long long)S * ((K + 31) / 32) * 32 * L.n_ttiles * UMMA_N * (KD / 4);
    mb_pack_spectrum_kernel<MM><<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(X, Bp, S, T, K, L.n_ttiles);
    DS_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)STAGES * (TILE_B + UMMA_N) * KD * 4 + (2 * STAGES + 2 * ACC_STAGES) * 8 + 16 + 128;
  auto kern = mb_tc_kernel<MM>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(L.n_btiles, L.n_ttiles, S);
  kern<<<grid, NTHREADS, smem, st>>>(Ap, Bp, Y, K, NB, L.n_btiles, L.n_ttiles, L.Tp);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // namespace mbtc
}  // namespace ds

using namespace ds;

extern "C" {

int ds_multibeam_tc_layout(int n_streams, int n_frames, int n_mics, int n_bins, int n_beams, int *frame_pitch,
                           size_t *workspace_bytes, size_t *y_bytes) {
  DS_CHECK_ARG(n_streams >= 1 && n_frames >= 1 && n_bins >= 1 && n_beams >= 1, "ds_multibeam_tc_layout: bad shape");
  DS_CHECK_ARG(n_mics == 4 || n_mics == 8 || n_mics == 16, "ds_multibeam_tc_layout: the tensor-core path is compiled for 4, 8 or 16 microphones");
  const mbtc::Layout L = mbtc::layout(n_streams, n_frames, n_mics, n_bins, n_beams);
  if (frame_pitch) *frame_pitch = L.Tp;
  if (workspace_bytes) *workspace_bytes = ((L.Ap_bytes + 255) & ~(size_t)255) + L.Bp_bytes;
  if (y_bytes) *y_bytes = L.Y_elems * sizeof(float2);
  return DS_OK;
}

int ds_multibeam_tc_run(int n_streams, int n_frames, int n_mics, int n_bins, int n_beams, const void *W, const void *X,
                        void *workspace, void *Y, void *stream) {
  DS_CHECK_ARG(W && X && workspace && Y, "ds_multibeam_tc_run: null argument");
  DS_CHECK_ARG(n_streams >= 1 && n_streams <= 65535 && n_frames >= 1 && n_bins >= 1 && n_beams >= 1, "ds_multibeam_tc_run: bad shape");
  DS_CHECK_ARG((reinterpret_cast<size_t>(workspace) & 127) == 0 && (reinterpret_cast<size_t>(Y) & 15) == 0, "ds_multibeam_tc_run: workspace must be 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_mics) {
    case 4: return mbtc::run<4>(n_streams, n_frames, n_bins, n_beams, (const float2 *)W, (const float2 *)X, (unsigned char *)workspace, (float2 *)Y, st);
    case 8: return mbtc::run<8>(n_streams, n_frames, n_bins, n_beams, (const float2 *)W, (const float2 *)X, (unsigned char *)workspace, (float2 *)Y, st);
    case 16: return mbtc::run<16>(n_streams, n_frames, n_bins, n_beams, (const float2 *)W, (const float2 *)X, (unsigned char *)workspace, (float2 *)Y, st);
  }
  set_error("ds_multibeam_tc_run: the tensor-core path is compiled for 4, 8 or 16 microphones (got %d)", n_mics);
  return DS_EUNSUPPORTED;
}

}  // extern "C"
