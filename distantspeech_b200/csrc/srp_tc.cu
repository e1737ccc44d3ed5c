// srp_tc.cu -- SRP-PHAT steered-response contraction on the 5th-generation tensor cores.
//
//   P[d, t] = sum_k | sum_m conj(a[d,k,m]) yhat[k,t,m] |          (doa/srp.py:45-51)
//
// Per frequency bin the inner sum is a real GEMM with a contraction depth of only 2M:
//   [Re Z | Im Z] (128 directions x 2*64 frames) = A' (128 x 2M) . B'^T (2M x 128)
//   A'[d]      = [ cos(w tau_dm) ... | -sin(w tau_dm) ... ]                 (a = cos - j sin)
//   B'[t]      = [ yr_m ... |  yi_m ... ]   -> Re Z = sum ar yr + ai yi
//   B'[64 + t] = [ yi_m ... | -yr_m ... ]   -> Im Z = sum ar yi - ai yr
// so the kernel is bound by what surrounds the MMA: generating the steering tile on
// chip (it is never stored: D x K x M complex would be 4 GB) and the |.| + sum_k
// epilogue out of tensor memory.  One CTA = 128 directions x 64 frames, looping over
// all bins:
//   warps 0-7   epilogue: tcgen05.ld the 128 x 128 fp32 accumulator, |z|, accumulate over bins
//               (warp w reads TMEM lanes 32 (w % 4) .. +31 and the frame half w / 4; the epilogue is
//               bound by the special-function unit -- one sqrt per (direction, frame, bin) -- so it
//               gets two warps per scheduler to keep that unit fed across TMEM-load latency)
//   warps 8-11  producers: steering tile generated on chip (phasor recurrence) into shared memory
//               in the canonical no-swizzle K-major core-matrix layout, 3 stages
//   warp  12    one elected thread issues tcgen05.mma (kind::tf32, M128 N128 K8, 2M/8 per bin)
//   warp  13    one elected thread streams the spectrum tile of each bin into shared memory with
//               cp.async.bulk (the tile was laid out and rounded to tf32 once by srp_pack_kernel)
//   warps 14-15 idle (they complete the fourth warpgroup: setmaxnreg is a warpgroup-wide instruction)
// Registers are rebalanced with setmaxnreg (launch 64: epilogue 72, producers 88, last warpgroup 24).
// TMEM: 2 accumulator stages x 128 columns.  Synchronisation: mbarriers
// (producer -> MMA -> producer, MMA -> epilogue -> MMA) with tcgen05.commit.
#include "common.cuh"
#include "tc_common.cuh"

namespace ds {

namespace tc {

constexpr int TILE_D = 128, TILE_T = 64, UMMA_N = 2 * TILE_T;
constexpr int STAGES = 3, ACC_STAGES = 2;
constexpr int EPI_WARPS = 8, PROD_WARPS = 4;
constexpr int EPI_COLS = TILE_T / (EPI_WARPS / 4);      // frames per epilogue thread
constexpr int RESYNC = 32;            // bins between exact re-evaluations of the steering phasors
constexpr int PROD_WARP0 = EPI_WARPS, MMA_WARP = EPI_WARPS + PROD_WARPS, COPY_WARP = MMA_WARP + 1;
constexpr int NTHREADS = (EPI_WARPS + PROD_WARPS + 4) * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;

// Spectrum tiles for the B operand, built once per call: for bin k and frame tile j the 128 x 2M tile
//   row r < 64 : [ yr_m ... |  yi_m ... ] of frame 64 j + r          (-> Re Z)
//   row 64 + r : [ yi_m ... | -yr_m ... ]                            (-> Im Z)
// rounded to tf32 and stored in the canonical layout, so the contraction kernel moves it with one
// bulk copy per bin.  One thread per 16-byte chunk.
template <int MM>
__global__ void srp_pack_kernel(const float2 *__restrict__ Yhat, unsigned char *__restrict__ Bp, int T, int K, int n_tiles) {
  constexpr int KD = 2 * MM, CHUNKS = KD / 4;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K * n_tiles * UMMA_N * CHUNKS;
  if (g >= total) return;
  const int q = (int)(g % CHUNKS);
  const int r = (int)((g / CHUNKS) % UMMA_N);
  const long long kt = g / ((long long)CHUNKS * UMMA_N);        // k * n_tiles + tile
  const int tile = (int)(kt % n_tiles);
  const int k = (int)(kt / n_tiles);
  const bool im_row = r >= TILE_T;
  const int t = tile * TILE_T + (im_row ? r - TILE_T : r);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (t < T) {
    const float2 *src = Yhat + ((size_t)k * T + t) * MM;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      const float2 y = src[c < MM ? c : c - MM];
      const float val = (c < MM) ? (im_row ? y.y : y.x) : (im_row ? -y.x : y.y);
      v[e] = to_tf32(val);
    }
  }
  *reinterpret_cast<float4 *>(Bp + (size_t)kt * (UMMA_N * KD * 4) + tile_off<UMMA_N>(r, 4 * q)) = make_float4(v[0], v[1], v[2], v[3]);
}

template <int MM>   // microphones
__global__ void __launch_bounds__(NTHREADS, 2) srp_tc_kernel(const float *__restrict__ tau, const unsigned char *__restrict__ Bp,
                                                             float *__restrict__ P, int D, int T, int K, float two_f0, int n_tiles) {
  constexpr int KD = 2 * MM;                      // contraction depth (fp32 / tf32 elements)
  constexpr int A_BYTES = TILE_D * KD * 4, B_BYTES = UMMA_N * KD * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *tiles = smem;                                            // [STAGES][A | B]
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *acc_full = empty_bar + STAGES;
  uint64_t *acc_empty = acc_full + ACC_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + ACC_STAGES);
  float *tau_s = reinterpret_cast<float *>(tmem_slot + 4);                 // [TILE_D][MM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = blockIdx.x * TILE_D, t0 = blockIdx.y * TILE_T;

  for (int i = threadIdx.x; i < TILE_D * MM; i += NTHREADS) {
    const int d = d0 + i / MM;
    tau_s[i] = (d < D) ? tau[(size_t)d * MM + (i % MM)] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], PROD_THREADS + 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {      // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(ACC_STAGES * UMMA_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARPS) {
    // ===================== epilogue: |z| and sum over bins ==========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
    float acc[EPI_COLS];
#pragma unroll
    for (int j = 0; j < EPI_COLS; ++j) acc[j] = 0.f;
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    for (int k = 0; k < K; ++k) {
      const int as = k % ACC_STAGES;
      mbar_wait(&acc_full[as], (k / ACC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tcol = tmem_base + lane_addr + as * UMMA_N + half * EPI_COLS;
#pragma unroll
      for (int h = 0; h < EPI_COLS / 16; ++h) {
        uint32_t re[16], im[16];
        tmem_ld16(tcol + h * 16, re);
        tmem_ld16(tcol + TILE_T + h * 16, im);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float zr = __uint_as_float(re[j]), zi = __uint_as_float(im[j]);
#ifdef SRP_DBG_EPI_LIGHT
          acc[h * 16 + j] += zr + zi;
#else
          float mag;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(zr, zr, zi * zi)));
          acc[h * 16 + j] += mag;
#endif
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc_empty[as]);
    }
    const int d = d0 + quad * 32 + lane;
    if (d < D) {
      const int tb = t0 + half * EPI_COLS;
      float *out = P + (size_t)d * T + tb;
      const bool vec = (tb + EPI_COLS <= T) && ((reinterpret_cast<size_t>(out) & 15) == 0);
      if (vec) {
#pragma unroll
        for (int j = 0; j < EPI_COLS; j += 4) *reinterpret_cast<float4 *>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < EPI_COLS; ++j)
          if (tb + j < T) out[j] = acc[j];
      }
    }
  } else if (warp >= MMA_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == MMA_WARP) {
    // ===================== MMA issuer ==================================================
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 128, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UMMA_N >> 3) << 17) | ((uint32_t)(TILE_D >> 4) << 24);
    for (int k = 0; k < K; ++k) {
      const int s = k % STAGES, as = k % ACC_STAGES;
      if (k >= ACC_STAGES) mbar_wait(&acc_empty[as], ((k / ACC_STAGES) - 1) & 1);
      mbar_wait(&full_bar[s], (k / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(tiles + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < KD / 8; ++kk) {
          // one K = 8 step = two 16-byte chunks of the canonical layout
          const uint64_t ad = make_desc(a_addr + kk * 2 * (TILE_D * 16), TILE_D * 16, 128);
          const uint64_t bd = make_desc(b_addr + kk * 2 * (UMMA_N * 16), UMMA_N * 16, 128);
          umma_tf32(tmem_base + as * UMMA_N, ad, bd, IDESC, kk > 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);     // smem stage may be refilled once these MMAs retire
        umma_commit(&acc_full[as]);     // accumulator ready for the epilogue
      }
      __syncwarp();
    }
    } else if (warp == COPY_WARP && lane == 0) {
    // ===================== spectrum tiles: one bulk copy per bin ==============================
      const unsigned char *src = Bp + (size_t)blockIdx.y * B_BYTES;
      for (int k = 0; k < K; ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(&empty_bar[s], ((k / STAGES) - 1) & 1);
        const uint32_t bar = smem_u32(&full_bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)B_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(tiles + s * STAGE_BYTES + A_BYTES)),
                     "l"(src + (size_t)k * n_tiles * B_BYTES), "r"((uint32_t)B_BYTES), "r"(bar)
                     : "memory");
      }
    }
  } else {
    // ===================== producers: steering tile ===========================================
    // Thread pt owns steering row pt (one direction, all MM mics): its phasors advance from bin to
    // bin by one complex rotation exp(-j 2 pi df tau) and are re-evaluated exactly every RESYNC bins.
    // All shared-memory traffic is 16-byte vectors.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int pt = threadIdx.x - PROD_WARP0 * 32;          // 0 .. 127
    float cs[MM], sn[MM], rc[MM], rs[MM];
#pragma unroll
    for (int m = 0; m < MM; ++m) sincospif(two_f0 * tau_s[pt * MM + m], &rs[m], &rc[m]);   // rotation per bin
    const int a_row = (pt >> 3) * 128 + (pt & 7) * 16;
    for (int k = 0; k < K; ++k) {
      const int s = k % STAGES;
      if ((k % RESYNC) == 0) {
        const float fk2 = two_f0 * (float)k;
#pragma unroll
        for (int m = 0; m < MM; ++m) sincospif(fk2 * tau_s[pt * MM + m], &sn[m], &cs[m]);
      }
      if (k >= STAGES) mbar_wait(&empty_bar[s], ((k / STAGES) - 1) & 1);
      unsigned char *At = tiles + s * STAGE_BYTES;
      // steering row: [cos_0 .. cos_{M-1} | -sin_0 .. -sin_{M-1}]   (a = cos - j sin)
#pragma unroll
      for (int q = 0; q < MM / 4; ++q) {
        *reinterpret_cast<float4 *>(At + q * (TILE_D * 16) + a_row) =
            make_float4(to_tf32(cs[4 * q]), to_tf32(cs[4 * q + 1]), to_tf32(cs[4 * q + 2]), to_tf32(cs[4 * q + 3]));
        *reinterpret_cast<float4 *>(At + (MM / 4 + q) * (TILE_D * 16) + a_row) =
            make_float4(-to_tf32(sn[4 * q]), -to_tf32(sn[4 * q + 1]), -to_tf32(sn[4 * q + 2]), -to_tf32(sn[4 * q + 3]));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (tensor core)
      mbar_arrive(&full_bar[s]);
      // advance the phasors to the next bin
#ifndef SRP_DBG_PROD_LIGHT
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        const float c = cs[m], sv = sn[m];
        cs[m] = fmaf(c, rc[m], -sv * rs[m]);
        sn[m] = fmaf(sv, rc[m], c * rs[m]);
      }
#endif
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC_STAGES * UMMA_N));
  }
}

template <int MM> static int launch(const float *tau, const float2 *Yhat, unsigned char *Bp, float *P, int D, int T, int K,
                                    float two_f0, cudaStream_t st) {
  constexpr int KD = 2 * MM;
  const int n_tiles = (T + TILE_T - 1) / TILE_T;
  const long long chunks = (long long)K * n_tiles * UMMA_N * (KD / 4);
  srp_pack_kernel<MM><<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(Yhat, Bp, T, K, n_tiles);
  DS_LAUNCH_CHECK();
  const size_t smem = (size_t)STAGES * (TILE_D + UMMA_N) * KD * 4 + (2 * STAGES + 2 * ACC_STAGES) * 8 + 16 + (size_t)TILE_D * MM * 4 + 128;
  auto kern = srp_tc_kernel<MM>;
  DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((D + TILE_D - 1) / TILE_D, n_tiles);
  kern<<<grid, NTHREADS, smem, st>>>(tau, Bp, P, D, T, K, two_f0, n_tiles);
  DS_LAUNCH_CHECK();
  return DS_OK;
}

}  // namespace tc

bool srp_tc_supported(int D, int T, int M, int K) { return (M == 4 || M == 8 || M == 16) && D >= 1 && T >= 1 && K >= 1; }

size_t srp_tc_workspace_bytes(int T, int M, int K) {
  return (size_t)K * ((T + tc::TILE_T - 1) / tc::TILE_T) * tc::UMMA_N * 2 * M * 4;
}

int srp_tc_launch(const float *tau, const float2 *Yhat, void *workspace, float *P, int D, int T, int M, int K, float two_f0,
                  cudaStream_t st) {
  unsigned char *Bp = (unsigned char *)workspace;
  switch (M) {
    case 4: return tc::launch<4>(tau, Yhat, Bp, P, D, T, K, two_f0, st);
    case 8: return tc::launch<8>(tau, Yhat, Bp, P, D, T, K, two_f0, st);
    case 16: return tc::launch<16>(tau, Yhat, Bp, P, D, T, K, two_f0, st);
  }
  set_error("srp tensor-core path: unsupported microphone count %d", M);
  return DS_EUNSUPPORTED;
}

}  // namespace ds
