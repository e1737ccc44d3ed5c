// srp_tc.cu -- SRP-PHAT steered-response contraction on the 5th-generation tensor cores.
//
//   P[d, t] = sum_k | sum_m conj(a[d,k,m]) yhat[k,t,m] |          (doa/srp.py:45-51)
//
// Per frequency bin the inner sum is a real GEMM with a contraction depth of only 2M:
//   [Re Z | Im Z] (128 directions x 2*128 frames) = A' (128 x 2M) . B'^T (2M x 256)
//   A'[d]      = [ cos(w tau_dm) ... | -sin(w tau_dm) ... ]                 (a = cos - j sin)
//   B'[t]      = [ yr_m ... |  yi_m ... ]   -> Re Z = sum ar yr + ai yi
//   B'[128 + t] = [ yi_m ... | -yr_m ... ]  -> Im Z = sum ar yi - ai yr
// so the kernel is bound by what surrounds the MMA: generating the steering tile on
// chip (it is never stored: D x K x M complex would be 4 GB) and the |.| + sum_k
// epilogue out of tensor memory.  One CTA (one per SM) = 128 directions x 128 frames, looping over
// all bins -- the steering tile costs as many issue slots per bin as half the epilogue, so it is
// shared by as many frames as tensor memory holds (two 256-column accumulators = all 512 columns):
//   warps 0-15  epilogue: tcgen05.ld the 128 x 256 fp32 accumulator, |z|, accumulate over bins
//               (warp w reads TMEM lanes 32 (w % 4) .. +31 and the frame quarter w / 4; packed fp32
//               (f32x2) squares and sums, one special-function sqrt per (direction, frame, bin): 16 results
//               per clock and SM, 1 024 cycles per bin and tile against 785 for the four MMAs)
//   warps 16-23 producers (two threads per direction, half the microphones each): steering tile generated
//               on chip (phasor recurrence on (cos, -sin)) into shared memory in the canonical no-swizzle
//               K-major core-matrix layout, 4 stages;
//               rounding to tf32 is one integer add per value (+ half an ulp; the tensor core
//               truncates the rest -- the same result as cvt.rna, which costs four instructions)
//   warp  24    one elected thread issues tcgen05.mma (kind::tf32, M128 N256 K8, 2M/8 per bin)
//   warp  25    one elected thread streams the spectrum tile of each bin into shared memory with
//               cp.async.bulk (the tile was laid out and rounded to tf32 once by srp_pack_kernel)
//   warps 26-27 idle (they complete the seventh warpgroup: setmaxnreg is a warpgroup-wide instruction)
// Registers are rebalanced with setmaxnreg (launch 72: epilogue 80, producers 64, last warpgroup 24).
// TMEM: 2 accumulator stages x 256 columns.  Synchronisation: mbarriers
// (producer -> MMA -> producer, MMA -> epilogue -> MMA) with tcgen05.commit.
// Round 1's shape (128 x 64 tiles, two CTAs per SM, cvt.rna, scalar epilogue) needed 2 980 issue slots per
// 8 192 accumulator elements (epilogue 1 200, producers 1 040, waits) and ran at 6.76 ms for config 5; this shape needs
// 1 520 and runs at 5.5 ms.  Where the time goes (cycles per bin and tile = 16 384 accumulator elements, role variants
// behind the SRP_DBG_* macros, profiles/ab_runs_r02.txt): the four MMAs free-running 785 (the K = 8 operand fetch from
// shared memory, not the 512-cycle math), + the two-stage accumulator hand-off 850, + producers and bulk copies 1 022,
// + epilogue 1 518; the sqrt unit alone would allow 1 024.  The costs add instead of overlapping: with all 512
// tensor-memory columns in two stages the MMA of bin k + 2 cannot start before the epilogue of bin k has drained.
#include "common.cuh"
#include "tc_common.cuh"
#include <type_traits>

namespace ds {

namespace tc {

constexpr int TILE_D = 128, TILE_T = 128, UMMA_N = 2 * TILE_T;
// The accumulator of a bin can be produced and handed over in NGRP frame groups ([Re | Im] of GRP_T frames each, one
// M128 N(2 GRP_T) MMA chain and one barrier pair per group), so that the epilogue warps of one group work while the tensor
// core still computes the next one.  MEASURED (config 5): NGRP = 2 5.95 ms against 5.56 ms for NGRP = 1 -- the N = 128
// MMAs cost more than the de-phasing of the epilogue warps gains; a software-pipelined tensor-memory load (next 8 columns
// in flight during the square roots of the current 8) was 5.87 ms.  Default: one group, plain load / wait / compute.
#ifndef SRP_NGRP
#define SRP_NGRP 1
#endif
constexpr int NGRP = SRP_NGRP, GRP_T = TILE_T / NGRP, GRP_N = 2 * GRP_T;
#ifndef SRP_STAGES
#define SRP_STAGES 4
#endif
constexpr int STAGES = SRP_STAGES, ACC_STAGES = 2;
constexpr int EPI_WARPS = 16, PROD_WARPS = 8;
constexpr int EPI_COLS = TILE_T / (EPI_WARPS / 4);      // frames per epilogue thread
constexpr int RESYNC = 32;            // bins between exact re-evaluations of the steering phasors
constexpr int PROD_WARP0 = EPI_WARPS, MMA_WARP = EPI_WARPS + PROD_WARPS, COPY_WARP = MMA_WARP + 1;
constexpr int NTHREADS = (EPI_WARPS + PROD_WARPS + 4) * 32;

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
  const int grp = r / GRP_N, rr = r % GRP_N;                    // rows: [Re grp 0 | Im grp 0 | Re grp 1 | Im grp 1]
  const bool im_row = rr >= GRP_T;
  const int t = tile * TILE_T + grp * GRP_T + (im_row ? rr - GRP_T : rr);
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
__global__ void __launch_bounds__(NTHREADS, 1) srp_tc_kernel(const float *__restrict__ tau, const unsigned char *__restrict__ Bp,
                                                             float *__restrict__ P, int D, int T, int K, float two_f0, int n_tiles) {
  constexpr int KD = 2 * MM;                      // contraction depth (fp32 / tf32 elements)
  constexpr int A_BYTES = TILE_D * KD * 4, B_BYTES = UMMA_N * KD * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // a steering row (one direction) is produced by PSPLIT threads, MH microphones each: one producer warp per scheduler
  // took ~1 000 cycles per bin (128 dependent-ish instructions, shared-memory stores, proxy fence) and was the critical
  // path of the whole pipeline (MMA + producers alone: 1 022 cycles per bin against 843 for the MMAs)
  constexpr int PSPLIT = MM >= 8 ? 2 : 1, MH = MM / PSPLIT, PROD_ACTIVE = TILE_D * PSPLIT;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *tiles = smem;                                            // [STAGES][A | B]
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *acc_full = empty_bar + STAGES;
  uint64_t *acc_empty = acc_full + ACC_STAGES * NGRP;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + ACC_STAGES * NGRP);
  float *tau_s = reinterpret_cast<float *>(tmem_slot + 4);                 // [TILE_D][MM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = blockIdx.x * TILE_D, t0 = blockIdx.y * TILE_T;

  for (int i = threadIdx.x; i < TILE_D * MM; i += NTHREADS) {
    const int d = d0 + i / MM;
    tau_s[i] = (d < D) ? tau[(size_t)d * MM + (i % MM)] : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], PROD_ACTIVE + 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES * NGRP; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS * 32 / NGRP); }
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    float acc[EPI_COLS];
#pragma unroll
    for (int j = 0; j < EPI_COLS; ++j) acc[j] = 0.f;
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int grp = half / (EPI_WARPS / 4 / NGRP), hh = half % (EPI_WARPS / 4 / NGRP);
    const int tb0 = t0 + grp * GRP_T + hh * EPI_COLS;
    const uint32_t tcol0 = tmem_base + lane_addr + grp * GRP_N + hh * EPI_COLS;
    // |z| of eight (direction, frame) pairs, two per instruction (packed fp32), into the running sums of column group h
    auto mag8 = [&](const uint32_t (&re)[8], const uint32_t (&im)[8], auto hc) {
      constexpr int h = decltype(hc)::value;
      if (tb0 + h * 8 < T) {                                  // warp-uniform: frame groups past the end stay zero
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
#ifdef SRP_DBG_EPI_LIGHT
          acc[h * 8 + j] += __uint_as_float(re[j]) + __uint_as_float(im[j]);
          acc[h * 8 + j + 1] += __uint_as_float(re[j + 1]) + __uint_as_float(im[j + 1]);
#else
          unsigned long long zr2, zi2, s2, a2, m2;
          asm("mov.b64 %0, {%1, %2};" : "=l"(zr2) : "r"(re[j]), "r"(re[j + 1]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(zi2) : "r"(im[j]), "r"(im[j + 1]));
          asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(s2) : "l"(zi2));
          asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(s2) : "l"(zr2), "l"(s2));
          float s_lo, s_hi, m_lo, m_hi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(s_lo), "=f"(s_hi) : "l"(s2));
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m_lo) : "f"(s_lo));
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m_hi) : "f"(s_hi));
          asm("mov.b64 %0, {%1, %2};" : "=l"(m2) : "f"(m_lo), "f"(m_hi));
          asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(acc[h * 8 + j]), "f"(acc[h * 8 + j + 1]));
          asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a2) : "l"(a2), "l"(m2));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[h * 8 + j]), "=f"(acc[h * 8 + j + 1]) : "l"(a2));
#endif
        }
      }
    };
    for (int k = 0; k < K; ++k) {
      const int as = k % ACC_STAGES;
      mbar_wait(&acc_full[as * NGRP + grp], (k / ACC_STAGES) & 1);     // (one polling / arriving lane per warp: measured slower)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tcol = tcol0 + as * UMMA_N;
#ifdef SRP_DBG_EPI_NONE
      if (k >= 0) { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); mbar_arrive(&acc_empty[as * NGRP + grp]); continue; }
#endif
      uint32_t re0[8], im0[8], re1[8], im1[8];
      tmem_ld8(tcol, re0); tmem_ld8(tcol + 8, re1); tmem_ld8(tcol + GRP_T, im0); tmem_ld8(tcol + GRP_T + 8, im1);
      tmem_wait_ld(re0, im0); tmem_wait_ld(re1, im1);
      mag8(re0, im0, std::integral_constant<int, 0>{});
      mag8(re1, im1, std::integral_constant<int, 1>{});
      tmem_ld8(tcol + 16, re0); tmem_ld8(tcol + 24, re1); tmem_ld8(tcol + GRP_T + 16, im0); tmem_ld8(tcol + GRP_T + 24, im1);
      tmem_wait_ld(re0, im0); tmem_wait_ld(re1, im1);
      mag8(re0, im0, std::integral_constant<int, 2>{});
      mag8(re1, im1, std::integral_constant<int, 3>{});
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc_empty[as * NGRP + grp]);
    }
    const int d = d0 + quad * 32 + lane;
    if (d < D) {
      const int tb = tb0;
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
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GRP_N >> 3) << 17) | ((uint32_t)(TILE_D >> 4) << 24);
    for (int k = 0; k < K; ++k) {
      const int s = k % STAGES, as = k % ACC_STAGES;
      mbar_wait(&full_bar[s], (k / STAGES) & 1);
#pragma unroll
      for (int g = 0; g < NGRP; ++g) {
#ifndef SRP_DBG_NO_ACC_WAIT
        if (k >= ACC_STAGES) mbar_wait(&acc_empty[as * NGRP + g], ((k / ACC_STAGES) - 1) & 1);
#endif
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(tiles + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES + g * (GRP_N / 8) * 128;     // rows GRP_N g .. of the 256-row tile
#pragma unroll
          for (int kk = 0; kk < KD / 8; ++kk) {
            // one K = 8 step = two 16-byte chunks of the canonical layout
            const uint64_t ad = make_desc(a_addr + kk * 2 * (TILE_D * 16), TILE_D * 16, 128);
            const uint64_t bd = make_desc(b_addr + kk * 2 * (UMMA_N * 16), UMMA_N * 16, 128);
#ifdef SRP_DBG_F16HACK
            // timing experiment only (wrong results): half as many MMAs of twice the depth on the same bytes
            if (kk < KD / 16) umma_f16(tmem_base + as * UMMA_N + g * GRP_N, ad, bd, IDESC & ~((7u << 7) | (7u << 10)), kk > 0 ? 1u : 0u);
#else
            umma_tf32(tmem_base + as * UMMA_N + g * GRP_N, ad, bd, IDESC, kk > 0 ? 1u : 0u);
#endif
          }
          if (g == NGRP - 1) umma_commit(&empty_bar[s]);     // smem stage may be refilled once these MMAs retire
          umma_commit(&acc_full[as * NGRP + g]);             // this group's accumulator is ready for its epilogue warps
        }
        __syncwarp();
      }
    }
    } else if (warp == COPY_WARP && lane == 0) {
    // ===================== spectrum tiles: one bulk copy per bin ==============================
      const unsigned char *src = Bp + (size_t)blockIdx.y * B_BYTES;
      for (int k = 0; k < K; ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(&empty_bar[s], ((k / STAGES) - 1) & 1);
        const uint32_t bar = smem_u32(&full_bar[s]);
#ifdef SRP_DBG_NO_COPY
        mbar_arrive(&full_bar[s]); continue;
#endif
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int pt = threadIdx.x - PROD_WARP0 * 32;          // 0 .. 255
    if (pt < PROD_ACTIVE) {
    const int row = pt % TILE_D, m0 = (pt / TILE_D) * MH;   // this thread's direction and first microphone
    // the phasor is carried as (cos, -sin) = the two steering coefficients themselves (a = cos - j sin)
    float cs[MH], ns[MH], rc[MH], rs[MH];
#pragma unroll
    for (int m = 0; m < MH; ++m) sincospif(two_f0 * tau_s[row * MM + m0 + m], &rs[m], &rc[m]);   // rotation per bin
    const int a_row = (row >> 3) * 128 + (row & 7) * 16;
    for (int k = 0; k < K; ++k) {
      const int s = k % STAGES;
      if ((k % RESYNC) == 0) {
        const float fk2 = two_f0 * (float)k;
#pragma unroll
        for (int m = 0; m < MH; ++m) { float sv; sincospif(fk2 * tau_s[row * MM + m0 + m], &sv, &cs[m]); ns[m] = -sv; }
      }
      if (k >= STAGES) mbar_wait(&empty_bar[s], ((k / STAGES) - 1) & 1);
      unsigned char *At = tiles + s * STAGE_BYTES;
#ifdef SRP_DBG_PROD_NONE
      if (k >= 0) { mbar_arrive(&full_bar[s]); continue; }
#endif
      // steering row: [cos_0 .. cos_{M-1} | -sin_0 .. -sin_{M-1}], rounded to tf32 by adding half an ulp of the
      // 10-bit mantissa to the bit pattern (|x| <= 1: no overflow) -- the MMA drops the low 13 bits
#define RB(x) (__float_as_uint(x) + 0x1000u)
#pragma unroll
      for (int q = 0; q < MH / 4; ++q) {
        *reinterpret_cast<uint4 *>(At + (m0 / 4 + q) * (TILE_D * 16) + a_row) = make_uint4(RB(cs[4 * q]), RB(cs[4 * q + 1]), RB(cs[4 * q + 2]), RB(cs[4 * q + 3]));
        *reinterpret_cast<uint4 *>(At + (MM / 4 + m0 / 4 + q) * (TILE_D * 16) + a_row) = make_uint4(RB(ns[4 * q]), RB(ns[4 * q + 1]), RB(ns[4 * q + 2]), RB(ns[4 * q + 3]));
      }
#undef RB
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (tensor core)
      mbar_arrive(&full_bar[s]);
      // advance the phasors to the next bin: (c, n) <- (c rc + n rs, n rc - c rs)
#ifndef SRP_DBG_PROD_LIGHT
#pragma unroll
      for (int m = 0; m < MH; ++m) {
        const float c = cs[m], nv = ns[m];
        cs[m] = fmaf(c, rc[m], nv * rs[m]);
        ns[m] = fmaf(nv, rc[m], -c * rs[m]);
      }
#endif
    }
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
  const size_t smem = (size_t)STAGES * (TILE_D + UMMA_N) * KD * 4 + (2 * STAGES + 2 * ACC_STAGES * NGRP) * 8 + 16 + (size_t)TILE_D * MM * 4 + 128;
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
