// mcspp_args.cuh -- kernel argument block shared by mcspp.cu and mcspp_fast.cu
#pragma once
#include "common.cuh"
#include "perbin.cuh"

namespace ds {

struct McsppArgs {
  double *state;            // [S][NE][K]
  const double2 *a0;        // [M][K] or null
  const void *X;            // [S][T][M][K] float2 or double2
  float2 *Yout;             // [S][T][K] or null
  double *tp, *txi, *tgamma, *tq, *tG;   // taps [S][T][K]
  double2 *tw_mvdr, *tw_pmwf;            // taps [S][T][M][K]
  double *tAinv;                         // [S][K][M][M] last frame only
  int S, K, T, frm_cnt, ell, k_first, apply_gain;
  double alpha, alpha_d, eps, q_min, q_max, p_min, p_max, snr_min, snr_max, Gmin, logGmin;
  McraConst mc;
};

// state blob: [S][NE][K] float64, element order PyyR[NP] PvvR[NP] mcra[5] PyyI[NQ] PvvI[NQ]
template <int M> __host__ __device__ constexpr int mcspp_state_elems() { return 2 * M * M + 5; }

int launch_mcspp_fast(int M, const McsppArgs &a, cudaStream_t st);

}  // namespace ds
