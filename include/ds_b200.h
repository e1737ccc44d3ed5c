/*
 * ds_b200.h -- C ABI of libds_b200.so: B200 (sm_100a) kernels for the
 * DistantSpeech multichannel enhancement hot path.
 *
 * The reference (wangwei2009/DistantSpeech) is pure Python/NumPy and has no
 * FFI layer; the boundary a maintainer binds is its Python call surface
 * (SURVEY.md 8b).  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  INTEGRATION.md shows
 * the ctypes stub that goes on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _h;
 *   - no allocation, no synchronisation and no host<->device copies inside a
 *     *_run call; kernels are enqueued on the cudaStream_t passed as `stream`
 *     (a `void*`, so this header needs no CUDA include);
 *   - state lives in caller-allocated blobs whose size is returned by the
 *     matching *_state_bytes(); a zero-filled blob is the reset state;
 *   - return value: DS_OK (0) or a negative DS_E* code; ds_last_error() gives
 *     a thread-local human-readable message for the last failure;
 *   - a state blob is not thread-safe; distinct blobs are independent;
 *   - complex arrays are interleaved (re, im); "c64" = 2 x float, "c128" =
 *     2 x double;
 *   - K = n_fft/2 + 1 bins, T frames, M mics/channels, S independent streams.
 *
 * Device layouts (chosen for coalescing: the bin index is innermost because
 * kernels map threads to bins; the sample index is innermost for audio):
 *   audio      x [S][M][N]   float32     (stream, mic, sample)
 *   spectrum   X [S][T][M][K] c64        (stream, frame, mic, bin)
 *   per-bin    p [S][T][K]   float64
 */
#ifndef DS_B200_H
#define DS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_VERSION 100

enum {
  DS_OK = 0,
  DS_EINVAL = -1,       /* bad shape / null pointer / inconsistent params        */
  DS_EUNSUPPORTED = -2, /* n_fft, M ... outside the compiled template set         */
  DS_ECUDA = -3,        /* CUDA runtime error (launch failure, wrong device ...)  */
  DS_ESTATE = -4        /* state blob too small                                   */
};

/* ---- library ---------------------------------------------------------- */
int ds_version(void);
const char *ds_last_error(void);
/* Builds the twiddle tables for every supported n_fft on the current device.
 * Idempotent. Call once before the first *_run (the *_run calls also do it
 * lazily, but that allocates on first use). */
int ds_init(void);
/* SM count and compute capability of the current device. */
int ds_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* Strided host <-> device copy on `stream` (cudaMemcpy2DAsync): `height` rows of `width` bytes; kind 0 = host to device,
 * 1 = device to host.  The host-buffer pipelines move a time slice of every stream with it (no reference counterpart:
 * the reference has no device).                                                                                  */
int ds_memcpy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, int kind,
                      void *stream);
/* fp64-pipe microbenchmark (8 independent DFMA chains per thread, 2 x 1024 threads per SM): launches on `stream`
 * and returns the floating-point operations executed (0 on error); the caller times it with CUDA events.  It is the
 * measured denominator of roofline.fp64_pipe in bench.py (no reference counterpart: measurement infrastructure). */
double ds_fp64_peak_run(int iters, double *scratch, void *stream);

/* ---- STFT / ISTFT  (transform/transform.py) --------------------------- */
enum {
  DS_STFT_STREAMING = 0, /* Transform.stft  :430-453: `history` (n_fft-hop samples per
                            channel) is prepended, T = N/hop, history is updated   */
  DS_STFT_CENTER = 1,    /* stft(center=True, pad_mode="reflect") :205-206, T=1+N/hop */
  DS_STFT_PLAIN = 2      /* stft(center=False)  :209, T = 1+(N-n_fft)/hop           */
};

typedef struct ds_stft_params {
  int32_t n_fft;     /* power of two, 128..2048                                   */
  int32_t hop;       /* 1..n_fft                                                  */
  int32_t n_streams; /* S                                                         */
  int32_t n_ch;      /* M (channels per stream)                                   */
  int32_t n_samples; /* N samples per channel in this call                        */
  int32_t mode;      /* DS_STFT_*                                                 */
  int32_t fft_fp64;  /* 0: fp32 FFT (default), 1: fp64 FFT                        */
  int32_t out_c128;  /* 0: X is complex64 (reference default dtype), 1: complex128 */
} ds_stft_params;

/* Number of frames a call produces (negative DS_E* on bad params). */
int ds_stft_num_frames(const ds_stft_params *p);

/* replaces transform.stft (transform.py:10-221) and Transform.stft (:430-453).
 *   window  [n_fft] float64 (already centre-padded to n_fft)
 *   history [S][M][n_fft-hop] float32, in/out, DS_STFT_STREAMING only (else NULL)
 *   x       [S][M][N] float32
 *   X       [S][T][M][K] c64 out (rounded to complex64 like :212) or c128        */
int ds_stft_run(const ds_stft_params *p, const double *window, float *history,
                const float *x, void *X, void *stream);
/* Same transform fed with int16 PCM x [S][M][N]: replaces load_audio's scaling
 * float32(pcm) / 32767 (beamformer/utils.py:182-187) followed by Transform.stft
 * (transform.py:430-453) -- the scaling runs inside the analysis kernel, the
 * float waveform never exists in memory; `history` stays float32.                */
int ds_stft_pcm16_run(const ds_stft_params *p, const double *window, float *history,
                      const int16_t *x_pcm, void *X, void *stream);

typedef struct ds_istft_params {
  int32_t n_fft;
  int32_t hop;
  int32_t n_streams;
  int32_t n_ch;
  int32_t n_frames; /* T                                                          */
  int32_t mode;     /* DS_STFT_STREAMING: Transform.istft (:455-481): out has hop*T
                       samples, `tail` carries previous_output, scaled by `scale`;
                       DS_STFT_PLAIN: istft(center=False): out has n_fft+hop*(T-1)
                       samples, no tail, no scaling                               */
  int32_t fft_fp64;
  int32_t in_c128; /* 0: Y is complex64, 1: complex128                            */
  double scale; /* Transform.istft: hop / sum(window^2) (:479)                    */
} ds_istft_params;

/* replaces transform.istft (transform.py:237-404, float32 overlap-add :224-234,
 * no window-sum normalisation) and Transform.istft (:455-481).
 *   Y    [S][T][C][K] c64/c128 tail [S][C][n_fft-hop] float32 in/out (streaming)
 *   y    [S][C][n_out] float32                                                   */
int ds_istft_run(const ds_istft_params *p, const double *window, float *tail,
                 const void *Y, float *y, void *stream);
/* Same synthesis writing int16 PCM y [S][C][n_out]: Transform.istft (:455-481)
 * followed by save_audio's (audio * 32767).astype(int16) (beamformer/utils.py:190-196,
 * product in double, truncation toward zero; out-of-range values saturate) fused
 * into the synthesis kernel's store.                                             */
int ds_istft_pcm16_run(const ds_istft_params *p, const double *window, float *tail,
                       const void *Y, int16_t *y_pcm, void *stream);
/* Same synthesis reading Y in the layout [S][C][K][frame_pitch] (frame index innermost, frame_pitch >= T) -- what the
 * tensor-core multi-beam kernel (ds_multibeam_tc_run) writes.                                                    */
int ds_istft_frames_inner_run(const ds_istft_params *p, const double *window, float *tail,
                              const void *Y, long long frame_pitch, float *y, void *stream);

/* ---- fixed beamformer (beamformer/fixedbeamformer.py) ----------------- */
typedef struct ds_fixedbf_params {
  int32_t n_fft;
  int32_t hop; /* must be n_fft/2 or n_fft/4                                      */
  int32_t n_streams;
  int32_t n_mics;
  int32_t n_samples; /* multiple of hop                                           */
  int32_t n_beams;   /* B >= 1 look directions evaluated at once                  */
  int32_t reserved0;
  int32_t reserved1;
  double scale; /* hop / sum(window^2)                                            */
} ds_fixedbf_params;

size_t ds_fixedbf_state_bytes(const ds_fixedbf_params *p);
/* replaces FixedBeamformer.process (fixedbeamformer.py:167-207): Transform.stft
 * -> Y[k,t] = sum_m conj(W[k,m]) X[k,t,m] (:163) -> Transform.istft, fused in
 * one kernel (the spectrum never goes to HBM).
 *   W     [B][K][M] c64 weights (as returned by compute_weights, NOT conjugated)
 *   state ds_fixedbf_state_bytes() blob: input history + output tail (zero = reset)
 *   x     [S][M][N] float32        y [S][B][N] float32                           */
int ds_fixedbf_run(const ds_fixedbf_params *p, const double *window, const void *W,
                   void *state, const float *x, float *y, void *stream);

/* ---- per-bin weight helpers (beamformer/beamformer.py module functions) -- */
/* replaces compute_mvdr_weight (beamformer.py:133-155): w = R^-1 a / (a^H R^-1 a).
 *   steer [bins][M] c128   Rvv_inv [bins][M][M] c128   w_out [bins][M] c128, M <= 16 */
int ds_mvdr_weight_run(int n_bins, int n_mics, const void *steer, const void *Rvv_inv,
                       void *w_out, void *stream);
/* replaces compute_pmwf_weight (beamformer.py:100-130, mcspp_base.py:220-240):
 * w = (Rvv_inv Rxx) u_1 / (beta + xi).  xi [bins] float64, Rxx/Rvv_inv [bins][M][M] c128 */
int ds_pmwf_weight_run(int n_bins, int n_mics, const double *xi, const void *Rxx,
                       const void *Rvv_inv, double beta, void *w_out, void *stream);
/* replaces process_freframe (fixedbeamformer.py:147-165): Y[s,t,k] = sum_m conj(W[k,m]) X[s,t,m,k].
 *   X [S][T][M][K] c64 or c128   W [K][M] c128   Y [S][T][K] c128                 */
int ds_apply_weights_run(int n_streams, int n_frames, int n_mics, int n_bins, const void *X,
                         int x_is_c128, const void *W, void *Y, void *stream);

/* |X|^2 the two ways the reference writes it.  via_abs = 0: np.real(X * np.conj(X)) (FDGSC.py:288,
 * GSC.py:286), re*re + im*im with each product rounded; via_abs = 1: np.abs(X) ** 2 (mcra.py:29-30,
 * what NoiseEstimationMCRA does to complex input), hypot then squared.
 * X: n complex values, c64 or c128; out: n float64.                                              */
int ds_power_run(long long n, const void *X, int x_is_c128, int via_abs, double *out, void *stream);
/* Y[i] * g[i] in complex128, g = G or sqrt(G) (FDGSC.py:292-294): spectral gain before Transform.istft.
 * Yin c64 or c128 (n values), G float64, Yout c128 (may alias Yin when it is c128).              */
int ds_spectral_gain_run(long long n, const void *Yin, int y_is_c128, const double *G, int take_sqrt,
                         void *Yout, void *stream);

/* replaces McSppBase.compute_omlsa_weight (mcspp_base.py:140-155):
 * G = clip((xi/(1+xi))^p Gmin^(1-p), Gmin, 1), G[:2] = 0 per row of n_bins.       */
int ds_omlsa_gain_run(int n_rows, int n_bins, const double *xi, const double *p, double Gmin,
                      double *G, double *G_H1, void *stream);

/* ---- MCRA noise estimator (noise_estimation/mcra.py) ------------------ */
typedef struct ds_mcra_params {
  int32_t n_bins;    /* K                                                         */
  int32_t n_streams; /* S                                                         */
  int32_t n_frames;  /* T                                                         */
  int32_t L;         /* minimum-search window (15)                      mcra.py:25 */
  int32_t frm_cnt;   /* frames already processed (host-tracked)          :72       */
  int32_t ell;       /* window counter at entry (1 at reset)             Base :18  */
  int32_t reserved0;
  int32_t reserved1;
  double alpha_d, alpha_s, delta_s, alpha_p, p_min, p_max; /* .95 .8 5 .2 1e-3 .999 */
} ds_mcra_params;

size_t ds_mcra_state_bytes(const ds_mcra_params *p); /* [S][5][K] float64: S,Smin,Stmp,p,lambda_d */
/* replaces NoiseEstimationMCRA.estimation (mcra.py:27-77) applied to T frames.
 *   Ypow [S][T][K] float64 power spectrum    lambda_out/p_out [S][T][K] or NULL
 * The caller advances frm_cnt/ell with ds_mcra_advance().                        */
int ds_mcra_run(const ds_mcra_params *p, void *state, const double *Ypow,
                double *lambda_out, double *p_out, void *stream);
/* Host-side helper: frm_cnt/ell after n_frames more frames (mcra.py:52-56,72-74). */
void ds_mcra_advance(int32_t L, int32_t n_frames, int32_t *frm_cnt, int32_t *ell);

/* ---- McSppBase estimator + MVDR/OMLSA chain --------------------------- */
typedef struct ds_mcspp_params {
  int32_t n_fft;
  int32_t n_streams;
  int32_t n_mics;   /* 2..8                                                       */
  int32_t n_frames; /* T                                                          */
  int32_t frm_cnt;  /* frames already processed (host-tracked)                    */
  int32_t ell;      /* inner MCRA window counter at entry                         */
  int32_t mcra_L;   /* 15                                         mcspp_base.py:77 */
  int32_t full_state; /* 1: track complex Phi_yy/Phi_vv, all K bins and the PMWF
                         weights (every public attribute of McSppBase);
                         0: output-only -- real parts and bins 2..K-1 only, which is
                         everything the chain output depends on (:278-284 use .real,
                         compute_omlsa_weight zeroes G[:2])                       */
  double alpha, alpha_d;              /* .92 .92                          :38-40  */
  double diag_eps;                    /* 1e-6                             :74     */
  double q_min, q_max, p_min, p_max;  /* .01 .99 .01 .99                  :120,290 */
  double snr_min, snr_max;            /* 1e-6 1e6  (xi, gamma clip)       :286-287 */
  double Gmin;                        /* 0.0631                           :140    */
  double mcra_alpha_d, mcra_alpha_s, mcra_delta_s, mcra_alpha_p, mcra_p_min, mcra_p_max;
} ds_mcspp_params;

/* Fills every constant of *p with the reference defaults. */
void ds_mcspp_default_params(ds_mcspp_params *p, int n_fft, int n_streams, int n_mics, int n_frames);
size_t ds_mcspp_state_bytes(const ds_mcspp_params *p);

typedef struct ds_mcspp_taps { /* optional per-frame outputs, any may be NULL     */
  double *p;      /* [S][T][K]    posterior SPP                                   */
  double *xi;     /* [S][T][K]                                                    */
  double *gamma;  /* [S][T][K]                                                    */
  double *q;      /* [S][T][K]                                                    */
  double *G;      /* [S][T][K]    OMLSA gain (compute_omlsa_weight)               */
  void *w_mvdr;   /* [S][T][M][K] c128 MVDR weights                               */
  void *w_pmwf;   /* [S][T][M][K] c128 PMWF weights (full_state only)             */
  double *Phi_vv_inv_last; /* [S][K][M][M] float64, inverse used by the LAST frame */
} ds_mcspp_taps;

/* replaces, per frame: McSppBase.estimation (mcspp_base.py:262-297),
 * compute_mvdr_weight(a0, Phi_vv_inv) (beamformer.py:133-155),
 * McSppBase.compute_omlsa_weight (:140-155) and the weight/gain apply
 * Y = (w^H y) G  (example/mvdr.ipynb, GSC.py:286).
 *   a0   [M][K] c128 steering vectors (or NULL: estimator only, Yout unused)
 *   X    [S][T][M][K] c64 (x_is_c128 = 0) or c128 (x_is_c128 = 1)
 *   Yout [S][T][K] c64 beamformed+postfiltered spectrum (or NULL)
 *   apply_gain: multiply by the OMLSA gain G (1) or output the plain MVDR (0)
 * Kernel selection: full_state = 0 with a0, Yout, c64 input and either no tap or the p tap alone runs
 * the output-only kernel (the headline path; with the p tap it visits every bin so that the mask of
 * the mask-based beamformers is complete); anything else runs the full-state kernel.             */
int ds_mcspp_run(const ds_mcspp_params *p, void *state, const void *a0, const void *X,
                 int x_is_c128, void *Yout, int apply_gain, const ds_mcspp_taps *taps,
                 void *stream);
/* state export for the Python attribute views: copies one named field of every
 * stream into `out`. field: 0 Phi_yy, 1 Phi_vv (c128 [S][K][M][M]),
 * 2 mcra block ([S][5][K] float64).                                              */
int ds_mcspp_export(const ds_mcspp_params *p, const void *state, int field, void *out, void *stream);

/* ---- McSpp with the McCDR prior (noise_estimation/mcspp.py, mccdr.py) ------ */
typedef struct ds_mcspp_cdr_params {
  int32_t n_fft;
  int32_t n_streams;
  int32_t n_mics;   /* 4..8.  4 is what the reference runs as shipped (McSpp builds McCDR
                       with its default 4 channels, mcspp.py:54, and raises IndexError
                       above); 5..8 follow the reference with McCDR(nfft, channels=M)  */
  int32_t n_frames;
  int32_t frm_cnt;  /* frames already processed (host-tracked; McSpp.frm_cnt)        */
  int32_t ell;      /* window counter of McCDR's MCRA at entry                       */
  int32_t mcra_L;   /* 65                                              mccdr.py:56   */
  int32_t cdr_only; /* bit 0: only McCDR.estimation (taps->cdr), state of the prior only;
                       bit 1: McSpp.estimation(repeat=True): second estimation_core pass
                       on the updated noise covariance (mcspp.py:282-284)             */
  int32_t band_lo_bin, band_hi_bin; /* int(500 nfft/16000), int(2000 nfft/16000) :266-267 */
  int32_t init_frames;              /* 10: Phi_vv = Phi_yy, q = q_init     :276-278  */
  int32_t fallback_loaded_frames;   /* 5: loaded fallback inverse          :224-227  */
  double alpha, alpha_d;            /* .92 .92                             :64-65    */
  double alpha_cdr;                 /* .9                               mccdr.py:126 */
  double load_min, load_max;        /* 1e-4 1e-1                           :262-263  */
  double snr_min, snr_max;          /* 1e-6 1e8  (xi, gamma clip)          :229,236  */
  double pmwf_beta;                 /* 10                                  :286      */
  double q_init;                    /* .99                                 :278      */
  double mcra_alpha_d, mcra_alpha_s, mcra_delta_s, mcra_alpha_p, mcra_p_min, mcra_p_max;
} ds_mcspp_cdr_params;
void ds_mcspp_cdr_default_params(ds_mcspp_cdr_params *p, int n_fft, int n_streams, int n_mics, int n_frames);
size_t ds_mcspp_cdr_state_bytes(const ds_mcspp_cdr_params *p);     /* zero-filled = reset */
size_t ds_mcspp_cdr_workspace_bytes(const ds_mcspp_cdr_params *p); /* scratch, per call   */

typedef struct ds_mcspp_cdr_taps { /* optional per-frame outputs, any may be NULL  */
  double *p;     /* [S][T][K] posterior SPP (the return value of McSpp.estimation) */
  double *xi;    /* [S][T][K]                                                      */
  double *gamma; /* [S][T][K]                                                      */
  double *q;     /* [S][T][K] prior speech absence probability used by compute_p   */
  double *cdr;   /* [S][T][K] McCDR.estimation return value sqrt(CDR^2 p_mcra)     */
  void *w;       /* [S][T][M][K] c128 PMWF weights (beta = pmwf_beta)              */
} ds_mcspp_cdr_taps;

/* replaces, per frame, McSpp.estimation (mcspp.py:248-305) including McCDR.estimation
 * (mccdr.py:164-177) and compute_pmwf_weight (mcspp_base.py:220-240).
 *   Fn   [K] float64: diffuse coherence Fvv[:, 1, 2] of the M-mic circular r = 0.032 array
 *        McCDR owns (mccdr.py:58-59, gen_noise_msc.py)
 *   X    [S][T][M][K] c64 (x_is_c128 = 0) or c128
 *   Yout [S][T][K] c64 = w^H y (or NULL)                                            */
int ds_mcspp_cdr_run(const ds_mcspp_cdr_params *p, void *state, void *workspace, const double *Fn,
                     const void *X, int x_is_c128, void *Yout, const ds_mcspp_cdr_taps *taps, void *stream);
/* state views. field: 0 Phi_yy, 1 Phi_vv, 2 Phi_vv_inv, 3 Phi_xx (c128 [S][K][M][M]; 2-3 as of the
 * last frame); 4 w [S][2M][K] (re[M], im[M]); 5 [S][4][K] xi gamma q cdr^2; 6 [S][M + M(M-1)][K]
 * Pxii[M] Re Pxij[NQ] Im Pxij[NQ] (NQ = M(M-1)/2 pairs, i < j row-major); 7 [S][6][K] MCRA S Smin Stmp p
 * lambda_d, posterior p                                                             */
int ds_mcspp_cdr_export(const ds_mcspp_cdr_params *p, const void *state, int field, void *out, void *stream);

/* ---- multi-beam fixed / superdirective weighting on the tensor cores (beamformer/fixedbeamformer.py) ---------- */
/* replaces, for n_beams look directions at once, the per-frame weighting of FixedBeamformer.process_freframe
 * (fixedbeamformer.py:147-165, einsum 'ij,ij->i' of conj(W) and X; weights from compute_weights :109-145):
 *   Y[s][b][k][t] = sum_m conj(W[b][k][m]) X[s][t][m][k]
 * as per-bin real GEMMs [128 beams x 6M] . [6M x 2*64 frames] on tcgen05 (kind::tf32, operands split into tf32 head +
 * tail so the result is fp32-accurate), accumulators in tensor memory, operand tiles streamed with cp.async.bulk.
 *   W [n_beams][K][M] c64, X [S][T][M][K] c64 (ds_stft_run's layout), n_mics in {4, 8, 16}
 *   Y [S][n_beams][K][frame_pitch] c64 (frame index innermost; feed it to ds_istft_frames_inner_run)
 *   workspace: scratch for the packed operand tiles, 128-byte aligned; sizes from ds_multibeam_tc_layout.        */
int ds_multibeam_tc_layout(int n_streams, int n_frames, int n_mics, int n_bins, int n_beams, int *frame_pitch,
                           size_t *workspace_bytes, size_t *y_bytes);
int ds_multibeam_tc_run(int n_streams, int n_frames, int n_mics, int n_bins, int n_beams, const void *W, const void *X,
                        void *workspace, void *Y, void *stream);

/* ---- adaptive (RLS) WPE dereverberation (dereverberation/awpe.py) ----------------------------- */
/* state blob [S][NE][K] float64 (NE from ds_wpe_state_bytes): W re/im [C][C L], P re/im [C L][C L], the last
 * delay + filter_len - 1 input frames re/im [.][C] (most recent first), var.  A fresh filter has P = 1e-3 I
 * (awpe.py:66-71) and everything else zero -- the caller writes the diagonal.                                  */
size_t ds_wpe_state_bytes(int n_streams, int n_bins, int n_ch, int filter_len, int delay);
/* replaces the arithmetic of Wpe.update (awpe.py:152-187) for T frames per call: X [S][T][C][K] c64 (x_is_c128 = 0)
 * or c128, the streaming STFT of the observed signals; Err [S][T][C][K] c128 receives the prior error d - W^H X of
 * every channel (the dereverberated spectrum).  delay = D frames (awpe.py:72: D hops in the time domain).      */
int ds_wpe_run(int n_streams, int n_bins, int n_frames, int n_ch, int filter_len, int delay, double forgetting_factor,
               double alpha_var, void *state, const void *X, int x_is_c128, void *Err, void *stream);

/* ---- McMcra + frequency-domain GSC (noise_estimation/mc_mcra.py, beamformer/GSC.py) -------- */
typedef struct ds_gsc_params {
  int32_t n_fft;
  int32_t n_streams;
  int32_t n_mics;      /* 2..8                                                              */
  int32_t n_frames;
  int32_t frm_cnt;     /* frames already processed (McMcra.frm_cnt, host-tracked)           */
  int32_t method;      /* GSC.process `method`: 0 = pass channel 0 through (GSC.py:242), else GSC */
  int32_t init_frames; /* 5: Phi_vv = Phi_yy                              mc_mcra.py:186-187 */
  int32_t reserved;
  double alpha, alpha_d;            /* .92 .95                             mc_mcra.py:34-36  */
  double diag_eps;                  /* 1e-6                                :191              */
  double psi_0;                     /* 100 (psi_0 = psi_tilde_0)           :62-63            */
  double q_min, q_max, p_min, p_max;/* .01 .99 .01 .99                     :89, :218         */
  double snr_min, snr_max;          /* 1e-6 1e6                            :194, :199        */
  double Gmin;                      /* 0.0631                              :152              */
  double mu;                        /* 0.01 NLMS step                      GSC.py:207        */
} ds_gsc_params;
void ds_gsc_default_params(ds_gsc_params *p, int n_fft, int n_streams, int n_mics, int n_frames);
/* state [S][NE][K] float64, zero = reset: Phi_yy, Phi_vv (packed real upper triangles, M(M+1)/2 each),
 * Re G[M-1], Im G[M-1] (noise-canceller weights), then p q xi gamma G of the last frame               */
size_t ds_gsc_state_bytes(const ds_gsc_params *p);
typedef struct ds_gsc_taps { /* optional per-frame outputs [S][T][K], any may be NULL */
  double *p, *G, *xi, *gamma, *q;
} ds_gsc_taps;
/* replaces, per frame, McMcra.estimation (mc_mcra.py:179-221) and -- when `a` and `Yout` are given -- the
 * per-bin loop of GSC.process (GSC.py:231-286): fixed beam a/(a^H a), Griffiths-Jim blocking matrix, NLMS
 * noise canceller gated by 1 - p, McMcra gain as postfilter.
 *   a    [M][K] c128 propagation vectors exp(-j w_k tau_m) (or NULL: estimator only)
 *   X    [S][T][M][K] c64 / c128     Yout [S][T][K] c64 (or NULL)                                        */
int ds_gsc_run(const ds_gsc_params *p, void *state, const void *a, const void *X, int x_is_c128, void *Yout,
               const ds_gsc_taps *taps, void *stream);

/* ---- STFT-domain NLMS filters (adaptivefilter/SubbandLMS.py, SubbandLmsMc.py) and SubbandGSC helpers -- */
typedef struct ds_subband_nlms_params {
  int32_t n_bins;      /* K = num_bands / 2 + 1                                                   */
  int32_t n_streams;   /* S                                                                       */
  int32_t n_filters;   /* F independent filters per stream that share the input X (SubbandGSC: the M
                          blocking filters all take the fixed beam as input, SubbandGSC.py:220-226) */
  int32_t n_frames;    /* T                                                                       */
  int32_t n_ch;        /* C input channels per filter (SubbandLMS: 1, SubbandLmsMc: channel)      */
  int32_t filter_len;  /* frame taps per bin (2 in SubbandGSC)                                    */
  int32_t one_minus_p; /* 1: gate with 1 - p (the canceller, SubbandGSC.py:236)                   */
  int32_t plain_lms;   /* 1: normalization=False, grad = buf conj(err) without the power term (:77) */
  double mu;    /* step: W += 2 mu p grad                            SubbandAF.py:84-87           */
  double alpha; /* power smoothing (0.9 default, 0.8 canceller)      SubbandLMS.py:70-74          */
  double eps;   /* 1e-4 regulariser (the `alpha` argument of update) SubbandLMS.py:75            */
} ds_subband_nlms_params;
size_t ds_subband_nlms_state_bytes(const ds_subband_nlms_params *p);   /* zero = fresh filters    */
/* replaces SubbandLMS.update (SubbandLMS.py:28-84) / SubbandLmsMc.update (SubbandLmsMc.py:144-191) over T frames,
 * spectra in, error spectra out (the Transform analysis / synthesis around them are ds_stft_run / ds_istft_run):
 *   X [S][T][C][K] c64 input spectra   D [S][T][F][K] c64 desired spectra   prob [S][T][K] float64 or NULL (p = 1)
 *   Err [S][T][F][K] c128                                                                               */
int ds_subband_nlms_run(const ds_subband_nlms_params *p, void *state, const void *X, const void *D, const double *prob,
                        void *Err, void *stream);
/* replaces SubbandRLS.update (adaptivefilter/SubbandRLS.py:44-71) over T frames: per-bin RLS, filter_len (1..4) frame
 * taps, one input channel.  state [S][NE][K] float64: Re W[L] Im W[L] Re buf[L] Im buf[L] Re P[L][L] Im P[L][L]; the caller
 * initialises P = I / 1e-3 (:38-40), everything else zero.  X, D [S][T][K] c64, Err [S][T][K] c128.                     */
size_t ds_subband_rls_state_bytes(int n_streams, int n_bins, int filter_len);
int ds_subband_rls_run(int n_streams, int n_bins, int n_frames, int filter_len, double mu, double forgetting_factor,
                       void *state, const void *X, const void *D, void *Err, void *stream);
/* FilterDcNotch16.filter_dc_notch16 (adaptivefilter/feature.py:37-49) in place on x [S][C][n_samples] float32,
 * memories mem [S][C][2] float64 in/out (zero = fresh filter).                                            */
int ds_dcnotch_run(int n_streams, int n_ch, int n_samples, double radius, double *mem, float *x, void *stream);
/* np.mean(x, axis=channel) (the fixed beamformer of the GSC pipelines, SubbandGSC.py:143, FDGSC.py:138):
 * x [S][C][n_samples] float64 -> out [S][n_samples] float64.                                              */
int ds_channel_mean_run(int n_streams, int n_ch, long long n_samples, const double *x, double *out, void *stream);

/* ---- constrained frequency-domain adaptive filter (adaptivefilter/FastFreqLms.py) and TDGSC helpers ---- */
typedef struct ds_fdaf_params {
  int32_t frame_len;    /* filter_len = hop_len; compiled for 256 (n_fft 512)           FastFreqLms.py:62-68 */
  int32_t n_streams;
  int32_t n_ch;         /* input channels, 1..7 (TDGSC: M - 1)                                               */
  int32_t n_samples;    /* multiple of frame_len                                                             */
  int32_t fir_truncate; /* taps zeroed at both ends after each update (:238-243); < 0 = None                 */
  int32_t non_causal;   /* 1: desired signal delayed by frame_len / 2 (:80-81, :156-157)                     */
  int32_t one_minus_p;  /* 1: the gate is 1 - prob (TDGSC.py:157)                                            */
  int32_t reserved;
  double mu;            /* 0.01                                                                   :55        */
  double alpha;         /* 0.9 power smoothing                                                    :58        */
} ds_fdaf_params;
size_t ds_fdaf_state_bytes(int n_streams, int n_ch);          /* float32 state, zero = fresh filter */
/* replaces FastFreqLms.update (FastFreqLms.py:203-245, base class: gradient constraint, factor 2, optional tap
 * truncation) over all blocks of a call:
 *   x [S][C][N] float32 inputs, d [S][N] float32 desired, prob [S][N/frame_len][257] float64 per-bin gate or NULL (1),
 *   e [S][N] float32 error output                                                                             */
int ds_fdaf_run(const ds_fdaf_params *p, void *state, const float *x, const float *d, const double *prob, float *e,
                void *stream);
/* the fixed blocking matrix of TDGSC (TDGSC.py:77-81): out[c] = x[c] - x[c + 1];
 * x [S][C][n_samples] float64 -> out [S][C-1][n_samples] float32                                              */
int ds_adjacent_diff_run(int n_streams, int n_ch, long long n_samples, const double *x, float *out, void *stream);

/* ---- postfilter gains ------------------------------------------------------ */
typedef struct ds_omlsa_multi_params {
  int32_t n_bins, n_streams, n_frames;
  int32_t n_mics;      /* M: beam output + M-1 references, 2..8                     */
  int32_t first_frame; /* 1 until the first frame has been seen (host-tracked) omlsa_multi.py:87 */
  int32_t frm_cnt, ell, mcra_L; /* shared by the M MCRA trackers (15)                */
  int32_t cal_weights; /* compute the OMLSA gain G                           :152   */
  int32_t u_const;     /* 1: u is [S][M-1][K], the same reference powers for every frame
                          (what FDGSC.process(postfilter=True) feeds it, FDGSC.py:285-289) */
  double alpha_d, alpha_s, alpha_xi; /* 0.85, 0.8, 0.921                  :53,70,96 */
  double beta;                       /* 1.47                                   :149  */
  double Gmin, q_min, q_max;         /* 10^-1.2, 1e-6, 0.9999998            :35-50   */
  double mcra_alpha_d, mcra_alpha_s, mcra_delta_s, mcra_alpha_p, mcra_p_min, mcra_p_max;
} ds_omlsa_multi_params;
void ds_omlsa_multi_default_params(ds_omlsa_multi_params *p, int n_bins, int n_streams, int n_frames, int n_mics);
size_t ds_omlsa_multi_state_bytes(const ds_omlsa_multi_params *p);
/* replaces NsOmlsaMulti.estimation (noise_estimation/omlsa_multi.py:73-156) over T frames.
 *   y [S][T][K] beam-output power, u [S][T][M-1][K] (u_const: [S][M-1][K]) reference powers (float64)
 *   G_out / lambda_out / p_out [S][T][K] or NULL                                  */
int ds_omlsa_multi_run(const ds_omlsa_multi_params *p, void *state, const double *y, const double *u,
                       double *G_out, double *lambda_out, double *p_out, void *stream);

size_t ds_zelinski_state_bytes(int n_streams, int n_mics, int n_bins);
/* replaces PostFilter.update_CSD_PSD + getweights (postfilter/postfilter.py:19-84).
 *   Z [S][T][M][K] c128, Fvv [K][M][M] float64 diffuse coherence, W [S][T][K] float64 out
 *   alpha 0.8 (:52), coh_max 0.7 (:68)                                             */
int ds_zelinski_run(int n_streams, int n_frames, int n_mics, int n_bins, double alpha, double coh_max,
                    void *state, const void *Z, const double *Fvv, double *W, void *stream);

/* ---- online MVDR with MCRA-VAD gate (beamformer/adaptivebeamformer.py) ---- */
typedef struct ds_amvdr_params {
  int32_t n_fft;
  int32_t n_streams;
  int32_t n_mics;   /* 2..8                                                       */
  int32_t n_frames; /* T                                                          */
  int32_t frm_cnt;  /* MCRA frames already processed (host-tracked)               */
  int32_t ell;      /* MCRA window counter at entry                               */
  int32_t mcra_L;   /* 15                                               mcra.py:25 */
  int32_t method;   /* AlgorithmList index: 0 src, 1 DS, 2 MVDR, 3 TFGSC      :37 */
  double alpha_y, alpha_v; /* 0.8, 0.9998                                  :65-66 */
  double diag;             /* 1e-6 diagonal loading                        :89    */
  double vad_thr;          /* 0.4: update Rvv where mcra.p < vad_thr       :94    */
  double mcra_alpha_d, mcra_alpha_s, mcra_delta_s, mcra_alpha_p, mcra_p_min, mcra_p_max;
} ds_amvdr_params;

void ds_amvdr_default_params(ds_amvdr_params *p, int n_fft, int n_streams, int n_mics, int n_frames);
size_t ds_amvdr_state_bytes(const ds_amvdr_params *p);
/* replaces the frame/bin loops of adaptivebeamfomer.process
 * (adaptivebeamformer.py:69-120): MCRA on channel 0, Ryy / gated Rvv recursions,
 * Hermitian inverse with diagonal loading, getweights (beamformer.py:306-336) and
 * the weight apply.
 *   a      [M][K] c128 propagation vectors exp(-j w_k tao_m)          (:84)
 *   X      [S][T][M][K] c64        Yout [S][T][K] c64
 *   H_last [S][K][M] c128 weights of the last frame or NULL
 *   p_out  [S][T][K] float64 MCRA speech presence or NULL                        */
int ds_amvdr_run(const ds_amvdr_params *p, void *state, const void *a, const void *X,
                 void *Yout, void *H_last, double *p_out, void *stream);
/* dense view of one state field: 0 Rvv, 1 Rvv_inv, 2 Ryy -> [S][K][M][M] c128    */
int ds_amvdr_export(const ds_amvdr_params *p, const void *state, int field, void *out, void *stream);

/* ---- frequency-domain GSC (beamformer/FDGSC.py) ------------------------- */
typedef struct ds_fdgsc_params {
  int32_t frame_len;  /* 256 (block = filter length; FFT length 512)               */
  int32_t n_streams;
  int32_t n_mics;     /* 2..8                                                      */
  int32_t n_samples;  /* multiple of frame_len                                     */
  int32_t filter_len; /* taps of the time-alignment FIR bank (<= 128)              */
  int32_t frm_cnt;    /* MCRA frames already processed (host-tracked)              */
  int32_t ell;        /* MCRA window counter at entry                              */
  int32_t mcra_L;     /* 60                                          FDGSC.py:99   */
  int32_t fp64;       /* 0: fp32 filters / FFTs, 1: fp64                           */
  int32_t dc_notch;   /* 1: apply FilterDcNotch16 in place first     FDGSC.py:211  */
  int32_t reserved;
  int32_t reserved2;
  double mu_bm, mu_aic;   /* 0.1, 0.1                                  FDGSC.py:66,77 */
  double alpha;           /* 0.9 power smoothing            FastFreqLms.py:158      */
  double notch_radius;    /* 0.98                                     FDGSC.py:115  */
  double maxnorm;         /* 0.003                                   gsc_aic.py:85  */
  double delta;           /* 0.001 tap bound                         gsc_bm.py:51   */
  double mcra_alpha_d, mcra_alpha_s, mcra_delta_s, mcra_alpha_p, mcra_p_min, mcra_p_max;
} ds_fdgsc_params;

void ds_fdgsc_default_params(ds_fdgsc_params *p, int n_streams, int n_mics, int n_samples, int filter_len);
size_t ds_fdgsc_state_bytes(const ds_fdgsc_params *p);
/* replaces FDGSC.process(x, postfilter=False, dc_notch=...) (FDGSC.py:201-317).
 *   delay_filter [M][filter_len] float64: fractional_delay_filter_bank of TimeAlignment
 *                (fixedbeamformer.py:66-72, multirate.py:4-51), host precompute
 *   window       [512] float64 sqrt-Hann (Transform of FDGSC.py:104)
 *   x            [S][M][N] float32, IN/OUT: overwritten with the DC-notched signal (FDGSC.py:213)
 *   y            [S][N] float32 enhanced output (tuple element 0)
 *   bm_out       [S][M][N] blocking-matrix outputs (element 4) or NULL
 *   fix_out      [S][N] fixed beamformer output (element 2) or NULL
 *   p_out        [S][N/256][257] float64 adaptation-control SPP (element 1) or NULL   */
int ds_fdgsc_run(const ds_fdgsc_params *p, const double *delay_filter, const double *window,
                 void *state, float *x, float *y, float *bm_out, float *fix_out, double *p_out,
                 void *stream);
/* The same computation as a pipeline of kernels cut along the algorithm's data dependences (feed-forward alignment /
 * spectra / detector, one warp per blocking filter, one CTA per canceller): several times faster for large batches, at
 * the price of a caller-provided scratch `workspace` of ds_fdgsc_workspace_bytes(p) bytes (about 3x the input).
 * Same arguments, state blob and results as ds_fdgsc_run.                                                      */
size_t ds_fdgsc_workspace_bytes(const ds_fdgsc_params *p);
int ds_fdgsc_run_ws(const ds_fdgsc_params *p, const double *delay_filter, const double *window, void *state,
                    void *workspace, float *x, float *y, float *bm_out, float *fix_out, double *p_out, void *stream);

/* The DC notch of FDGSC.process on its own (FDGSC.py:211-213, feature.py:37-49), in place on
 * x [S][M][n_samples] with the notch memories of `state`: the reference filters the WHOLE input,
 * also the trailing samples beyond the last full block that ds_fdgsc_run does not consume.       */
int ds_fdgsc_notch_run(const ds_fdgsc_params *p, void *state, float *x, int n_samples, void *stream);

/* replaces TimeAlignment.process / fir_filter (fixedbeamformer.py:13-93): streaming per-channel
 * FIR y[n] = sum_k h[k] x[n-k] with a (filter_len-1)-sample cache.  All float64:
 *   h [C][filter_len]   cache [S][C][filter_len-1] in/out   x, y, scratch [S][C][N]   */
int ds_fir_run(int n_streams, int n_ch, int n_samples, int filter_len, const double *h,
               double *cache, const double *x, double *y, double *scratch, void *stream);

/* ---- PCM ingest / egress (beamformer/utils.py) ------------------------------ */
/* replaces the arithmetic of load_audio (utils.py:182-187): out = float32(pcm) / 32767.0f       */
int ds_pcm16_to_float_run(size_t n, const void *pcm_int16, float *out, void *stream);
/* replaces the arithmetic of save_audio (utils.py:190-196): pcm = int16(audio * 32767)          */
int ds_float_to_pcm16_run(size_t n, const float *in, void *pcm_int16, void *stream);
/* the same for a float64 array (what Transform.istft returns): product in double   */
int ds_double_to_pcm16_run(size_t n, const double *in, void *pcm_int16, void *stream);

/* ---- SRP-PHAT (doa/srp.py) ------------------------------------------------ */
/* PHAT normalisation + transpose for the contraction (srp.py:49-50):
 *   X [T][M][K] c64 (one stream of ds_stft_run)  ->  Yhat [K][T][M] c64 = X / (|X| + 1e-6) (phat=1) or X */
int ds_phat_run(int n_frames, int n_mics, int n_bins, int phat, const void *X, void *Yhat, void *stream);
/* replaces the direction x frame loops of srp.compute_angle_spectrum (srp.py:45-51):
 *   P[d, t] = sum_k | sum_m conj(a[d,k,m]) Yhat[k,t,m] |,  a = exp(-j 2 pi f_k tau[d,m]),  f_k = k fs / n_fft
 *   tau [D][M] float32 (MicArray.compute_tau per direction), P [D][T] float32
 *   use_tensor_cores = 1: tcgen05 (tf32) path, n_mics in {4, 8, 16}; 0: CUDA-core path, n_mics <= 16
 *   workspace: ds_srp_workspace_bytes() bytes, 128-byte aligned (tensor path: the spectrum re-tiled and
 *   rounded to tf32 once per call; may be NULL for the CUDA-core path)                              */
size_t ds_srp_workspace_bytes(int n_frames, int n_mics, int n_bins, int use_tensor_cores);
int ds_srp_run(int n_dirs, int n_frames, int n_mics, int n_bins, double fs, int n_fft, const float *tau,
               const void *Yhat, void *workspace, float *P, int use_tensor_cores, void *stream);

/* ---- data-driven steering, mask-based MVDR and GEV weights ---- */
/* replaces the covariance accumulations of example/mvdr.ipynb cell 6 (mask-based MVDR) and the
 * frame-range averages of cell 2:
 *   Phi_xx[s,k] += scale * sum_{t in [t0,t1)} p[s,t,k]       * y y^H
 *   Phi_vv[s,k] += scale * sum_{t in [t0,t1)} (1 - p[s,t,k]) * y y^H        y = X[s,t,:,k]
 *   X [S][T][M][K] c64 or c128   p [S][T][K] float64 or NULL (= 1)   Phi_* [S][K][M][M] c128
 *   Phi_vv may be NULL (it must be when p is NULL).  The caller zero-fills before the first call.   */
int ds_masked_cov_run(int n_streams, int n_frames, int n_mics, int n_bins, int t0, int t1, const void *X,
                      int x_is_c128, const double *p, double scale, void *Phi_xx, void *Phi_vv, void *stream);
/* replaces steering() (beamformer/beamformer.py:10-31): principal eigenvector of each Hermitian
 * matrix (lower triangle read, like numpy.linalg.eigh), unit norm, phase referenced to sensor 0.
 *   XXs [n][M][M] c128 -> out [n][M] c128,  1 <= M <= 8                                            */
int ds_steering_run(long long n, int n_mics, const void *XXs, void *out, void *stream);
/* replaces get_gev_vector() (beamformer.py:77-97): last generalised eigenvector of (target, noise)
 * per matrix pair, normalised to w^H noise w = 1 like scipy.linalg.eigh.  LAPACK leaves the phase of
 * that vector to its tridiagonal solver (component 0 of the standard-form vector real, either sign);
 * here that component is real and NON-NEGATIVE, so results equal the reference's up to a sign per
 * matrix -- which phase_correction() removes for every bin but the first.  A noise matrix that is not
 * positive definite gives the reference's fallback ones * M / trace(noise).
 *   target, noise [n][M][M] c128 -> out [n][M] c128                                               */
int ds_gev_run(long long n, int n_mics, const void *target, const void *noise, void *out, void *stream);
/* replaces phase_correction() (beamformer.py:64-74), in place: W [S][K][M] c128                    */
int ds_phase_correction_run(int n_streams, int n_bins, int n_mics, void *W, void *stream);
/* replaces blind_analytic_normalization() (beamformer.py:34-61):
 *   out = vector * |sqrt(v^H N N v)| / (|v^H N v| + eps);  vector [n][M], noise [n][M][M], out [n][M] c128 */
int ds_ban_run(long long n, int n_mics, const void *vector, const void *noise, double eps, void *out, void *stream);

/* replaces compute_mvdr_weight(steer, np.linalg.inv(Phi_vv)) of mvdr.ipynb cell 6 (beamformer.py:133-155
 * fed with a fresh inverse): w = R^-1 a / (a^H R^-1 a) by Cholesky and two triangular solves.
 *   steer [n][M], Rvv [n][M][M] (Hermitian, lower triangle read) -> w_out [n][M], all c128; NaN where Rvv
 *   is not positive definite                                                                         */
int ds_mvdr_from_cov_run(long long n, int n_mics, const void *steer, const void *Rvv, void *w_out, void *stream);
/* einsum('inj,ij->in', D, w.conj()) with one weight set per stream (mvdr.ipynb cells 6, 8):
 *   X [S][T][M][K] c64 or c128   W [S][K][M] c128   Y [S][T][K] c128                                 */
int ds_apply_stream_weights_run(int n_streams, int n_frames, int n_mics, int n_bins, const void *X, int x_is_c128,
                                const void *W, void *Y, void *stream);

/* ---- Idoa: spatial speech-presence probability (doa/idoa.py) ---- */
/* replaces the RTF recursion of Idoa.estimate (idoa.py:119-127), shared by every direction:
 *   Y_smooth <- (1-alpha) Y_smooth + alpha |X_0|^2 ; Y_xcorr <- (1-alpha) Y_xcorr + alpha X_i conj(X_0) ; B = Y_xcorr / Y_smooth
 *   X [S][T][M][K] c64 or c128 -> B [S][T][2(M-1)+1][K] float64 (Re, Im per channel pair, then ||B||)
 *   state: ds_idoa_rtf_state_bytes() bytes, zero-filled = reset                                          */
size_t ds_idoa_rtf_state_bytes(int n_streams, int n_mics, int n_bins);
int ds_idoa_rtf_run(int n_streams, int n_frames, int n_mics, int n_bins, double alpha, const void *X, int x_is_c128,
                    void *state, double *B, void *stream);
/* replaces the per-direction part of Idoa.estimate (idoa.py:129-165) and, with Yout, the gain of Idoa.process
 * (:199): similarity Delta to the free-field RTF Psi, its H0 / Hd statistics, beta_n from mean(mu_Delta[72:128]),
 * presence probability p.  Directions are independent: theta[n_slots] lists the direction index held by each
 * state slot (any subset of the grid).  only_theta >= 0 reproduces estimate(X, theta=int): every other direction
 * sees Delta = 0.
 *   Psi [n_theta][M-1][K] c128   B from ds_idoa_rtf_run   state [S][n_slots][4][K] float64 (zero = reset)
 *   p_out [S][T][n_slots][K] float64 or NULL
 *   Yout [S][T][K] c128 or NULL: max(mean(p[64:128]), 0.01) * X[:, :, 0, :]; needs n_slots == 1 and X
 *   128 <= K <= 288 (the reference hard-codes bins 64..127 / 72..127)                                     */
size_t ds_idoa_spp_state_bytes(int n_streams, int n_slots, int n_bins);
int ds_idoa_spp_run(int n_streams, int n_frames, int n_mics, int n_bins, int n_slots, const int *theta, int only_theta,
                    const void *Psi, const double *B, void *state, double *p_out, const void *X, int x_is_c128,
                    void *Yout, void *stream);

/* ---- config-4 chain: STFT -> McSppBase -> MVDR -> OMLSA -> ISTFT ------- */
typedef struct ds_chain_params {
  ds_mcspp_params est; /* n_frames is derived: n_samples / hop                    */
  int32_t hop;
  int32_t n_samples; /* multiple of hop                                           */
  int32_t fft_fp64;
  int32_t apply_gain;
  double scale; /* hop / sum(window^2)                                            */
} ds_chain_params;

size_t ds_chain_state_bytes(const ds_chain_params *p);
size_t ds_chain_workspace_bytes(const ds_chain_params *p);
/* replaces the composition pinned in SURVEY.md 8c (mcsppbase.ipynb cell 3,
 * mvdr.ipynb cell 4): x [S][M][N] float32 -> y [S][N] float32.                   */
int ds_chain_run(const ds_chain_params *p, const double *window, const void *a0,
                 void *state, void *workspace, const float *x, float *y, void *stream);

/* Same chain with int16 PCM on either side (the reference's on-disk format,
 * load_audio / save_audio, beamformer/utils.py:182-196): x is [S][M][N] float32 or
 * int16, y is [S][N] float32 or int16; the scalings are fused into the analysis and
 * synthesis kernels (ds_stft_pcm16_run / ds_istft_pcm16_run).                    */
int ds_chain_run_io(const ds_chain_params *p, const double *window, const void *a0,
                    void *state, void *workspace, const void *x, int x_is_pcm16,
                    void *y, int y_is_pcm16, void *stream);

/* Profiling variant (bench only): same work, but records CUDA events between the
 * three kernels, SYNCHRONISES, and returns their durations in milliseconds in
 * phase_ms_h[3] = {analysis, per-bin estimator/beamformer, synthesis} (host). */
int ds_chain_run_profiled(const ds_chain_params *p, const double *window, const void *a0,
                          void *state, void *workspace, const float *x, float *y, void *stream,
                          float *phase_ms_h);

#ifdef __cplusplus
}
#endif
#endif /* DS_B200_H */
