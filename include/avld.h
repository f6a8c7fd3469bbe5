/* avld.h -- C ABI of the B200-native encode+detect hot path ("amphibian VAE latent detector").
 *
 * The reference (vpobleteacustica/amphibian-vae-latent-detector) is pure Python and has no FFI;
 * its de-facto interface for this path is a set of Python functions.  Each
 * entry point below names the reference function(s) it replaces (paths relative to
 * latent_space_exploration/ in the reference).  A Python maintainer binds these with ctypes --
 * see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every function returns AVLD_OK (0) or a negative error code; nothing throws across the ABI;
 *     avld_last_error() returns a thread-local message for the last failure.
 *   - pointers documented "dev" are caller-owned CUDA device memory, valid on ctx's device, never
 *     freed or retained past the call (exception: none -- encoder weights are copied at load).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous unless
 *     the function is documented as synchronous.
 *   - one ctx per (device, host thread); a ctx is not safe for concurrent calls.
 *   - per-chunk failures never fail the call: they are reported in `ok`/`pred`, mirroring the
 *     reference's count-and-continue loops (08_fit_radial_detector.py:489-506,
 *     10_benchmark_folder_detection.py:397-418).
 *   - there is no CPU fallback: without a CUDA device avld_ctx_create fails with AVLD_ERR_CUDA.
 */
#ifndef AVLD_H
#define AVLD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVLD_ABI_VERSION 1

enum {
  AVLD_OK = 0,
  AVLD_ERR_INVALID = -1,      /* bad argument                                                     */
  AVLD_ERR_CUDA = -2,         /* CUDA runtime/driver error (message has the CUDA error string)     */
  AVLD_ERR_UNSUPPORTED = -3,  /* valid request the kernels do not implement (e.g. hop % 64 != 0)   */
  AVLD_ERR_STATE = -4,        /* call order (e.g. forward before encoder_load)                     */
  AVLD_ERR_NOMEM = -5
};

typedef struct avld_ctx avld_ctx;

/* Feature parameters = the CLI defaults shared by 07/08/09/10 (07_encode_wav_to_latent.py:424-432,
 * 08_fit_radial_detector.py:348-354) plus librosa.power_to_db's amin/top_db
 * (map_detector_core.py:229). chunk_len = int(sr * duration) (map_detector_core.py:213). */
typedef struct avld_params {
  int32_t sr;             /* 48000 */
  int32_t chunk_len;      /* samples per chunk, e.g. 144000 (3 s) or 240000 (5 s) */
  int32_t n_fft;          /* 2048 */
  int32_t hop;            /* 384 */
  int32_t n_mels;         /* 64 */
  float fmin;             /* 150 */
  float fmax;             /* 15000 */
  int32_t target_frames;  /* 192 */
  float amin;             /* 1e-10 */
  float top_db;           /* 80 */
  int32_t max_batch;      /* chunks per internal pass; device scratch is sized from this */
} avld_params;

/* One encoder layer, BatchNorm already folded (host pointers; copied during avld_encoder_load).
 * kind 0 = Conv2d(+ReLU)(+MaxPool2d(2)) on NHWC activations, weight [c_out][k][k][c_in];
 * kind 1 = Linear(+ReLU), weight [c_out][c_in] with c_in in NHWC-flatten order. */
typedef struct avld_layer {
  int32_t kind;
  int32_t c_in, c_out;
  int32_t ksize, stride, pad;
  int32_t relu, pool;
  int32_t in_h, in_w;
  const float* weight;
  const float* bias;
} avld_layer;

/* One operation of an encoder PROGRAM: a dataflow graph over numbered tensors (tensor 0 = one feature segment,
 * [seg_frames, n_mels], single channel; images are NHWC on the device).  What encoder.py::export_program emits for an
 * nn.Module built from Conv2d / BatchNorm2d / ReLU / Max- / AvgPool2d / AdaptiveAvgPool2d(1) / residual add / Flatten /
 * Linear; weights are host pointers, copied (and padded to the kernels' channel granularity) at load. */
enum {
  AVLD_OP_CONV = 0,    /* out = pool(act(conv(in0) + bias)); weight [c_out][k][k][c_in]; stride 1 or 2; pool 0 none, 1 max 2x2,
                          2 average 2x2 (after the activation) */
  AVLD_OP_LINEAR = 1,  /* out = act(in0 . W^T + bias); weight [c_out][c_in], c_in in NHWC-flatten order of in0 */
  AVLD_OP_ADD = 2,     /* out = act(in0 + in1): the residual connection */
  AVLD_OP_AFFINE = 3,  /* out = act(in0 * weight[c] + bias[c]): a BatchNorm that cannot be folded, a lone ReLU */
  AVLD_OP_POOL = 4     /* pool 1 max / 2 average over ksize x ksize windows with `stride`, no padding; ksize 0 = global average */
};
typedef struct avld_op {
  int32_t kind;
  int32_t in0, in1, out;       /* tensor ids; in1 = the second operand of AVLD_OP_ADD, or for AVLD_OP_CONV an optional residual of the output's
                                * shape added before the activation (no pooling then), else -1; every id is written exactly once */
  int32_t c_in, c_out;
  int32_t ksize, stride, pad;
  int32_t relu;                /* act = ReLU */
  int32_t pool;
  int32_t in_h, in_w;          /* of in0 (checked against the producer) */
  const float* weight;
  const float* bias;
} avld_op;

/* A requested order statistic: the `rank`-th smallest (0-based) of radii[:, species] over the rows
 * whose label == species (side 0, "in class") or label != species and label >= 0 (side 1). */
typedef struct avld_rank_query {
  int32_t species;
  int32_t side;
  int64_t rank;
} avld_rank_query;

int avld_abi_version(void);
const char* avld_last_error(void);

/* ---- context ---------------------------------------------------------------------------------- */
int avld_ctx_create(int device, const avld_params* params, avld_ctx** out);
void avld_ctx_destroy(avld_ctx* ctx);
/* derived sizes: frames per chunk F = 1 + chunk_len / hop, latent dim D (0 before encoder_load) */
int avld_ctx_info(const avld_ctx* ctx, int32_t* n_frames, int32_t* latent_dim, int32_t* sm_count);

/* Constants and scalar arithmetic of rms_normalize (00_normalize_dataset_rms.py:29-38) for this context.
 * scalar_semantics 0 (default): numpy >= 2 -- `rms + eps`, `target_rms / (...)` and the `rms < rms_min` gate are float32
 * operations (a Python float next to a float32 scalar is "weak").  1: numpy 1.x value-based casting, which the reference's
 * pinned numpy==1.26.4 executes -- the three are float64 operations and the scale is rounded to float32 once (the two modes
 * differ by one ulp of the scale in about a third of all chunks).  The fused host calls (avld_encode_detect_host*) take
 * their constants from here; the per-call float parameters of the other entry points must equal (float) of these values
 * when scalar_semantics is 1.  Defaults: 0, 0.05, 1e-4, 1e-8. */
int avld_ctx_set_normalization(avld_ctx* ctx, int scalar_semantics, double target_rms, double rms_min, double eps);

/* How the STFT is evaluated on this context (accounting for bench.py's roofline line): `mode` receives a static string
 * ("fold3": the three-times folded GEMM of dftf3.cu, the only form the library carries), algorithmic = the flops of the plain windowed DFT GEMM
 * restricted to the bins with mel weight (2 * F * n_fft * 2 * bins), issued = the tensor-core
 * flops the selected kernel actually issues per chunk (all split-precision passes, padded tiles included). */
int avld_ctx_dft_info(const avld_ctx* ctx, const char** mode, double* algorithmic_flops_per_chunk,
                      double* issued_flops_per_chunk);

/* ---- accounting -------------------------------------------------------------------------------
 * Every kernel launch is counted per kernel family ("stage"); with profiling enabled each launch is
 * additionally bracketed by CUDA events on its own stream.  avld_profile_collect synchronises on the
 * recorded events and returns, per stage, the summed device time (ms), the number of timed launches
 * and the total launch count; arrays of avld_stage_count() entries, any may be NULL. */
int avld_profile_enable(avld_ctx* ctx, int on);
int avld_profile_collect(avld_ctx* ctx, double* ms, int64_t* timed_launches, uint64_t* launches, int reset);
int avld_stage_count(void);
const char* avld_stage_name(int stage);

/* ---- R1/R2: rms_normalize (00_normalize_dataset_rms.py:29-38), batched -------------------------
 * x, y: dev float32 [n, chunk_len]; ok: dev uint8 [n] (1 = scaled, 0 = silence gate: copied
 * unchanged); rms: dev float32 [n] or NULL.  Bit-exact with numpy-2 float32 semantics (pairwise
 * sum tree, non-fused ops).  quantize_pcm16 != 0 additionally applies the sf.write -> librosa.load
 * PCM_16 round trip of process_folder (00:55-57): y = rint(y * 32767) / 32768. */
int avld_rms_normalize(avld_ctx* ctx, const float* x, float* y, uint8_t* ok, float* rms, int64_t n,
                       float target_rms, float rms_min, float eps, int quantize_pcm16, void* stream);

/* ---- M1: the resampling of `librosa.load(path, sr=sr)` (map_detector_core.py:210, 00_normalize_dataset_rms.py:51) for a
 * file whose rate differs from sr: librosa 0.9.2 resample(res_type="kaiser_best") = resampy's Kaiser-windowed sinc
 * interpolation, restated tap for tap (float64 weights, float32 running sum).  x: dev float32 [n_in] mono samples at
 * sr_in -> y: dev float32 [n_out], n_out = avld_resample_len(n_in, sr_in, sr_out) = ceil(n_in * sr_out / sr_in).
 * Synchronous (returns after the kernel). */
int64_t avld_resample_len(int64_t n_in, int32_t sr_in, int32_t sr_out);
int avld_resample(avld_ctx* ctx, const float* x, int64_t n_in, int32_t sr_in, int32_t sr_out, float* y, int64_t n_out,
                  void* stream);

/* ---- M2-M5 + E0: wav_to_mel after the load (map_detector_core.py:219-237) and the transpose of
 * map_detector_core.py:267-268.  y: dev float32 [n, chunk_len] -> feat: dev float32
 * [n, target_frames, n_mels] (the encoder's [B,1,T,M] input). */
int avld_logmel(avld_ctx* ctx, const float* y, float* feat, int64_t n, void* stream);

/* fused R1(+R2) + features: x -> feat; ok/rms as in avld_rms_normalize (either may be NULL). */
int avld_normalize_logmel(avld_ctx* ctx, const float* x, float* feat, uint8_t* ok, float* rms,
                          int64_t n, float target_rms, float rms_min, float eps, int quantize_pcm16,
                          void* stream);

/* ---- L1/E1/E2: encoder (map_detector_core.py:150-179, :270-300) --------------------------------
 * encoder_load is synchronous (copies and pre-splits the weights, builds TMA descriptors). */
int avld_encoder_load(avld_ctx* ctx, const avld_layer* layers, int32_t n_layers);
/* General form.  out_tensor = id of the latent; out_is_map != 0: the latent is a feature map [C, H, W], flattened in NCHW
 * order as the reference does for rank > 2 outputs (map_detector_core.py:294-295).  The feature image [target_frames,
 * n_mels] is cut into n_seg segments of seg_frames frames (seg_frames * n_seg == target_frames), each one runs through
 * the program, and the chunk's latent is the mean over its segments -- the reference's `t.mean(dim=1)` for encoders that
 * return [B, n_seg, C] (map_detector_core.py:292-293, 07_encode_wav_to_latent.py:287-291). */
int avld_encoder_load_program(avld_ctx* ctx, const avld_op* ops, int32_t n_ops, int32_t out_tensor, int32_t out_is_map,
                              int32_t seg_frames, int32_t n_seg);
/* feat: dev float32 [n, target_frames, n_mels] -> mu: dev float32 [n, D] (the latent mean). */
int avld_encoder_forward(avld_ctx* ctx, const float* feat, float* mu, int64_t n, void* stream);

/* whole encode half in one call: x dev [n, chunk_len] -> mu dev [n, D]
 * (= rms_normalize + sf.write/load + encode_wav_to_latent, map_detector_core.py:240-300). */
int avld_encode(avld_ctx* ctx, const float* x, float* mu, uint8_t* ok, int64_t n, float target_rms,
                float rms_min, float eps, int quantize_pcm16, void* stream);

/* same, with the chunks as device-resident PCM_16 samples (what the reference's WAV datasets hold; decoded as s / 32768
 * like librosa.load, 00:51 / core:210): half the HBM bytes of the float32 form, bit-identical results. */
int avld_encode_pcm16(avld_ctx* ctx, const int16_t* pcm, float* mu, uint8_t* ok, int64_t n, float target_rms,
                      float rms_min, float eps, int quantize_pcm16, void* stream);

/* ---- F1-F3: radial fit pieces (08_fit_radial_detector.py:105-106, :310-333, :530-558) ----------
 * per-species latent sums for the centroid (np.mean(Z_in, axis=0), 08:316): ACCUMULATES into
 * sum (dev float64 [K, D]) and cnt (dev int64 [K]); rows with label < 0 or >= K are skipped. */
int avld_centroid_accumulate(avld_ctx* ctx, const float* Z, const int32_t* label, double* sum,
                             int64_t* cnt, int64_t n, int32_t K, int32_t D, void* stream);
/* radii[i, k] = || Z[i] - centroid[k] ||_2 (08:105-106, :318, :325; 09:354-355). */
int avld_radii(avld_ctx* ctx, const float* Z, const float* centroid, float* radii, int64_t n,
               int32_t K, int32_t D, void* stream);
/* exact order statistics of radii columns (np.quantile's partition step, 08:109-112); the caller
 * interpolates.  radii dev [n, K], label dev [n], queries HOST [n_q], out HOST float32 [n_q].
 * Synchronous on `stream`. */
int avld_order_stats(avld_ctx* ctx, const float* radii, const int32_t* label, int64_t n, int32_t K,
                     const avld_rank_query* queries, int32_t n_q, float* out, void* stream);

/* ---- multi-GPU fit without torch.distributed (the two exchanges of 08:530-558 when the rows are sharded over
 * ranks, one process and one context per GPU).  NCCL is loaded at run time (dlopen of libnccl.so.2); the host ships the
 * 128-byte id from rank 0 to the other ranks by its own means (file, socket, MPI ...).
 *   avld_comm_unique_id       rank 0: a fresh id
 *   avld_comm_init            every rank, collectively: the context's communicator
 *   avld_allreduce_centroids  sum dev float64 [K, D] and cnt dev int64 [K] (as filled by avld_centroid_accumulate) are
 *                             summed over the ranks IN PLACE: every rank then forms bit-identical centroids (08:316)
 *   avld_allgather_radii      radii dev [n_local, K] + label dev [n_local] -> radii_all dev [world * shard_rows, K],
 *                             label_all dev [world * shard_rows]; n_local <= shard_rows, padding rows get label -1 (skipped
 *                             by avld_order_stats): every rank then selects the same exact order statistics (08:319, :326)
 * All stream-ordered and asynchronous. */
#define AVLD_COMM_ID_BYTES 128
int avld_comm_unique_id(void* id_out);
int avld_comm_init(avld_ctx* ctx, const void* nccl_unique_id, int32_t rank, int32_t world);
int avld_comm_destroy(avld_ctx* ctx);
int avld_allreduce_centroids(avld_ctx* ctx, double* sum, int64_t* cnt, int32_t K, int32_t D, void* stream);
int avld_allgather_radii(avld_ctx* ctx, const float* radii, const int32_t* label, int64_t n_local, int64_t shard_rows,
                         int32_t K, float* radii_all, int32_t* label_all, void* stream);

/* ---- D2: decision (09_evaluate_wav_detection.py:416-436; 10_benchmark_folder_detection.py:175-199)
 * accept k iff (double)radii[i,k] <= thr[k]; pred[i] = accepted k with the smallest
 * priority_rank[k] or -1 (NO_DETECT); best_d[i] = min_k radii[i,k].  thr dev float64 [K] (a NaN
 * threshold never accepts but its distance still counts for best_d, 10:177-187; a species ABSENT from the
 * config's thresholds is left out of the arrays by the caller, 09:418-419), priority_rank dev int32 [K]. */
int avld_decide(avld_ctx* ctx, const float* radii, const double* thr, const int32_t* priority_rank,
                int32_t* pred, float* best_d, int64_t n, int32_t K, void* stream);

/* ---- N1: Gaussian-MAP detector on latents (map_detector_core.py:319-323; 09n_evaluate_wav_detection.py:114-140;
 * 10b_benchmark_folder_detection_map.py:146-169) ------------------------------------------------------------------
 * score[i,k] = -0.5 * (d^T P_k d + a_const[k]) + log_prior[k], d = Z[i] - mean[k], a_const = logdet_cov + D ln(2 pi),
 * log_prior = ln(prior + 1e-12); the quadratic form is accumulated in float32 (as `diff.T @ prec @ diff` is), the rest in
 * float64.  pred[i] = first k (caller passes the species in sorted-name order) with the strictly largest score, or -1
 * when use_tau and best < tau; best[i] = the largest score.  mean dev [K,D] f32, precision dev [K,D,D] f32, a_const /
 * log_prior dev [K] f64, scores dev [n,K] f64 or NULL.  D <= 256. */
int avld_map_score(avld_ctx* ctx, const float* Z, const float* mean, const float* precision, const double* a_const,
                   const double* log_prior, double tau, int use_tau, int32_t* pred, double* best, double* scores,
                   int64_t n, int32_t K, int32_t D, void* stream);
/* second moments for the MAP fit (08b_fit_map_detector.py:60-81, :276-296: np.cov of centred latents, float64):
 * out[D,D] (dev f64, ACCUMULATES) += sum over rows with label == k_sel (k_sel < 0: every labelled row, each centred by
 * its own class mean = the LDA pooling) of (z - mean[label]) (z - mean[label])^T. */
int avld_cov_accumulate(avld_ctx* ctx, const float* Z, const int32_t* label, const float* mean, int32_t k_sel,
                        double* out, int64_t n, int32_t K, int32_t D, void* stream);

/* ---- end to end with HOST buffers (the call the drop-in Python layer makes per batch) -----------
 * x_host float32 [n, chunk_len] (pinned or pageable) -> pred_host int32 [n], best_host float32 [n],
 * mu_host float32 [n, D] (nullable), ok_host uint8 [n] (nullable).  centroid/thr/priority_rank are
 * HOST arrays ([K, D] float32, [K] float64, [K] int32).  Copies are double-buffered against compute.
 * Synchronous. */
int avld_encode_detect_host(avld_ctx* ctx, const float* x_host, int64_t n, int quantize_pcm16,
                            const float* centroid, const double* thr, const int32_t* priority_rank,
                            int32_t K, int32_t* pred_host, float* best_host, float* mu_host,
                            uint8_t* ok_host);

/* Same, with the audio as PCM_16 samples (what the reference's WAV files hold; librosa.load decodes them as
 * s / 32768, 00:51 / core:210): pcm_host int16 [n, chunk_len].  Halves the bytes crossing PCIe; the decode is
 * fused into the normalisation kernel and is exact. */
int avld_encode_detect_host_pcm16(avld_ctx* ctx, const int16_t* pcm_host, int64_t n, int quantize_pcm16,
                                  const float* centroid, const double* thr, const int32_t* priority_rank,
                                  int32_t K, int32_t* pred_host, float* best_host, float* mu_host,
                                  uint8_t* ok_host);

/* ---- host-side helpers (no device work; used by tests and by the Python layer) -----------------*/
/* numpy's float32 pairwise-summation plan for a length-n reduction: writes up to cap leaves
 * (offset, length) in in-order traversal; returns the number of leaves (or a negative error). */
int64_t avld_pairwise_plan(int64_t n, int64_t* leaf_offset, int64_t* leaf_len, int64_t cap);
/* slaney mel filterbank taps: for FFT bin b, weight[0] feeds filter first[b], weight[1] feeds
 * first[b] + 1 (first[b] = -1 when the bin feeds nothing).  Arrays of n_fft/2 + 1 entries. */
int avld_mel_taps(const avld_params* params, int32_t* first, float* w0, float* w1);

/* ---- test / bring-up entry (exercises the tcgen05 split-precision GEMM core on plain matrices) --
 * C[M,N] = A[M,K] * B[N,K]^T, all dev float32 row-major, K % 64 == 0, N % 16 == 0, N <= 256*tiles.
 * mode 0: fp16 hi + fp16 lo operands; mode 1: bf16 hi + bf16 lo. Synchronous. */
int avld_dbg_gemm(avld_ctx* ctx, const float* A, const float* B, float* C, int64_t M, int32_t N,
                  int32_t K, int32_t mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVLD_H */
