// common.cuh -- context, error plumbing and shared structs of libavld (host side).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/avld.h"

namespace avld {

void set_error(const char* fmt, ...);

// kernel families, for the launch counter and the optional per-stage CUDA-event timing
enum Stage {
  ST_PREP = 0, ST_STFT_MEL, ST_LOGMEL_POST, ST_CONV_DIRECT, ST_CONV_GEMM, ST_DENSE_GEMM, ST_RADII, ST_DECIDE,
  ST_CENTROID, ST_SELECT, ST_SPLIT, ST_FOLD, ST_MAP, ST_COUNT
};

#define AVLD_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::avld::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AVLD_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define AVLD_CHECK(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      ::avld::set_error(__VA_ARGS__);    \
      return (code);                     \
    }                                    \
  } while (0)

#define AVLD_TRY(expr)            \
  do {                            \
    int _r = (expr);              \
    if (_r != AVLD_OK) return _r; \
  } while (0)

// mel filterbank tap of one FFT bin: feeds filter `first` with w0 and `first + 1` with w1
struct __align__(16) MelTap {
  int32_t first;
  float w0, w1;
  int32_t pad;
};

struct PairNode {
  int32_t a, b;  // indices into the value array (leaves first, then internal nodes by height)
};

// encoder layer on the device
struct LayerDev {
  int kind;  // 0 conv (tcgen05, one box per tap), 1 linear (tcgen05), 2 conv direct (CUDA cores, tiny C_in),
             // 3 conv (tcgen05, halo reuse: convh.cu)
  int c_in, c_out, ksize, stride, pad, relu, pool;
  int in_h, in_w, out_h, out_w;  // out = after pooling
  int bn, swz, cblk, cblocks;    // tcgen05 tiling
  int tw, th, tiles_w, tiles_h;
  float* w_f32 = nullptr;  // direct conv weights [c_out][k][k][c_in]
  float* bias = nullptr;   // [c_out]
  __nv_bfloat16* w_hi = nullptr;
  __nv_bfloat16* w_lo = nullptr;  // [c_out][K]
  int64_t K = 0;
  CUtensorMap tm_w_hi, tm_w_lo;
};

}  // namespace avld

struct avld_ctx {
  int device = 0;
  avld_params p{};
  int sm_count = 0;
  int smem_optin = 0;

  // derived feature geometry
  int L = 0, F = 0, R = 0, hpb = 0, kblocks = 0;
  int bin_lo = 0, nbins_pad = 0, ncols = 0, n_tiles_n = 0;
  int T = 0, M = 0;            // target_frames, n_mels
  int crop_start = 0, pad_left = 0, frames_copy = 0;
  bool features_ok = true;     // false: chunk too short / long for the feature kernels (RMS normalisation still works)

  // numpy pairwise-sum plan
  int n_leaves = 0, n_nodes = 0, n_levels = 0;
  int32_t* d_leaf_off = nullptr;
  int32_t* d_leaf_len = nullptr;
  avld::PairNode* d_nodes = nullptr;
  int32_t* d_level_start = nullptr;  // [n_levels + 1]

  avld::MelTap* d_taps = nullptr;  // [nbins_pad]
  __half* d_Bhi = nullptr;         // windowed DFT matrix * 2^dft_scale_log2, [ncols][n_fft], fp16 hi part
  __half* d_Blo = nullptr;         //                                                        fp16 lo part
  int dft_scale_log2 = 10;
  CUtensorMap tm_B_hi, tm_B_lo;

  // folded STFT (default): the Hann window is symmetric (w[k] = w[N-k], w[0] = 0), so
  //   Re X[b] =  sum_{k=1..N/2} (x[k] + x[N-k]) w[k] cos(2 pi k b / N)      (k = N/2: x[N/2] alone)
  //   Im X[b] = -sum_{k=1..N/2-1} (x[k] - x[N-k]) w[k] sin(2 pi k b / N)
  // i.e. two GEMMs with K = N/2 instead of one with K = N: half the tensor-core work.  The folded frames
  // E | O are materialised per pass by fold_kernel as fp16 hi/lo rows [frame][N/2 + N/2].
  int dft_fold = 1;
  int n_tiles2 = 0, last_tile_bins = 0;   // 256-bin N tiles; the last one may hold only 128
  int fold_bk = 64;                       // K elements per pipeline stage of the folded GEMM (64: 2 x 96 KB stages, 32: 4 x 48 KB)
  float4* d_chunk_par = nullptr;   // [max_batch] (scale, pow2, scaled, -) written by prep_kernel, read by fold_kernel
  const float* cur_x = nullptr;    // operand source of the current pass (set by launch_prep, read by launch_fold)
  const int16_t* cur_x16 = nullptr;
  int cur_quantize = 0;
  const uint16_t* cur_q16 = nullptr;   // normalised PCM_16 samples (+32768) of the current pass, or NULL (launch_prep)
  __half* d_A2hi = nullptr;        // [max_batch * F + 128][n_fft] folded frames, E in columns [0, N/2), O in [N/2, N)
  __half* d_A2lo = nullptr;
  __half* d_B2hi = nullptr;        // [n_tiles2 * 512][N/2]: per tile 256 cos rows then 256 (-sin) rows
  __half* d_B2lo = nullptr;
  CUtensorMap tm_A2_hi, tm_A2_lo, tm_B2_hi, tm_B2_lo;
  CUtensorMap tm_B2h_hi, tm_B2h_lo;   // 128-row boxes of B2 for the CTA-pair kernel
  CUtensorMap tm_A2pf_hi, tm_A2pf_lo; // un-swizzled 256-tap x 128-row boxes of A2, used only for L2 prefetch (4x fewer TMA rows)
  int dft_pair = 1;                   // folded GEMM on CTA pairs (cta_group::2); AVLD_DFT_MODE=fold1 selects the 1-CTA kernel

  // twice-folded STFT ("fold2", default when n_fft % 512 == 0): a second time-reversal fold splits the bins by parity,
  //   even b:  Re X = sum_{k<N/4} (E[k] + E[N/2-k]) cos(2 pi k b / N) + E[N/4] cos(pi b / 2),   -Im X = sum (O[k] - O[N/2-k]) sin(.)
  //   odd  b:  Re X = sum_{k<N/4} (E[k] - E[N/2-k]) cos(.),   -Im X = sum (O[k] + O[N/2-k]) sin(.) + O[N/4] sin(pi b / 2)
  // (E/O = first fold of the WINDOWED frame), i.e. K = N/4 per cos / sin GEMM: a quarter of the plain DFT GEMM's tensor
  // work.  fold2.cu writes the four folded sequences per frame (same bytes as one fold) and the two edge terms; dftf3.cu
  // runs 160-bin work items per class and adds the edge term in its epilogue; the two classes accumulate mel power into
  // separate planes (summed by logmel_post_kernel) so that the float sums stay order independent.
  // f2_levels == 3 (default) folds the even bins once more (b = 0 / 2 mod 4: K = N/8 per cos / sin GEMM; the odd bins'
  // symmetry is spent), see fold2.cu::fold3_kernel; AVLD_DFT_MODE=fold2 keeps the two-level form.
  int dft_fold2 = 0;
  int f2_levels = 3;
  int f2_items = 0, f2_classes = 0;
  struct F2Item { int a_col0, kbp, cls, edge_im; };   // A column of the cos part, 64-tap K blocks per part, bin class
  F2Item f2_item[8] = {};                              // (= mel plane and edge component), edge term goes to Im?
  __half* d_B3hi = nullptr;        // [f2_items * 2 * 160][N/4]: per item 160 cos rows then 160 sin rows (no window)
  __half* d_B3lo = nullptr;
  CUtensorMap tm_B3_hi, tm_B3_lo;  // 64-tap x 80-row boxes (one CTA's half of an item)
  avld::MelTap* d_taps3 = nullptr; // [f2_items * 160], .pad = bits of the edge coefficient
  float4* d_edge = nullptr;        // [max_batch * F + 256] per frame: the self-paired tap of each bin class
  uint16_t* d_q16 = nullptr;       // [max_batch * L + 64] normalised, PCM_16-rounded samples biased by 32768 (three-level fold, quantize passes)
  float* d_win = nullptr;          // [N/2 + 1] periodic Hann
  bool planes_dirty = false;       // a GEMM pass accumulated into the planes and logmel_post has not consumed them yet
  long long melpow_plane = 0;      // elements per mel-power plane (fold2: one plane per bin class)

  // per-pass scratch (max_batch chunks)
  int max_batch = 0;
  __half* d_Ahi = nullptr;         // padded, pow2-scaled audio rows [max_batch * R + 128][hop]
  __half* d_Alo = nullptr;
  CUtensorMap tm_A_hi, tm_A_lo;
  float* d_inv2 = nullptr;         // [max_batch] 2^(-2 s_c)
  float* d_melpow = nullptr;       // [max_batch * R][n_mels]
  float* d_feat = nullptr;         // [max_batch][T][M]
  float* d_mu = nullptr;           // [max_batch][D]
  float* d_radii = nullptr;        // [max_batch][K<=64]
  uint8_t* d_ok = nullptr;
  float* d_rms = nullptr;

  // encoder
  std::vector<avld::LayerDev> layers;
  int latent_dim = 0;
  size_t act_elems = 0;            // per-chunk max activation elements
  __nv_bfloat16* d_act_hi[2] = {nullptr, nullptr};
  __nv_bfloat16* d_act_lo[2] = {nullptr, nullptr};
  std::vector<CUtensorMap> tm_act_hi, tm_act_lo;  // per layer input maps

  // host end-to-end path
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  float* d_xbuf[2] = {nullptr, nullptr};
  float* d_cent = nullptr;
  double* d_thr = nullptr;
  int32_t* d_prio = nullptr;
  int32_t* d_pred = nullptr;
  float* d_best = nullptr;
  void* h_stage = nullptr;         // pinned staging for results (a D2H into pageable memory would block the host
  size_t h_stage_bytes = 0;        // thread until the slab's kernels finish and serialise copy against compute)

  // order-statistics scratch
  unsigned int* d_hist = nullptr;
  size_t hist_bytes = 0;

  // launch accounting (always on) and per-stage event timing (avld_profile_enable)
  uint64_t launches[avld::ST_COUNT] = {};
  bool profiling = false;
  struct TimedLaunch { int stage; cudaEvent_t start, stop; };
  std::vector<TimedLaunch> timed;
  std::vector<cudaEvent_t> event_pool;
};

namespace avld {

// tensor-map helpers (driver entry point fetched at runtime; libcuda is not linked)
int encode_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, uint64_t dim0, uint64_t dim1,
                   uint64_t stride1_bytes, uint32_t box0, uint32_t box1, uint32_t swizzle_bytes);
int encode_tmap_4d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, const uint64_t dims[4],
                   const uint64_t strides_bytes[3], const uint32_t box[4], uint32_t swizzle_bytes);

// host math
int64_t pairwise_plan(int64_t n, std::vector<int64_t>& off, std::vector<int64_t>& len, std::vector<PairNode>& nodes,
                      std::vector<int32_t>& level_start);
int mel_taps_host(const avld_params& p, std::vector<int32_t>& first, std::vector<float>& w0, std::vector<float>& w1,
                  int* bin_lo, int* bin_hi);

// RAII bracket around one kernel launch: counts it and, when profiling, records CUDA events on the
// launching stream (so the measured duration is that kernel's, inside the real step).
struct LaunchScope {
  avld_ctx* c;
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  LaunchScope(avld_ctx* ctx, int stage, cudaStream_t stream);
  ~LaunchScope();
};

// stage launchers (each in its own translation unit)
int launch_prep(avld_ctx* c, const float* x, const int16_t* x16, float* y_out, bool write_operand, bool normalize, uint8_t* ok, float* rms,
                int n, float target_rms, float rms_min, float eps, int quantize, cudaStream_t st);
int launch_stft_mel(avld_ctx* c, int n, cudaStream_t st);
int launch_fold(avld_ctx* c, int n, cudaStream_t st);
int launch_fold2(avld_ctx* c, int n, cudaStream_t st);
int launch_stft_mel_fold2(avld_ctx* c, int n, cudaStream_t st);
bool dftf4_supported(const avld_ctx* c);                               // dftf4.cu: experimental, AVLD_DFT_DUAL=1
int launch_stft_mel_fold2_dual(avld_ctx* c, int n, cudaStream_t st);
bool dftg_supported(const avld_ctx* c);
int launch_stft_mel_gen(avld_ctx* c, const float* x, const int16_t* x16, int n, cudaStream_t st);
int launch_stft_mel_pair(avld_ctx* c, int n, cudaStream_t st);
int launch_logmel_post(avld_ctx* c, float* feat, int n, cudaStream_t st);
int launch_encoder(avld_ctx* c, const float* feat, float* mu, int n, cudaStream_t st);
int convh_encode_input_map(CUtensorMap* out, const void* base, int n, int H, int W, int C, int cblk);
bool convh_supported(int c_in, int c_out, int ksize, int w);
int launch_convh(const LayerDev& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int n, __nv_bfloat16* out_hi,
                 __nv_bfloat16* out_lo, int sm_count, cudaStream_t st);
int launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t st);
int launch_split_f16(const float* src, __half* hi, __half* lo, size_t n, cudaStream_t st);

}  // namespace avld
