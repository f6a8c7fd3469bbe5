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
  ST_CENTROID, ST_SELECT, ST_SPLIT, ST_FOLD, ST_MAP, ST_ELEMENTWISE, ST_COUNT
};   // names: ctx.cu::avld_stage_name

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

// one operation of the encoder program on the device (encoder.cu); tensors are NHWC bf16 hi + lo planes in numbered slots
struct OpDev {
  int kind;   // OP_* below
  int src = -1, src2 = -1, dst = -1;   // tensor ids
  int c_in = 0, c_out = 0;             // padded channel counts (tensor layouts), logical ones only matter at load time
  int ksize = 1, stride = 1, pad = 0, relu = 0;
  int pool = 1;                        // fused 2x2 pooling after the activation (1 = none, 2 = yes), pool_avg: average instead of max
  int pool_avg = 0;
  int in_h = 0, in_w = 0, out_h = 0, out_w = 0;   // out = after pooling
  int bn = 0, swz = 0, cblk = 0, cblocks = 0;     // tcgen05 tiling
  int tw = 0, th = 0, tiles_w = 0, tiles_h = 0;
  int conv_h = 0, conv_w = 0;          // convolution output size before pooling
  float* w_f32 = nullptr;  // first-layer conv weights [c_out][k][k][c_in]; affine: scale [c]
  float* bias = nullptr;   // [c_out]; affine: shift [c]
  __nv_bfloat16* w_hi = nullptr;
  __nv_bfloat16* w_lo = nullptr;  // [c_out][K]
  int64_t K = 0;
  CUtensorMap tm_w_hi, tm_w_lo, tm_in_hi, tm_in_lo;
};
enum {
  OP_CONV_GEMM = 0,    // implicit GEMM on tcgen05, one TMA box per filter tap (any odd k, stride 1 / 2)
  OP_LINEAR = 1,       // plain GEMM on tcgen05
  OP_CONV_FIRST = 2,   // first layer (fp32 single-channel feature image in), CUDA cores
  OP_CONV_HALO = 3,    // 3x3 stride-1 implicit GEMM with halo reuse (convh.cu)
  OP_ADD = 4, OP_AFFINE = 5, OP_POOL = 6, OP_GAP = 7
};

struct TensorDev {
  int c = 0, h = 0, w = 0;   // logical; vectors: h = w = 1
  int c_pad = 0;             // channels of the device layout (multiple of 32 / 64, pad channels hold zeros)
  int slot = -1;             // activation slot (tensor 0 = the fp32 feature image: no slot)
  size_t elems() const { return static_cast<size_t>(h) * w * c_pad; }
};

}  // namespace avld

struct avld_ctx {
  int device = 0;
  avld_params p{};
  int sm_count = 0;
  int smem_optin = 0;

  // R1 constants and scalar arithmetic of rms_normalize (avld_ctx_set_normalization).  scalar_f64 = 0: numpy >= 2 rules
  // (`rms + eps` and `target / (...)` in float32, the gate against float32(rms_min)); 1: numpy 1.x value-based casting, what
  // the reference's pinned numpy==1.26.4 executes (both scalar operations and the gate comparison in float64, the scale
  // rounded to float32 once).  The per-call float parameters of the ABI are used in mode 0, these doubles in mode 1.
  int scalar_f64 = 0;
  double norm_target = 0.05, norm_rms_min = 1e-4, norm_eps = 1e-8;

  // derived feature geometry
  int L = 0, F = 0;            // samples and STFT frames per chunk
  int T = 0, M = 0;            // target_frames, n_mels
  int crop_start = 0, pad_left = 0, frames_copy = 0;
  bool features_ok = true;     // false: chunk too short / long for the feature kernels (RMS normalisation still works)

  // numpy pairwise-sum plan
  int n_leaves = 0, n_nodes = 0, n_levels = 0;
  bool leaves_regular = false;       // every leaf is 8 .. 128 samples, a multiple of 8, at a multiple-of-8 offset (3 s and 5 s chunks)
  int min_leaf_rows = 0;             // shortest leaf / 8
  bool tree_perfect = false;         // 2^h leaves under a tree of height h: siblings are adjacent leaves / nodes at every level
  int32_t* d_leaf_off = nullptr;
  int32_t* d_leaf_len = nullptr;
  avld::PairNode* d_nodes = nullptr;
  int32_t* d_level_start = nullptr;  // [n_levels + 1]

  // Three-times folded STFT.  The Hann window is symmetric and a real DFT has time-reversal symmetry, so
  // with E / O the even / odd fold of the WINDOWED frame the bins split by parity, and the even ones once more:
  //   odd b        Re X = sum_{k<N/4} (E[k] - E[N/2-k]) cos(2 pi k b / N),   -Im X = sum (O[k] + O[N/2-k]) sin(.) + O[N/4] sin(pi b / 2)
  //   b = 0 mod 4  Re X = sum_{k<N/8} (P[k] + P[N/4-k]) cos(.) + P[N/8] cos(pi b / 4),   -Im X = sum (R[k] - R[N/4-k]) sin(.)
  //   b = 2 mod 4  Re X = sum_{k<N/8} (P[k] - P[N/4-k]) cos(.),   -Im X = sum (R[k] + R[N/4-k]) sin(.) + R[N/8] sin(pi b / 4)
  // fold3.cu writes the folded sequences of every frame (tile-major fp16 hi / lo) and the three self-paired edge terms;
  // dftf3.cu runs 160-bin work items per class, adds the edge term in its epilogue and accumulates mel power into one plane
  // per class (summed in a fixed order by logmel_post_kernel) so that the float sums stay order independent.
  static constexpr int kMaxItems = 12;
  int dft_scale_log2 = 10;
  int f2_items = 0, f2_classes = 0;
  struct F2Item { int a_col0, kbp, cls, edge_im; };   // A column of the cos part, 64-tap K blocks per part, bin class
  F2Item f2_item[kMaxItems] = {};                      // (= mel plane and edge component), edge term goes to Im?
  float4* d_chunk_par = nullptr;   // [max_batch] (scale, pow2, scaled, -) written by prep_kernel, read by fold3_kernel
  const float* cur_x = nullptr;    // operand source of the current pass (set by launch_prep, read by launch_fold3)
  const int16_t* cur_x16 = nullptr;
  int cur_quantize = 0;
  const uint16_t* cur_q16 = nullptr;   // normalised PCM_16 samples (+32768) of the current pass, or NULL (launch_prep)
  __half* d_A3 = nullptr;          // folded frames, tile-major: [tile][K block][hi | lo][128][64] (fold3.cu)
  size_t a3_tiles = 0;
  int a3_kblocks = 0;              // n_fft / 64
  CUtensorMap tm_A3;               // 64-tap x 256-row boxes = one (tile, K block): 32 KB of contiguous memory
  __half* d_B3hi = nullptr;        // [f2_items * 2 * 160][N/4]: per item 160 cos rows then 160 sin rows (no window)
  __half* d_B3lo = nullptr;
  CUtensorMap tm_B3_hi, tm_B3_lo;  // 64-tap x 80-row boxes (one CTA's half of an item)
  avld::MelTap* d_taps3 = nullptr; // [f2_items * 160], .pad = bits of the edge coefficient
  float4* d_edge = nullptr;        // [max_batch * F + 256] per frame: the self-paired tap of each bin class
  uint16_t* d_q16 = nullptr;       // [max_batch * L + 64] normalised, PCM_16-rounded samples biased by 32768 (quantize passes)
  float* d_win = nullptr;          // [N/2 + 1] periodic Hann + the per-block table of fold3_kernel
  bool planes_dirty = false;       // a GEMM pass accumulated into the planes and logmel_post has not consumed them yet
  long long melpow_plane = 0;      // elements per mel-power plane (one plane per bin class)

  // per-pass scratch (max_batch chunks)
  int max_batch = 0;
  float* d_inv2 = nullptr;         // [max_batch] 2^(-2 s_c)
  float* d_melpow = nullptr;       // [classes][max_batch * F + 256][n_mels]
  float* d_feat = nullptr;         // [max_batch][T][M]
  float* d_mu = nullptr;           // [max_batch][D]
  float* d_radii = nullptr;        // [max_batch][K<=64]
  uint8_t* d_ok = nullptr;
  float* d_rms = nullptr;

  // encoder program (avld_encoder_load_program)
  std::vector<avld::OpDev> ops;
  std::vector<avld::TensorDev> tensors;
  int enc_out = -1;                // tensor id of the latent
  int enc_out_nchw = 0;            // the latent is a feature map: flattened in NCHW order
  int n_seg = 1, seg_frames = 0;   // segments per chunk (the latent is their mean), frames per segment
  int latent_dim = 0;
  std::vector<__nv_bfloat16*> d_slot_hi, d_slot_lo;   // activation slots, each max_batch * n_seg images of the largest tensor
  float* d_lat = nullptr;          // [max_batch * n_seg][latent_dim] per-segment latents (n_seg > 1 or a non-linear head)

  bool conv1_tensor = false;       // bring-up builds: AVLD_CONV1_TENSOR at context creation sends the first conv to conv1t.cu

  // optional NCCL communicator of this context (comm.cu)
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 0;
  float* d_gather_r = nullptr;
  int32_t* d_gather_l = nullptr;
  int64_t gather_rows = 0;

  // host end-to-end path
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  float* d_xbuf[2] = {nullptr, nullptr};
  float* d_cent = nullptr;
  double* d_thr = nullptr;
  int32_t* d_prio = nullptr;
  int32_t* d_pred = nullptr;
  float* d_best = nullptr;
  std::string host_trace_path;     // AVLD_HOST_TRACE at context creation: per-slab timeline of the host path (tools/trace_run.py)
  void* h_stage = nullptr;         // pinned staging for results (a D2H into pageable memory would block the host
  size_t h_stage_bytes = 0;        // thread until the slab's kernels finish and serialise copy against compute)

  // resampling filter table of the last (sr_in, sr_out) pair (resample.cu)
  double* d_rs_win = nullptr;
  double* d_rs_delta = nullptr;
  int rs_sr_in = 0, rs_sr_out = 0;

  // order-statistics scratch
  unsigned int* d_hist = nullptr;
  size_t hist_bytes = 0;

  std::vector<const void*> smem_configured;   // kernels whose dynamic shared-memory limit was raised on this device

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
                   const uint64_t strides_bytes[3], const uint32_t box[4], uint32_t swizzle_bytes,
                   const uint32_t* elem_strides = nullptr);   // traversal strides (strided convolutions), default 1

// Opt-in to more than 48 KB of dynamic shared memory, once per (context, kernel).  The attribute belongs to the device the
// context lives on, so the record is kept in the context (a process-wide flag would leave the kernels of a second GPU's
// context without it).
int ensure_dyn_smem(avld_ctx* c, const void* kernel, int bytes);

// Every extern "C" entry point that touches the device starts with this: NULL check + cudaSetDevice(ctx's device), so a
// host thread that drives several contexts (one per GPU) always launches on the right one.
#define AVLD_ENTER(c)                                             \
  do {                                                            \
    AVLD_CHECK((c) != nullptr, AVLD_ERR_INVALID, "ctx is NULL");  \
    AVLD_CUDA(cudaSetDevice((c)->device));                        \
  } while (0)

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
int launch_stft_mel(avld_ctx* c, int n, cudaStream_t st);            // fold3 + GEMM
int launch_fold3(avld_ctx* c, int n, cudaStream_t st);
int launch_dftf3(avld_ctx* c, int n, cudaStream_t st);
int launch_logmel_post(avld_ctx* c, float* feat, int n, cudaStream_t st);
int launch_encoder(avld_ctx* c, const float* feat, float* mu, int n, cudaStream_t st);
int convh_encode_input_map(CUtensorMap* out, const void* base, int n, int H, int W, int C, int cblk);
bool convh_supported(int c_in, int c_out, int ksize, int w);
bool conv1t_supported(int ksize, int stride, int pad, int c_in, int c_out_pad, int h, int w, int pool);
int launch_conv1t(avld_ctx* c, const float* feat, const OpDev& L, int n, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t st);
int launch_convh(avld_ctx* c, const OpDev& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int n, __nv_bfloat16* out_hi,
                 __nv_bfloat16* out_lo, const __nv_bfloat16* res_hi, const __nv_bfloat16* res_lo, cudaStream_t st);
int launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t st);
int launch_split_f16(const float* src, __half* hi, __half* lo, size_t n, cudaStream_t st);

}  // namespace avld
