// ctx.cu -- context life cycle, host-side tables (numpy pairwise-sum plan, slaney mel taps,
// windowed DFT matrix) and TMA descriptor encoding.
#include <nvtx3/nvToolsExt.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace avld {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// launch accounting / per-stage timing
// ------------------------------------------------------------------------------------------------
static cudaEvent_t take_event(avld_ctx* c) {
  if (!c->event_pool.empty()) {
    cudaEvent_t e = c->event_pool.back();
    c->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// Every kernel launch of the library sits inside an NVTX range named after its stage (header-only NVTX3: a no-op of a few
// nanoseconds unless a tool such as `ncu --nvtx --nvtx-include "avld/dftf3_kernel/"` or Nsight Systems is attached).
LaunchScope::LaunchScope(avld_ctx* ctx, int stage, cudaStream_t stream) : c(ctx), st(stream) {
  nvtxRangePushA(avld_stage_name(stage));
  c->launches[stage] += 1;
  if (c->profiling) {
    cudaEvent_t a = take_event(c);
    stop = take_event(c);
    cudaEventRecord(a, st);
    c->timed.push_back({stage, a, stop});
  }
}
LaunchScope::~LaunchScope() {
  if (stop) cudaEventRecord(stop, st);
  nvtxRangePop();
}

int ensure_dyn_smem(avld_ctx* c, const void* kernel, int bytes) {
  for (const void* k : c->smem_configured)
    if (k == kernel) return AVLD_OK;
  AVLD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  c->smem_configured.push_back(kernel);
  return AVLD_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static CUtensorMapSwizzle swizzle_enum(uint32_t bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                   : bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, uint64_t dim0, uint64_t dim1,
                   uint64_t stride1_bytes, uint32_t box0, uint32_t box1, uint32_t swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return AVLD_ERR_CUDA;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVLD_CHECK(r == CUDA_SUCCESS, AVLD_ERR_CUDA,
             "cuTensorMapEncodeTiled(2d) failed with CUresult %d (dims %llu x %llu, stride %llu, box %u x %u)", (int)r,
             (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)stride1_bytes, box0, box1);
  return AVLD_OK;
}

int encode_tmap_4d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, const uint64_t dims_[4],
                   const uint64_t strides_bytes[3], const uint32_t box_[4], uint32_t swizzle_bytes,
                   const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return AVLD_ERR_CUDA;
  cuuint64_t dims[4] = {dims_[0], dims_[1], dims_[2], dims_[3]};
  cuuint64_t strides[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t box[4] = {box_[0], box_[1], box_[2], box_[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < 4; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(out, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AVLD_CHECK(r == CUDA_SUCCESS, AVLD_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
  return AVLD_OK;
}

// ------------------------------------------------------------------------------------------------
// numpy pairwise summation plan (numpy/core/src/umath/loops_utils.h.src::pairwise_sum):
//   n < 8: plain loop; n <= 128: 8 interleaved accumulators; else split at n/2 rounded down to a
//   multiple of 8.  This is the tree np.mean(y**2) walks in 00_normalize_dataset_rms.py:30.
// ------------------------------------------------------------------------------------------------
namespace {
struct PlanBuilder {
  std::vector<int64_t>* off;
  std::vector<int64_t>* len;
  struct Raw { int l, r, height; };
  std::vector<Raw> raw;  // internal nodes; child id >= 0 -> leaf id, < 0 -> ~(raw index)
  int build(int64_t o, int64_t n, int* height) {
    if (n <= 128) {
      off->push_back(o);
      len->push_back(n);
      *height = 0;
      return static_cast<int>(off->size()) - 1;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    int hl, hr;
    const int l = build(o, n2, &hl);
    const int r = build(o + n2, n - n2, &hr);
    *height = (hl > hr ? hl : hr) + 1;
    raw.push_back({l, r, *height});
    return ~static_cast<int>(raw.size() - 1);
  }
};
}  // namespace

int64_t pairwise_plan(int64_t n, std::vector<int64_t>& off, std::vector<int64_t>& len, std::vector<PairNode>& nodes,
                      std::vector<int32_t>& level_start) {
  off.clear();
  len.clear();
  nodes.clear();
  level_start.clear();
  PlanBuilder b{&off, &len, {}};
  int height = 0;
  b.build(0, n, &height);
  const int n_leaves = static_cast<int>(off.size());
  // order internal nodes by height so that each level only depends on earlier values
  std::vector<int> new_index(b.raw.size(), -1);
  level_start.push_back(0);
  for (int h = 1; h <= height; ++h) {
    for (size_t i = 0; i < b.raw.size(); ++i)
      if (b.raw[i].height == h) new_index[i] = static_cast<int>(nodes.size()), nodes.push_back({0, 0});
    level_start.push_back(static_cast<int32_t>(nodes.size()));
  }
  for (size_t i = 0; i < b.raw.size(); ++i) {
    auto resolve = [&](int c) { return c >= 0 ? c : n_leaves + new_index[~c]; };
    nodes[new_index[i]] = {resolve(b.raw[i].l), resolve(b.raw[i].r)};
  }
  return n_leaves;
}

// ------------------------------------------------------------------------------------------------
// slaney mel filterbank (librosa.filters.mel, norm="slaney", htk=False), reduced to <= 2 taps per bin
// ------------------------------------------------------------------------------------------------
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

int mel_taps_host(const avld_params& p, std::vector<int32_t>& first, std::vector<float>& w0, std::vector<float>& w1,
                  int* bin_lo, int* bin_hi) {
  const int nb = p.n_fft / 2 + 1, nm = p.n_mels;
  AVLD_CHECK(nm >= 1 && nb >= 2, AVLD_ERR_INVALID, "invalid n_mels / n_fft");
  std::vector<double> mel_f(nm + 2), fftf(nb);
  const double m0 = hz_to_mel(p.fmin), m1 = hz_to_mel(p.fmax);
  const double mstep = (m1 - m0) / (nm + 1);
  for (int i = 0; i < nm + 2; ++i) mel_f[i] = mel_to_hz(i == nm + 1 ? m1 : m0 + mstep * i);  // np.linspace
  const double fstep = (p.sr / 2.0) / (nb - 1);
  for (int b = 0; b < nb; ++b) fftf[b] = (b == nb - 1) ? p.sr / 2.0 : fstep * b;
  first.assign(nb, -1);
  w0.assign(nb, 0.f);
  w1.assign(nb, 0.f);
  *bin_lo = -1;
  *bin_hi = -1;
  int prev_first = 0;
  for (int b = 0; b < nb; ++b) {
    int cnt = 0, f0 = -1;
    float ww[2] = {0.f, 0.f};
    for (int m = 0; m < nm; ++m) {
      const double lower = -(mel_f[m] - fftf[b]) / (mel_f[m + 1] - mel_f[m]);
      const double upper = (mel_f[m + 2] - fftf[b]) / (mel_f[m + 2] - mel_f[m + 1]);
      const double tri = std::fmax(0.0, std::fmin(lower, upper));
      float wf = static_cast<float>(tri);                                             // weights[i] = ... (float32 row)
      wf = static_cast<float>(static_cast<double>(wf) * (2.0 / (mel_f[m + 2] - mel_f[m])));  // weights *= enorm
      if (wf != 0.f) {
        if (cnt == 0) f0 = m;
        AVLD_CHECK(m - f0 <= 1, AVLD_ERR_UNSUPPORTED, "FFT bin %d feeds non-adjacent mel filters", b);
        ww[m - f0] = wf;
        ++cnt;
      }
    }
    if (cnt > 0) {
      AVLD_CHECK(f0 >= prev_first, AVLD_ERR_UNSUPPORTED, "mel filter order is not monotone at bin %d", b);
      first[b] = f0;
      w0[b] = ww[0];
      w1[b] = ww[1];
      prev_first = f0;
      if (*bin_lo < 0) *bin_lo = b;
      *bin_hi = b;
    }
  }
  AVLD_CHECK(*bin_lo >= 0, AVLD_ERR_INVALID, "mel filterbank is empty for these parameters");
  return AVLD_OK;
}

}  // namespace avld

using namespace avld;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int avld_abi_version(void) { return AVLD_ABI_VERSION; }
extern "C" const char* avld_last_error(void) { return g_err; }

extern "C" int64_t avld_pairwise_plan(int64_t n, int64_t* leaf_offset, int64_t* leaf_len, int64_t cap) {
  if (n < 0) {
    set_error("negative length");
    return AVLD_ERR_INVALID;
  }
  std::vector<int64_t> off, len;
  std::vector<PairNode> nodes;
  std::vector<int32_t> ls;
  const int64_t nl = pairwise_plan(n, off, len, nodes, ls);
  for (int64_t i = 0; i < nl && i < cap; ++i) {
    if (leaf_offset) leaf_offset[i] = off[i];
    if (leaf_len) leaf_len[i] = len[i];
  }
  return nl;
}

extern "C" int avld_mel_taps(const avld_params* params, int32_t* first, float* w0, float* w1) {
  AVLD_CHECK(params != nullptr, AVLD_ERR_INVALID, "params is NULL");
  std::vector<int32_t> f;
  std::vector<float> a, b;
  int lo, hi;
  AVLD_TRY(mel_taps_host(*params, f, a, b, &lo, &hi));
  for (size_t i = 0; i < f.size(); ++i) {
    if (first) first[i] = f[i];
    if (w0) w0[i] = a[i];
    if (w1) w1[i] = b[i];
  }
  return AVLD_OK;
}

template <typename T>
static int dev_alloc(T** p, size_t count) {
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return AVLD_OK;
}

static int build_ctx(avld_ctx* c) {
  const avld_params& p = c->p;
  AVLD_CHECK(p.sr > 0 && p.n_fft >= 64 && p.hop > 0 && p.n_mels > 0 && p.target_frames > 0 && p.chunk_len > 0,
             AVLD_ERR_INVALID, "non-positive feature parameter");
  AVLD_CHECK(p.n_fft % 512 == 0, AVLD_ERR_UNSUPPORTED,
             "n_fft must be a multiple of 512 for the folded STFT GEMM (got %d)", p.n_fft);
  AVLD_CHECK(p.hop % 8 == 0, AVLD_ERR_UNSUPPORTED, "hop_length must be a multiple of 8 (got %d)", p.hop);
  AVLD_CHECK(p.n_mels <= 256, AVLD_ERR_UNSUPPORTED, "n_mels > 256");
  AVLD_CHECK(p.amin > 0.f && p.top_db >= 0.f, AVLD_ERR_INVALID, "amin must be > 0 and top_db >= 0");
  AVLD_CHECK(p.max_batch >= 1, AVLD_ERR_INVALID, "max_batch must be >= 1");
  AVLD_CHECK(c->sm_count % 2 == 0, AVLD_ERR_UNSUPPORTED, "the STFT GEMM runs on CTA pairs: odd SM count %d", c->sm_count);
  c->L = p.chunk_len;
  c->F = 1 + p.chunk_len / p.hop;
  c->T = p.target_frames;
  c->M = p.n_mels;
  c->max_batch = p.max_batch;
  if (c->F >= c->T) {
    c->crop_start = (c->F - c->T) / 2;
    c->pad_left = 0;
    c->frames_copy = c->T;
  } else {
    c->crop_start = 0;
    c->pad_left = (c->T - c->F) / 2;
    c->frames_copy = c->F;
  }
  // reflect padding needs L > n_fft/2 (as librosa does); the log-mel kernel keeps a whole chunk's [F, M] in shared memory
  c->features_ok = p.chunk_len > p.n_fft / 2 && static_cast<size_t>(c->F) * c->M * sizeof(float) <= 200 * 1024;

  // ---- pairwise plan
  std::vector<int64_t> off, len;
  std::vector<PairNode> nodes;
  std::vector<int32_t> level_start;
  c->n_leaves = static_cast<int>(pairwise_plan(c->L, off, len, nodes, level_start));
  c->n_nodes = static_cast<int>(nodes.size());
  c->n_levels = static_cast<int>(level_start.size()) - 1;
  AVLD_CHECK(static_cast<size_t>(c->n_leaves + c->n_nodes) * 4 <= 160 * 1024, AVLD_ERR_UNSUPPORTED, "chunk_len too large");
  std::vector<int32_t> off32(off.begin(), off.end()), len32(len.begin(), len.end());
  c->leaves_regular = c->L % 8 == 0;
  for (size_t i = 0; i < off32.size(); ++i)
    c->leaves_regular = c->leaves_regular && off32[i] % 8 == 0 && len32[i] % 8 == 0 && len32[i] >= 8 && len32[i] <= 128;
  c->min_leaf_rows = 1 << 30;
  for (int32_t v : len32) c->min_leaf_rows = std::min(c->min_leaf_rows, v / 8);
  c->tree_perfect = c->n_leaves >= 256 && (c->n_leaves & (c->n_leaves - 1)) == 0 && (1 << c->n_levels) == c->n_leaves;
  AVLD_TRY(dev_alloc(&c->d_leaf_off, off32.size()));
  AVLD_TRY(dev_alloc(&c->d_leaf_len, len32.size()));
  AVLD_TRY(dev_alloc(&c->d_nodes, nodes.size() + 1));
  AVLD_TRY(dev_alloc(&c->d_level_start, level_start.size()));
  AVLD_CUDA(cudaMemcpy(c->d_leaf_off, off32.data(), off32.size() * 4, cudaMemcpyHostToDevice));
  AVLD_CUDA(cudaMemcpy(c->d_leaf_len, len32.data(), len32.size() * 4, cudaMemcpyHostToDevice));
  if (!nodes.empty())
    AVLD_CUDA(cudaMemcpy(c->d_nodes, nodes.data(), nodes.size() * sizeof(PairNode), cudaMemcpyHostToDevice));
  AVLD_CUDA(cudaMemcpy(c->d_level_start, level_start.data(), level_start.size() * 4, cudaMemcpyHostToDevice));

  // ---- mel taps, restricted to the FFT bins that carry weight
  std::vector<int32_t> first;
  std::vector<float> w0, w1;
  int bin_lo, bin_hi;
  AVLD_TRY(mel_taps_host(p, first, w0, w1, &bin_lo, &bin_hi));

  // ---- folded operands: three bin classes, each covered by work items of 160 bins; item rows of the DFT matrix B3 =
  // 160 cos rows then 160 sin rows, Q = N/4 taps wide (the classes = 0 / 2 mod 4 use the first N/8), no window (it is
  // applied to the frames by fold3_kernel), scaled by 2^dft_scale_log2 to keep the lo parts out of the fp16 subnormals
  const int nf = p.n_fft, half = nf / 2, Q = nf / 4, kItem = 160;
  const double bscale = std::ldexp(1.0, c->dft_scale_log2);
  struct ClassDef { int mod, rem, a_col0, K, edge_im; };
  const ClassDef classes[3] = {{2, 1, 0, Q, 1},                 // odd bins: K = N/4, edge O[N/4] sin(pi b / 2) -> Im
                               {4, 0, half, Q / 2, 0},          // b = 0 mod 4: K = N/8, edge P[N/8] cos(pi b / 4) -> Re
                               {4, 2, half + Q, Q / 2, 1}};     // b = 2 mod 4: K = N/8, edge R[N/8] sin(pi b / 4) -> Im
  c->f2_classes = 3;
  std::vector<std::vector<int>> item_bins;
  c->f2_items = 0;
  for (int ci = 0; ci < 3; ++ci) {
    std::vector<int> bins;
    for (int b = bin_lo; b <= bin_hi; ++b)
      if (b % classes[ci].mod == classes[ci].rem) bins.push_back(b);
    for (size_t o = 0; o < bins.size(); o += kItem) {
      AVLD_CHECK(c->f2_items < avld_ctx::kMaxItems, AVLD_ERR_UNSUPPORTED,
                 "too many FFT bins carry mel weight for the STFT kernel (%d bins, limit %d)", bin_hi - bin_lo + 1,
                 avld_ctx::kMaxItems * kItem);
      c->f2_item[c->f2_items++] = {classes[ci].a_col0, classes[ci].K / 64, ci, classes[ci].edge_im};
      item_bins.emplace_back(bins.begin() + o, bins.begin() + std::min(bins.size(), o + kItem));
    }
  }
  const size_t rows3 = static_cast<size_t>(c->f2_items) * 2 * kItem;
  std::vector<__half> h3(rows3 * Q), l3(rows3 * Q);
  std::vector<MelTap> taps3(static_cast<size_t>(c->f2_items) * kItem);
  for (int it = 0; it < c->f2_items; ++it) {
    const ClassDef& cd = classes[c->f2_item[it].cls];
    const std::vector<int>& bins = item_bins[it];
    int run = 0;
    for (size_t u = 0; u < bins.size(); ++u)             // `first` must be monotone from the item's first column on
      if (first[bins[u]] >= 0) { run = first[bins[u]]; break; }
    for (int j = 0; j < kItem; ++j) {
      const bool have = j < static_cast<int>(bins.size());
      const int bin = have ? bins[j] : -1;
      // epilogue taps: mel filter pair of the bin and the coefficient of the class's self-paired tap:
      // cos(pi b / 2), sin(pi b / 2), cos(pi b / 4) or sin(pi b / 4) at the class's bins, all +-1
      MelTap tp{run, 0.f, 0.f, 0};
      if (have && first[bin] >= 0) {
        run = first[bin];
        const double ang = M_PI * bin / (cd.K == Q ? 2.0 : 4.0);
        const double cf = cd.edge_im ? std::sin(ang) : std::cos(ang);
        const float coef = static_cast<float>(bscale * std::round(cf));
        int32_t bits;
        memcpy(&bits, &coef, 4);
        tp = {first[bin], w0[bin], w1[bin], bits};
      }
      taps3[static_cast<size_t>(it) * kItem + j] = tp;
      for (int part = 0; part < 2; ++part) {
        const size_t r = (static_cast<size_t>(it) * 2 + part) * kItem + j;
        for (int k = 0; k < Q; ++k) {
          double v = 0.0;
          if (have && k < cd.K) {
            const long long ph = (static_cast<long long>(k) * bin) % nf;
            const double ang = 2.0 * M_PI * static_cast<double>(ph) / nf;
            v = bscale * (part == 0 ? std::cos(ang) : std::sin(ang));
          }
          const __half h = __float2half_rn(static_cast<float>(v));
          h3[r * Q + k] = h;
          l3[r * Q + k] = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(h))));
        }
      }
    }
  }
  AVLD_TRY(dev_alloc(&c->d_B3hi, h3.size()));
  AVLD_TRY(dev_alloc(&c->d_B3lo, l3.size()));
  AVLD_CUDA(cudaMemcpy(c->d_B3hi, h3.data(), h3.size() * 2, cudaMemcpyHostToDevice));
  AVLD_CUDA(cudaMemcpy(c->d_B3lo, l3.data(), l3.size() * 2, cudaMemcpyHostToDevice));
  AVLD_TRY(encode_tmap_2d(&c->tm_B3_hi, c->d_B3hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Q, rows3, static_cast<uint64_t>(Q) * 2, 64, kItem / 2, 128));
  AVLD_TRY(encode_tmap_2d(&c->tm_B3_lo, c->d_B3lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Q, rows3, static_cast<uint64_t>(Q) * 2, 64, kItem / 2, 128));
  AVLD_TRY(dev_alloc(&c->d_taps3, taps3.size()));
  AVLD_CUDA(cudaMemcpy(c->d_taps3, taps3.data(), taps3.size() * sizeof(MelTap), cudaMemcpyHostToDevice));

  // ---- folded frames A3, tile-major (fold3.cu): [tile of 128 frames][K block][hi | lo][128][64] fp16; one (tile, K block)
  // = 256 rows of 128 bytes = one 32 KB box of the tensor map below
  c->a3_kblocks = nf / 64;
  const size_t frames = static_cast<size_t>(c->max_batch) * c->F + 256;
  c->a3_tiles = (frames + 127) / 128;
  const size_t a3_rows = c->a3_tiles * c->a3_kblocks * 256;
  AVLD_CHECK(a3_rows < (1ull << 31), AVLD_ERR_UNSUPPORTED, "max_batch too large for the folded operand's tensor map");
  AVLD_TRY(dev_alloc(&c->d_A3, a3_rows * 64));
  AVLD_CUDA(cudaMemset(c->d_A3, 0, a3_rows * 64 * 2));
  AVLD_TRY(encode_tmap_2d(&c->tm_A3, c->d_A3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 64, a3_rows, 128, 64, 256, 128));
  AVLD_TRY(dev_alloc(&c->d_chunk_par, c->max_batch));
  // periodic Hann [half + 1], then the same values regrouped per 8-tap block kb of fold3_kernel as nine float4 planes
  // [9][E / 8]: w[k] (2), w[H-k] (2), w[Q-k] (2), w[Q+k] (2) for k = 8 kb + 0..7, and (w[Q-k], w[Q+k]) at k = 8 kb + 8 --
  // one coalesced 16-byte load per plane instead of 34 strided scalar loads per thread
  {
    const int wt_off = (half + 1 + 3) & ~3, wt_blocks = nf / 64;
    std::vector<float> win(wt_off + 9 * wt_blocks * 4, 0.f);
    for (int k = 0; k <= half; ++k) win[k] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * M_PI * k / nf));   // scipy get_window('hann', n, fftbins=True)
    for (int kb = 0; kb < wt_blocks; ++kb) {
      auto cell = [&](int plane, int j) -> float& { return win[wt_off + (plane * wt_blocks + kb) * 4 + j]; };
      for (int q = 0; q < 8; ++q) {
        const int k = kb * 8 + q;
        cell(0 + q / 4, q % 4) = win[k];
        cell(2 + q / 4, q % 4) = win[half - k];
        cell(4 + q / 4, q % 4) = win[Q - k];
        cell(6 + q / 4, q % 4) = win[Q + k];
      }
      cell(8, 0) = win[Q - (kb * 8 + 8)];
      cell(8, 1) = win[Q + (kb * 8 + 8)];
    }
    AVLD_TRY(dev_alloc(&c->d_win, win.size()));
    AVLD_CUDA(cudaMemcpy(c->d_win, win.data(), win.size() * 4, cudaMemcpyHostToDevice));
  }
  // AVLD_NO_Q16 (read here, once): PCM_16 round-trip passes then re-normalise every sample inside fold3_kernel instead of
  // reading prep_kernel's integers -- the same bits by construction, kept selectable so that a test can prove it
  if (std::getenv("AVLD_NO_Q16") == nullptr) AVLD_TRY(dev_alloc(&c->d_q16, static_cast<size_t>(p.max_batch) * c->L + 64));
  AVLD_TRY(dev_alloc(&c->d_edge, frames));
  AVLD_CUDA(cudaMemset(c->d_edge, 0, frames * sizeof(float4)));

  // ---- per-pass scratch
  AVLD_TRY(dev_alloc(&c->d_inv2, c->max_batch));
  c->melpow_plane = static_cast<long long>(frames) * c->M;
  AVLD_TRY(dev_alloc(&c->d_melpow, static_cast<size_t>(c->melpow_plane) * c->f2_classes));
  AVLD_CUDA(cudaMemset(c->d_melpow, 0, static_cast<size_t>(c->melpow_plane) * c->f2_classes * sizeof(float)));
  AVLD_TRY(dev_alloc(&c->d_feat, static_cast<size_t>(c->max_batch) * c->T * c->M));
  AVLD_TRY(dev_alloc(&c->d_ok, c->max_batch));
  AVLD_TRY(dev_alloc(&c->d_rms, c->max_batch));
  AVLD_TRY(dev_alloc(&c->d_radii, static_cast<size_t>(c->max_batch) * 64));
  AVLD_TRY(dev_alloc(&c->d_pred, c->max_batch));
  AVLD_TRY(dev_alloc(&c->d_best, c->max_batch));
  AVLD_CUDA(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
  AVLD_CUDA(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    AVLD_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
    AVLD_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
  }
  return AVLD_OK;
}

extern "C" int avld_ctx_create(int device, const avld_params* params, avld_ctx** out) {
  AVLD_CHECK(params != nullptr && out != nullptr, AVLD_ERR_INVALID, "NULL argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  AVLD_CHECK(e == cudaSuccess && ndev > 0, AVLD_ERR_CUDA,
             "no CUDA device available (%s); libavld has no CPU fallback", cudaGetErrorString(e));
  AVLD_CHECK(device >= 0 && device < ndev, AVLD_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
  AVLD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AVLD_CUDA(cudaGetDeviceProperties(&prop, device));
  AVLD_CHECK(prop.major == 10, AVLD_ERR_UNSUPPORTED,
             "libavld is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
  avld_ctx* c = new avld_ctx();
  c->device = device;
  c->p = *params;
  c->sm_count = prop.multiProcessorCount;
  if (const char* t = getenv("AVLD_HOST_TRACE")) c->host_trace_path = t;
#ifdef AVLD_BRINGUP
  c->conv1_tensor = getenv("AVLD_CONV1_TENSOR") != nullptr;
#endif
  c->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  const int r = build_ctx(c);
  if (r != AVLD_OK) {
    avld_ctx_destroy(c);
    return r;
  }
  *out = c;
  return AVLD_OK;
}

extern "C" int avld_comm_destroy(avld_ctx* c);

extern "C" void avld_ctx_destroy(avld_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  avld_comm_destroy(c);
  void* ptrs[] = {c->d_leaf_off, c->d_leaf_len, c->d_nodes, c->d_level_start, c->d_chunk_par, c->d_A3, c->d_B3hi, c->d_B3lo,
                  c->d_taps3, c->d_edge, c->d_q16, c->d_win, c->d_inv2, c->d_melpow, c->d_feat, c->d_mu, c->d_radii, c->d_ok,
                  c->d_rms, c->d_lat, c->d_rs_win, c->d_rs_delta, c->d_xbuf[0], c->d_xbuf[1], c->d_cent, c->d_thr, c->d_prio, c->d_pred, c->d_best, c->d_hist};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto* p : c->d_slot_hi)
    if (p) cudaFree(p);
  for (auto* p : c->d_slot_lo)
    if (p) cudaFree(p);
  for (auto& l : c->ops) {
    if (l.w_f32) cudaFree(l.w_f32);
    if (l.bias) cudaFree(l.bias);
    if (l.w_hi) cudaFree(l.w_hi);
    if (l.w_lo) cudaFree(l.w_lo);
  }
  if (c->s_compute) cudaStreamDestroy(c->s_compute);
  if (c->s_copy) cudaStreamDestroy(c->s_copy);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
  }
  if (c->h_stage) cudaFreeHost(c->h_stage);
  for (auto& t : c->timed) {
    cudaEventDestroy(t.start);
    cudaEventDestroy(t.stop);
  }
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  delete c;
}

extern "C" int avld_ctx_set_normalization(avld_ctx* c, int scalar_semantics, double target_rms, double rms_min, double eps) {
  AVLD_ENTER(c);
  AVLD_CHECK(scalar_semantics == 0 || scalar_semantics == 1, AVLD_ERR_INVALID, "scalar_semantics must be 0 (numpy 2) or 1 (numpy 1.x)");
  AVLD_CHECK(target_rms > 0.0 && rms_min >= 0.0 && eps >= 0.0, AVLD_ERR_INVALID, "bad normalisation constants");
  c->scalar_f64 = scalar_semantics;
  c->norm_target = target_rms;
  c->norm_rms_min = rms_min;
  c->norm_eps = eps;
  return AVLD_OK;
}

extern "C" int avld_profile_enable(avld_ctx* c, int on) {
  AVLD_ENTER(c);
  c->profiling = on != 0;
  return AVLD_OK;
}

extern "C" int avld_profile_collect(avld_ctx* c, double* ms, int64_t* timed_launches, uint64_t* launches, int reset) {
  AVLD_ENTER(c);
  for (int i = 0; i < ST_COUNT; ++i) {
    if (ms) ms[i] = 0.0;
    if (timed_launches) timed_launches[i] = 0;
    if (launches) launches[i] = c->launches[i];
  }
  for (auto& t : c->timed) {
    AVLD_CUDA(cudaEventSynchronize(t.stop));
    float e = 0.f;
    AVLD_CUDA(cudaEventElapsedTime(&e, t.start, t.stop));
    if (ms) ms[t.stage] += e;
    if (timed_launches) timed_launches[t.stage] += 1;
    c->event_pool.push_back(t.start);
    c->event_pool.push_back(t.stop);
  }
  c->timed.clear();
  if (reset)
    for (int i = 0; i < ST_COUNT; ++i) c->launches[i] = 0;
  return AVLD_OK;
}

extern "C" int avld_stage_count(void) { return ST_COUNT; }
extern "C" const char* avld_stage_name(int stage) {
  static const char* names[ST_COUNT] = {"prep_kernel", "dftf3_kernel", "logmel_post_kernel", "conv1_kernel", "convh_kernel",
                                        "gemm3_kernel", "radii_kernel", "decide_kernel", "centroid_kernel", "select_hist_kernel",
                                        "split_kernel", "fold3_kernel", "map_kernels", "encoder_elementwise_kernels"};
  return (stage >= 0 && stage < ST_COUNT) ? names[stage] : "?";
}

extern "C" int avld_ctx_dft_info(const avld_ctx* c, const char** mode, double* algorithmic, double* issued) {
  AVLD_ENTER(c);
  const double F = c->F, N = c->p.n_fft;
  int first_bin = 0, bins = 0;
  {
    std::vector<int32_t> first;
    std::vector<float> w0, w1;
    int lo = 0, hi = 0;
    AVLD_TRY(mel_taps_host(c->p, first, w0, w1, &lo, &hi));
    for (int b = lo; b <= hi; ++b) bins += first[b] >= 0;
    first_bin = lo;
  }
  (void)first_bin;
  if (algorithmic) *algorithmic = 2.0 * F * N * 2.0 * bins;
  // items x (cos + sin) x K x 160 columns, three split-precision passes
  const char* m = "fold3";
  double iss = 0.0;
  for (int it = 0; it < c->f2_items; ++it) iss += 3.0 * 2.0 * F * 2.0 * (c->f2_item[it].kbp * 64.0) * 160.0;
  if (mode) *mode = m;
  if (issued) *issued = iss;
  return AVLD_OK;
}

extern "C" int avld_ctx_info(const avld_ctx* c, int32_t* n_frames, int32_t* latent_dim, int32_t* sm_count) {
  AVLD_ENTER(c);
  if (n_frames) *n_frames = c->F;
  if (latent_dim) *latent_dim = c->latent_dim;
  if (sm_count) *sm_count = c->sm_count;
  return AVLD_OK;
}
