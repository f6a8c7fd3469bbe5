// logmel.cu -- M2..M5 + E0: librosa.feature.melspectrogram / power_to_db(ref=np.max) / global
// z-score / centre crop (map_detector_core.py:219-237) and the [M,T] -> [T,M] transpose of
// map_detector_core.py:267-268.
//
//   prep_kernel (rms.cu)            audio -> per-chunk scale + normalised PCM_16 integers
//   fold3_kernel (fold3.cu)         windowed, three-times folded frames (fp16 hi / lo tiles)
//   dftf3_kernel (dftf3.cu)         folded DFT as a tcgen05 GEMM on CTA pairs; epilogue = |X|^2, un-scale,
//                                   sparse slaney-mel accumulation (<= 2 taps per FFT bin)
//   logmel_post_kernel              per chunk: ref = max, 10 log10, -top_db floor, mean/std over ALL
//                                   F frames (statistics before the crop), z-score, crop/pad, store
#include "common.cuh"

namespace avld {

// M2: folded operand (fold3.cu) -> CTA-pair GEMM with the |X|^2 / mel epilogue (dftf3.cu)
int launch_stft_mel(avld_ctx* c, int n, cudaStream_t st) {
  AVLD_TRY(launch_fold3(c, n, st));
  return launch_dftf3(c, n, st);
}

struct PostParams {
  float* melpow;        // [planes][rows][M]; the per-class planes are zeroed again right after they are read, so the
                        // GEMM epilogue can accumulate into them on the next pass without a separate memset
  long long plane2;     // stride between the per-class planes, added in a fixed order
  int n_planes;
  float* feat;          // [n][T][M]
  int R, F, M, T, crop_start, pad_left, frames_copy;
  float amin, top_db;
};

template <typename T>
__device__ __forceinline__ T block_reduce(T v, T* s_buf, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    const T u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (u > v ? u : v) : v + u;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = s_buf[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? (s_buf[w] > r ? s_buf[w] : r) : r + s_buf[w];
  return r;
}

__global__ void __launch_bounds__(512) logmel_post_kernel(const PostParams P) {
  extern __shared__ float s_db[];                 // [F*M]
  __shared__ double s_red_d[16];
  __shared__ float s_red_f[16];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int n = P.F * P.M;
  float* __restrict__ src = P.melpow + static_cast<size_t>(c) * P.R * P.M;   // frames 0..F-1 are contiguous

  // ref = np.max(S); NaN anywhere poisons the chunk exactly like numpy's max would
  float mx = -INFINITY;
  bool has_nan = false;
  if (P.plane2 && (P.M & 3) == 0 && (P.plane2 & 3) == 0) {
    // folded STFT: sum the per-class planes in a fixed order and clear them for the next pass, 16 bytes at a time
    float4* src4 = reinterpret_cast<float4*>(src);
    const long long p4 = P.plane2 >> 2;
    for (int i = tid; i < (n >> 2); i += blockDim.x) {
      float4 v = src4[i];
      for (int pl = 1; pl < P.n_planes; ++pl) {
        const float4 u = src4[pl * p4 + i];
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      }
      for (int pl = 0; pl < P.n_planes; ++pl) src4[pl * p4 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(s_db)[i] = v;
      has_nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
      mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  } else {
    for (int i = tid; i < n; i += blockDim.x) {
      float v = src[i];
      for (int pl = 1; pl < P.n_planes; ++pl) v += src[pl * P.plane2 + i];
      if (P.plane2) {
        for (int pl = 0; pl < P.n_planes; ++pl) src[pl * P.plane2 + i] = 0.f;
      }
      s_db[i] = v;
      has_nan |= (v != v);
      mx = fmaxf(mx, v);
    }
  }
  mx = block_reduce<float>(mx, s_red_f, true);
  const float nan_cnt = block_reduce<float>(has_nan ? 1.f : 0.f, s_red_f, false);
  if (nan_cnt > 0.f) mx = NAN;

  // log_spec = 10*log10(max(amin, S)) - 10*log10(max(amin, ref)); then max over log_spec
  const float ref_db = 10.0f * log10f(fmaxf(P.amin, mx));
  float mx_db = -INFINITY;
  for (int i = tid; i < n; i += blockDim.x) {
    const float v = 10.0f * log10f(fmaxf(P.amin, s_db[i])) - ref_db;
    s_db[i] = v;
    mx_db = fmaxf(mx_db, v);
  }
  mx_db = block_reduce<float>(mx_db, s_red_f, true);
  if (nan_cnt > 0.f) mx_db = NAN;
  const float floor_db = mx_db - P.top_db;         // np.maximum(log_spec, log_spec.max() - top_db)

  double sum = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    float v = s_db[i];
    v = (v != v || floor_db != floor_db) ? NAN : fmaxf(v, floor_db);
    s_db[i] = v;
    sum += static_cast<double>(v);
  }
  sum = block_reduce<double>(sum, s_red_d, false);
  const float mean = static_cast<float>(sum / n);
  double ss = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double d = static_cast<double>(s_db[i]) - static_cast<double>(mean);
    ss += d * d;
  }
  ss = block_reduce<double>(ss, s_red_d, false);
  const float sd = static_cast<float>(sqrt(ss / n));
  const float denom = sd + 1e-8f;                   // (S_db - mean) / (std + 1e-8)

  // crop_or_pad_time + transpose: feat[c][t][m]
  float* __restrict__ dst = P.feat + static_cast<size_t>(c) * P.T * P.M;
  const int total = P.T * P.M;
  for (int i = tid; i < total; i += blockDim.x) {
    const int t = i / P.M, m = i - t * P.M;
    const int f = t - P.pad_left + P.crop_start;
    float v = 0.f;                                  // np.pad(..., mode="constant") after the z-score
    if (t >= P.pad_left && t < P.pad_left + P.frames_copy) v = (s_db[f * P.M + m] - mean) / denom;
    dst[i] = v;
  }
}

int launch_logmel_post(avld_ctx* c, float* feat, int n, cudaStream_t st) {
  if (n <= 0) return AVLD_OK;
  PostParams P{c->d_melpow, c->melpow_plane, c->f2_classes, feat, c->F, c->F, c->M, c->T, c->crop_start, c->pad_left, c->frames_copy, c->p.amin, c->p.top_db};
  const size_t smem = static_cast<size_t>(c->F) * c->M * sizeof(float);
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(logmel_post_kernel), 200 * 1024));
  { LaunchScope ls(c, ST_LOGMEL_POST, st); logmel_post_kernel<<<n, 512, smem, st>>>(P); }
  c->planes_dirty = false;
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld

using namespace avld;

static int features_pass(avld_ctx* c, const float* x, float* feat, uint8_t* ok, float* rms, int64_t n, bool normalize,
                         float target_rms, float rms_min, float eps, int quantize, cudaStream_t st) {
  AVLD_CHECK(c->features_ok, AVLD_ERR_UNSUPPORTED,
             "chunk_len %d: features need n_fft/2 < chunk_len and frames x mels x 4 <= 200 KB (got %d frames)", c->L, c->F);
  for (int64_t i = 0; i < n; i += c->max_batch) {
    const int m = static_cast<int>(n - i < c->max_batch ? n - i : c->max_batch);
    AVLD_TRY(launch_prep(c, x + i * c->L, nullptr, nullptr, true, normalize, ok ? ok + i : nullptr, rms ? rms + i : nullptr, m,
                         target_rms, rms_min, eps, quantize, st));
    AVLD_TRY(launch_stft_mel(c, m, st));
    AVLD_TRY(launch_logmel_post(c, feat + i * c->T * c->M, m, st));
  }
  return AVLD_OK;
}

extern "C" int avld_logmel(avld_ctx* c, const float* y, float* feat, int64_t n, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(y && feat, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0, AVLD_ERR_INVALID, "negative n");
  return features_pass(c, y, feat, nullptr, nullptr, n, false, 0.f, 0.f, 0.f, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int avld_normalize_logmel(avld_ctx* c, const float* x, float* feat, uint8_t* ok, float* rms, int64_t n,
                                     float target_rms, float rms_min, float eps, int quantize_pcm16, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(x && feat, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0, AVLD_ERR_INVALID, "negative n");
  return features_pass(c, x, feat, ok, rms, n, true, target_rms, rms_min, eps, quantize_pcm16,
                       static_cast<cudaStream_t>(stream));
}
