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

// max and NaN flag of the whole block in one exchange (NaN anywhere poisons the chunk exactly like numpy's max would)
__device__ __forceinline__ float block_max_poison(float mx, bool bad, float* s_buf) {
  float v = bad ? NAN : mx;
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = (v != v || u != u) ? NAN : fmaxf(u, v);
  }
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_buf[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = (r != r || s_buf[w] != s_buf[w]) ? NAN : fmaxf(r, s_buf[w]);
  return r;
}

// sum and sum of squares of the whole block (float64: sum v^2 - n mean^2 loses nothing at 2.4e4 elements of |v| <= 80)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* s_buf) {
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { s_buf[2 * (threadIdx.x >> 5)] = a; s_buf[2 * (threadIdx.x >> 5) + 1] = b; }
  __syncthreads();
  a = s_buf[0]; b = s_buf[1];
  for (int w = 1; w < (blockDim.x >> 5); ++w) { a += s_buf[2 * w]; b += s_buf[2 * w + 1]; }
}

// 10 log10(x) through the hardware log2 (MUFU.LG2: |error| < 2e-6 dB over the 100 dB this sees; the features carry the
// fp16-split GEMM's 1e-5 anyway); explicitly rounded so that `db(x) - db(ref)` is exactly 0 at the reference element
__device__ __forceinline__ float power_db(float x) { return __fmul_rn(3.01029995663981195f, __log2f(x)); }

// Three sweeps per chunk: (1) sum + clear the per-class planes into shared memory, ref = max; (2) dB, floor, float64 sum and
// sum of squares; (3) z-score + crop / pad + store.  log_spec.max() is exactly 0 (the dB of the reference element minus
// itself), so the floor is -top_db and needs no reduction of its own.
__global__ void __launch_bounds__(512, 2) logmel_post_kernel(const PostParams P) {
  extern __shared__ float s_db[];                 // [F*M]
  __shared__ double s_red_d[32];
  __shared__ float s_red_f[16];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int n = P.F * P.M;
  float* __restrict__ src = P.melpow + static_cast<size_t>(c) * P.R * P.M;   // frames 0..F-1 are contiguous

  float mx = -INFINITY;
  bool has_nan = false;
  const bool vec = (P.M & 3) == 0 && (P.plane2 & 3) == 0;
  if (P.n_planes == 3 && vec) {
    // folded STFT: sum the per-class planes in a fixed order and clear them for the next pass, 16 bytes at a time; the
    // nine loads of three steps are issued before their first use (the clears would otherwise order every load behind them)
    float4* src4 = reinterpret_cast<float4*>(src);
    const long long p4 = P.plane2 >> 2;
    const int n4 = n >> 2, stride = blockDim.x;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i0 = tid; i0 < n4; i0 += 3 * stride) {
      float4 a[3], b[3], d[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = i0 + u * stride;
        if (i < n4) { a[u] = __ldcs(src4 + i); b[u] = __ldcs(src4 + p4 + i); d[u] = __ldcs(src4 + 2 * p4 + i); }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = i0 + u * stride;
        if (i < n4) {
          float4 v = a[u];
          v.x += b[u].x; v.y += b[u].y; v.z += b[u].z; v.w += b[u].w;
          v.x += d[u].x; v.y += d[u].y; v.z += d[u].z; v.w += d[u].w;
          src4[i] = zero; src4[p4 + i] = zero; src4[2 * p4 + i] = zero;
          reinterpret_cast<float4*>(s_db)[i] = v;
          has_nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
          mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
      }
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
  mx = block_max_poison(mx, has_nan, s_red_f);      // (its barrier also publishes s_db)
  // an infinite power makes log_spec.max() NaN in numpy (inf - inf at that element): the whole chunk is NaN there too
  const bool poisoned = !(mx < INFINITY);

  // log_spec = 10*log10(max(amin, S)) - 10*log10(max(amin, ref)), floored at log_spec.max() - top_db = -top_db
  const float ref_db = power_db(fmaxf(P.amin, mx));
  const float floor_db = -P.top_db;
  double sum = 0.0, sq = 0.0;
  auto to_db = [&](float sv) -> float {
    const float v = fmaxf(__fsub_rn(power_db(fmaxf(P.amin, sv)), ref_db), floor_db);
    const double dv = static_cast<double>(v);
    sum += dv;
    sq = fma(dv, dv, sq);
    return v;
  };
  if (!poisoned) {
    if ((n & 3) == 0) {
      float4* s4 = reinterpret_cast<float4*>(s_db);
      for (int i = tid; i < (n >> 2); i += blockDim.x) {
        float4 v = s4[i];
        v.x = to_db(v.x); v.y = to_db(v.y); v.z = to_db(v.z); v.w = to_db(v.w);
        s4[i] = v;
      }
    } else {
      for (int i = tid; i < n; i += blockDim.x) s_db[i] = to_db(s_db[i]);
    }
  }
  block_sum2(sum, sq, s_red_d);                     // (barrier: s_db holds the dB values)
  const double mean_d = sum / n;
  const float mean = poisoned ? NAN : static_cast<float>(mean_d);
  const float sd = static_cast<float>(sqrt(fmax(sq / n - mean_d * mean_d, 0.0)));
  const float denom = sd + 1e-8f;                   // (S_db - mean) / (std + 1e-8)

  // crop_or_pad_time + transpose: feat[c][t][m] = z(S_db[t - pad_left + crop_start][m]); source and destination are both
  // frame-major, so the copy is one contiguous shift
  float* __restrict__ dst = P.feat + static_cast<size_t>(c) * P.T * P.M;
  const int total = P.T * P.M, lo = P.pad_left * P.M, hi = (P.pad_left + P.frames_copy) * P.M;
  const int shift = (P.crop_start - P.pad_left) * P.M;
  if (vec) {
    for (int i = 4 * tid; i < total; i += 4 * blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);   // np.pad(..., mode="constant") after the z-score
      if (i >= lo && i < hi) {
        const float4 sv = *reinterpret_cast<const float4*>(s_db + i + shift);
        v = make_float4(__fdiv_rn(sv.x - mean, denom), __fdiv_rn(sv.y - mean, denom), __fdiv_rn(sv.z - mean, denom), __fdiv_rn(sv.w - mean, denom));
      }
      __stcs(reinterpret_cast<float4*>(dst + i), v);
    }
  } else {
    for (int i = tid; i < total; i += blockDim.x) dst[i] = (i >= lo && i < hi) ? __fdiv_rn(s_db[i + shift] - mean, denom) : 0.f;
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
