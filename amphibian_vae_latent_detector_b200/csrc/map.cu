// map.cu -- row N1: Gaussian-MAP detector on latents.
//   map_detector_core.py:319-323   gaussian_logpdf_from_precision: -0.5 (d^T P d + logdet + D ln 2pi)
//   09n_evaluate_wav_detection.py:114-140, 10b_benchmark_folder_detection_map.py:146-169: + ln(prior + 1e-12),
//       argmax over species in sorted-name order with a strict '>', reject when best < tau
//   08b_fit_map_detector.py:60-81, :276-296: class / pooled covariance of centred latents (np.cov, float64)
#include <algorithm>

#include "common.cuh"

namespace avld {

// ------------------------------------------------------------------------------------------------
// scoring: one warp = 4 latents; the precision matrix row i is read once (coalesced, L2 resident) and applied to
// the 4 difference vectors held in shared memory.  quad is accumulated in float32 like the reference's
// `diff.T @ prec @ diff`, everything after it in float64 like the reference's Python floats.
// ------------------------------------------------------------------------------------------------
constexpr int kMapRows = 4;       // latents per warp
constexpr int kMapMaxQ = 8;       // D <= 256

__global__ void __launch_bounds__(128) map_score_kernel(const float* __restrict__ Z, const float* __restrict__ mean,
                                                        const float* __restrict__ prec, const double* __restrict__ a_const,
                                                        const double* __restrict__ log_prior, double tau, int use_tau,
                                                        int32_t* __restrict__ pred, double* __restrict__ best,
                                                        double* __restrict__ scores, long long n, int K, int D) {
  extern __shared__ float s_diff[];                       // [warps][kMapRows][D]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* dbuf = s_diff + static_cast<size_t>(wib) * kMapRows * D;
  const long long groups = (n + kMapRows - 1) / kMapRows;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const int nq = (D + 31) / 32;
  for (long long grp = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + wib; grp < groups; grp += warps) {
    const long long r0 = grp * kMapRows;
    double bs[kMapRows];
    int bk[kMapRows];
#pragma unroll
    for (int r = 0; r < kMapRows; ++r) { bs[r] = -INFINITY; bk[r] = -1; }
    for (int k = 0; k < K; ++k) {
      float dreg[kMapRows][kMapMaxQ];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < kMapRows; ++r)
#pragma unroll
        for (int q = 0; q < kMapMaxQ; ++q) {
          const int j = lane + 32 * q;
          float d = 0.f;
          if (q < nq && j < D && r0 + r < n) d = Z[(r0 + r) * D + j] - mean[static_cast<size_t>(k) * D + j];
          dreg[r][q] = d;
          if (q < nq && j < D) dbuf[r * D + j] = d;
        }
      __syncwarp();
      float acc[kMapRows][kMapMaxQ];
#pragma unroll
      for (int r = 0; r < kMapRows; ++r)
#pragma unroll
        for (int q = 0; q < kMapMaxQ; ++q) acc[r][q] = 0.f;
      const float* Pk = prec + static_cast<size_t>(k) * D * D;
      for (int i = 0; i < D; ++i) {
        float p[kMapMaxQ];
#pragma unroll
        for (int q = 0; q < kMapMaxQ; ++q) {
          const int j = lane + 32 * q;
          p[q] = (q < nq && j < D) ? Pk[static_cast<size_t>(i) * D + j] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) {
          const float di = dbuf[r * D + i];
#pragma unroll
          for (int q = 0; q < kMapMaxQ; ++q) acc[r][q] = fmaf(p[q], di, acc[r][q]);
        }
      }
#pragma unroll
      for (int r = 0; r < kMapRows; ++r) {
        float quad = 0.f;
#pragma unroll
        for (int q = 0; q < kMapMaxQ; ++q) quad = fmaf(acc[r][q], dreg[r][q], quad);
        for (int o = 16; o > 0; o >>= 1) quad += __shfl_xor_sync(0xffffffffu, quad, o);
        const double s = -0.5 * (static_cast<double>(quad) + a_const[k]) + log_prior[k];
        if (lane == 0 && scores != nullptr && r0 + r < n) scores[(r0 + r) * K + k] = s;
        if (s > bs[r]) { bs[r] = s; bk[r] = k; }           // strict '>': the first maximum in species order wins
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < kMapRows; ++r)
        if (r0 + r < n) {
          int p = bk[r];
          if (p >= 0 && use_tau && bs[r] < tau) p = -1;     // `best_score < tau` -> NO_DETECT
          pred[r0 + r] = p;
          best[r0 + r] = bs[r];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// second moments of centred latents, float64:  out[i][j] += sum_r (z_ri - mu_{c(r),i}) (z_rj - mu_{c(r),j})
// over the rows of class k_sel, or of every labelled row (k_sel < 0, LDA pooling: each row centred by its class mean)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cov_accumulate_kernel(const float* __restrict__ Z, const int32_t* __restrict__ label,
                                                             const float* __restrict__ mean, int k_sel,
                                                             double* __restrict__ out, long long n, int K, int D,
                                                             long long rows_per_block) {
  __shared__ float s_a[64][17], s_b[64][17];
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int i0 = blockIdx.x * 16, j0 = blockIdx.y * 16;
  const long long r_begin = blockIdx.z * rows_per_block;
  const long long r_end = r_begin + rows_per_block < n ? r_begin + rows_per_block : n;
  double acc = 0.0;
  for (long long rb = r_begin; rb < r_end; rb += 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int rr = e >> 4, cc = e & 15;
      const long long r = rb + rr;
      float a = 0.f, b = 0.f;
      if (r < r_end) {
        const int lb = label[r];
        if (lb >= 0 && lb < K && (k_sel < 0 || lb == k_sel)) {
          if (i0 + cc < D) a = Z[r * D + i0 + cc] - mean[static_cast<size_t>(lb) * D + i0 + cc];
          if (j0 + cc < D) b = Z[r * D + j0 + cc] - mean[static_cast<size_t>(lb) * D + j0 + cc];
        }
      }
      s_a[rr][cc] = a;
      s_b[rr][cc] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) acc += static_cast<double>(s_a[rr][ti]) * static_cast<double>(s_b[rr][tj]);
  }
  if (i0 + ti < D && j0 + tj < D && acc != 0.0) atomicAdd(&out[static_cast<size_t>(i0 + ti) * D + j0 + tj], acc);
}

}  // namespace avld

using namespace avld;

extern "C" int avld_map_score(avld_ctx* c, const float* Z, const float* mean, const float* precision,
                              const double* a_const, const double* log_prior, double tau, int use_tau, int32_t* pred,
                              double* best, double* scores, int64_t n, int32_t K, int32_t D, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(Z && mean && precision && a_const && log_prior && pred && best, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(K >= 1 && D >= 1 && D <= 32 * kMapMaxQ, AVLD_ERR_UNSUPPORTED, "need 1 <= D <= %d", 32 * kMapMaxQ);
  if (n <= 0) return AVLD_OK;
  const size_t smem = static_cast<size_t>(4) * kMapRows * D * sizeof(float);
  const long long groups = (n + kMapRows - 1) / kMapRows;
  const int grid = static_cast<int>(std::min<long long>((groups + 3) / 4, static_cast<long long>(c->sm_count) * 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { LaunchScope ls(c, ST_MAP, st); map_score_kernel<<<grid, 128, smem, st>>>(Z, mean, precision, a_const, log_prior, tau, use_tau, pred, best, scores, n, K, D); }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

extern "C" int avld_cov_accumulate(avld_ctx* c, const float* Z, const int32_t* label, const float* mean, int32_t k_sel,
                                   double* out, int64_t n, int32_t K, int32_t D, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(Z && label && mean && out, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(K >= 1 && D >= 1 && k_sel < K, AVLD_ERR_INVALID, "bad K / D / class");
  if (n <= 0) return AVLD_OK;
  const int tiles = (D + 15) / 16;
  int splits = std::max(1, (c->sm_count * 4) / (tiles * tiles));
  const long long rows_per_block = std::max<long long>(64, ((n + splits - 1) / splits + 63) / 64 * 64);
  splits = static_cast<int>((n + rows_per_block - 1) / rows_per_block);
  dim3 grid(tiles, tiles, splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { LaunchScope ls(c, ST_MAP, st); cov_accumulate_kernel<<<grid, 256, 0, st>>>(Z, label, mean, k_sel, out, n, K, D, rows_per_block); }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}
