// radial.cu -- F1-F3 and D2: per-species centroid sums, radii to K centroids, exact order statistics
// for the q_in / q_out quantiles, and the accept / priority decision.  All streaming, HBM-bound.
//   08_fit_radial_detector.py:105-106 (l2_norm_rows), :310-333 (fit_species_with_fp_control)
//   09_evaluate_wav_detection.py:354-355 (l2), :416-436 (accept set + PRIORITY_ORDER)
//   10_benchmark_folder_detection.py:175-199 (best_distance)
#include <algorithm>
#include <map>
#include <tuple>

#include "common.cuh"

namespace avld {

// ------------------------------------------------------------------------------------------------
// centroid sums: float64 accumulation (rank-count independent after the all-reduce)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) centroid_kernel(const float* __restrict__ Z, const int32_t* __restrict__ label,
                                                       double* __restrict__ sum, long long* __restrict__ cnt,
                                                       long long n, int K, int D, long long rows_per_block) {
  extern __shared__ double s_sum[];               // [K][D]
  __shared__ long long s_cnt[64];
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) s_sum[i] = 0.0;
  if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {     // thread owns column d: no smem conflicts
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {
      int lb[4];
      float zv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        lb[q] = label[r + q];
        zv[q] = Z[(r + q) * D + d];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (lb[q] >= 0 && lb[q] < K) s_sum[lb[q] * D + d] += static_cast<double>(zv[q]);
    }
    for (; r < r1; ++r) {
      const int lb = label[r];
      if (lb >= 0 && lb < K) s_sum[lb * D + d] += static_cast<double>(Z[r * D + d]);
    }
  }
  if (threadIdx.x == 0) {
    for (long long r = r0; r < r1; ++r) {
      const int lb = label[r];
      if (lb >= 0 && lb < K) s_cnt[lb] += 1;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * D; i += blockDim.x)
    if (s_sum[i] != 0.0) atomicAdd(&sum[i], s_sum[i]);
  if (threadIdx.x < K && s_cnt[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(&cnt[threadIdx.x]), static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// ------------------------------------------------------------------------------------------------
// radii: one warp per latent, centroids staged in shared memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) radii_kernel(const float* __restrict__ Z, const float* __restrict__ cent,
                                                    float* __restrict__ radii, long long n, int K, int D) {
  extern __shared__ float s_c[];                  // [K][D]
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) s_c[i] = cent[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const float* z = Z + row * D;
    for (int k = 0; k < K; ++k) {
      float acc = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float t = z[d] - s_c[k * D + d];
        acc = fmaf(t, t, acc);
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) radii[row * K + k] = sqrtf(acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// decision
// ------------------------------------------------------------------------------------------------
__global__ void decide_kernel(const float* __restrict__ radii, const double* __restrict__ thr,
                              const int32_t* __restrict__ prio, int32_t* __restrict__ pred, float* __restrict__ best,
                              long long n, int K) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float bd = INFINITY;
    int sel = -1, sel_rank = 0x7fffffff;
    for (int k = 0; k < K; ++k) {
      const double rk = thr[k];                                // NaN (e.g. a species fitted on no data): `d <= rk` is false,
      const float d = radii[i * K + k];                        // but d still counts for best_d, as in 10:177-187.  A species
      bd = fminf(bd, d);                                       // that is ABSENT from `thresholds` is dropped by the caller
                                                               // (09:418-419).  min(best_d, d): NaN d never lowers it
      if (static_cast<double>(d) <= rk && prio[k] < sel_rank) {  // accept iff d <= rk; first in priority order wins
        sel = k;
        sel_rank = prio[k];
      }
    }
    pred[i] = sel;
    best[i] = bd;
  }
}

// ------------------------------------------------------------------------------------------------
// exact order statistics by 3-round histogram selection on the float bit pattern (radii >= 0, so the
// IEEE bits are monotone as unsigned integers): 11 + 11 + 9 bits... rounds use shifts 20, 9, 0.
// A bucket = (species column, side, lo, shift): keys in [lo, lo + 2048 << shift) are counted into
// 2048 bins of width 1 << shift.  Shared-memory privatised histograms, <= 8 buckets per launch.
// ------------------------------------------------------------------------------------------------
struct Bucket {
  int32_t k, side;
  uint32_t lo, shift;
};
struct BucketGroup {
  Bucket b[8];
  int nb;
};

__global__ void __launch_bounds__(512) select_hist_kernel(const float* __restrict__ radii, const int32_t* __restrict__ label,
                                                          long long n, int K, const BucketGroup G, unsigned int* __restrict__ hist) {
  extern __shared__ unsigned int s_h[];           // [nb][2048]
  for (int i = threadIdx.x; i < G.nb * 2048; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int lb = label[i];
    if (lb < 0) continue;
    for (int b = 0; b < G.nb; ++b) {
      const Bucket bk = G.b[b];
      const int side = (lb == bk.k) ? 0 : 1;
      if (side != bk.side) continue;
      const uint32_t key = __float_as_uint(radii[i * K + bk.k]);
      if (key < bk.lo) continue;
      const uint32_t d = (key - bk.lo) >> bk.shift;
      if (d < 2048u) atomicAdd(&s_h[b * 2048 + d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G.nb * 2048; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

}  // namespace avld

using namespace avld;

extern "C" int avld_centroid_accumulate(avld_ctx* c, const float* Z, const int32_t* label, double* sum, int64_t* cnt,
                                        int64_t n, int32_t K, int32_t D, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(Z && label && sum && cnt, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(K >= 1 && K <= 64 && D >= 1 && static_cast<size_t>(K) * D * 8 <= 96 * 1024, AVLD_ERR_UNSUPPORTED,
             "K must be in [1,64] and K*D*8 <= 96 KB");
  if (n <= 0) return AVLD_OK;
  const size_t smem = static_cast<size_t>(K) * D * sizeof(double);
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(centroid_kernel), 96 * 1024));
  const long long rows_per_block = std::max<long long>(64, (n + c->sm_count * 4 - 1) / (c->sm_count * 4));
  const int grid = static_cast<int>((n + rows_per_block - 1) / rows_per_block);
  { LaunchScope ls(c, ST_CENTROID, static_cast<cudaStream_t>(stream)); centroid_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(Z, label, sum, reinterpret_cast<long long*>(cnt), n, K, D, rows_per_block); }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

extern "C" int avld_radii(avld_ctx* c, const float* Z, const float* centroid, float* radii, int64_t n, int32_t K,
                          int32_t D, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(Z && centroid && radii, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(K >= 1 && D >= 1 && static_cast<size_t>(K) * D * 4 <= 96 * 1024, AVLD_ERR_UNSUPPORTED, "K*D*4 must be <= 96 KB");
  if (n <= 0) return AVLD_OK;
  const size_t smem = static_cast<size_t>(K) * D * sizeof(float);
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(radii_kernel), 96 * 1024));
  const long long want = (n + 7) / 8;
  const int grid = static_cast<int>(std::min<long long>(want, static_cast<long long>(c->sm_count) * 8));
  { LaunchScope ls(c, ST_RADII, static_cast<cudaStream_t>(stream)); radii_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(Z, centroid, radii, n, K, D); }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

extern "C" int avld_decide(avld_ctx* c, const float* radii, const double* thr, const int32_t* priority_rank,
                           int32_t* pred, float* best_d, int64_t n, int32_t K, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(radii && thr && priority_rank && pred && best_d, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(K >= 1, AVLD_ERR_INVALID, "K must be >= 1");
  if (n <= 0) return AVLD_OK;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(c->sm_count) * 8));
  { LaunchScope ls(c, ST_DECIDE, static_cast<cudaStream_t>(stream)); decide_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(radii, thr, priority_rank, pred, best_d, n, K); }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

extern "C" int avld_order_stats(avld_ctx* c, const float* radii, const int32_t* label, int64_t n, int32_t K,
                                const avld_rank_query* queries, int32_t n_q, float* out, void* stream) {
  AVLD_ENTER(c);
  AVLD_CHECK(radii && label && queries && out, AVLD_ERR_INVALID, "NULL argument");   // an order statistic of
  AVLD_CHECK(n > 0 && K >= 1 && n_q >= 1 && n_q <= 4096, AVLD_ERR_INVALID, "bad n / K / n_q");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  struct QState { uint32_t lo; int64_t rank; };
  std::vector<QState> qs(n_q);
  for (int q = 0; q < n_q; ++q) {
    AVLD_CHECK(queries[q].species >= 0 && queries[q].species < K && (queries[q].side == 0 || queries[q].side == 1) && queries[q].rank >= 0,
               AVLD_ERR_INVALID, "query %d is malformed", q);
    qs[q] = {0u, queries[q].rank};
  }
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(select_hist_kernel), 8 * 2048 * 4));
  const uint32_t shifts[3] = {20u, 9u, 0u};
  std::vector<unsigned int> h_hist;
  for (int round = 0; round < 3; ++round) {
    const uint32_t shift = shifts[round];
    // distinct buckets of this round
    std::map<std::tuple<int, int, uint32_t>, int> index;
    std::vector<Bucket> buckets;
    std::vector<int> q2b(n_q);
    for (int q = 0; q < n_q; ++q) {
      auto key = std::make_tuple(queries[q].species, queries[q].side, qs[q].lo);
      auto it = index.find(key);
      if (it == index.end()) {
        it = index.emplace(key, static_cast<int>(buckets.size())).first;
        buckets.push_back({queries[q].species, queries[q].side, qs[q].lo, shift});
      }
      q2b[q] = it->second;
    }
    const size_t need = buckets.size() * 2048 * sizeof(unsigned int);
    if (need > c->hist_bytes) {
      if (c->d_hist) AVLD_CUDA(cudaFree(c->d_hist));
      c->d_hist = nullptr;
      AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_hist), need));
      c->hist_bytes = need;
    }
    AVLD_CUDA(cudaMemsetAsync(c->d_hist, 0, need, st));
    const int grid = static_cast<int>(std::min<long long>((n + 511) / 512, static_cast<long long>(c->sm_count) * 2));
    for (size_t g0 = 0; g0 < buckets.size(); g0 += 8) {
      BucketGroup G{};
      G.nb = static_cast<int>(std::min<size_t>(8, buckets.size() - g0));
      for (int b = 0; b < G.nb; ++b) G.b[b] = buckets[g0 + b];
      { LaunchScope ls(c, ST_SELECT, st); select_hist_kernel<<<grid, 512, static_cast<size_t>(G.nb) * 2048 * 4, st>>>(radii, label, n, K, G, c->d_hist + g0 * 2048); }
      AVLD_CUDA(cudaGetLastError());
    }
    h_hist.resize(buckets.size() * 2048);
    AVLD_CUDA(cudaMemcpyAsync(h_hist.data(), c->d_hist, need, cudaMemcpyDeviceToHost, st));
    AVLD_CUDA(cudaStreamSynchronize(st));
    for (int q = 0; q < n_q; ++q) {
      const unsigned int* h = h_hist.data() + static_cast<size_t>(q2b[q]) * 2048;
      int64_t r = qs[q].rank;
      int bin = -1;
      for (int b = 0; b < 2048; ++b) {
        if (r < static_cast<int64_t>(h[b])) { bin = b; break; }
        r -= h[b];
      }
      AVLD_CHECK(bin >= 0, AVLD_ERR_INVALID, "query %d: rank %lld is beyond the population of (species %d, side %d)", q,
                 static_cast<long long>(queries[q].rank), queries[q].species, queries[q].side);
      qs[q].lo += static_cast<uint32_t>(bin) << shift;
      qs[q].rank = r;
    }
  }
  for (int q = 0; q < n_q; ++q) {
    float f;
    memcpy(&f, &qs[q].lo, 4);
    out[q] = f;
  }
  return AVLD_OK;
}
