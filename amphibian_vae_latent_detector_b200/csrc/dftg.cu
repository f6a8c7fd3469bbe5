// dftg.cu -- folded STFT GEMM whose A operand never exists in global memory.  EXPERIMENTAL, opt-in (AVLD_DFT_GEN=1):
// correct (same parity tests as the default path) but slower in this form, see the measurements at the end of this comment.
//
// dftf3.cu reads the folded, windowed, hi/lo-split frames that fold3_kernel wrote: 3.08 MB per chunk written (that kernel
// is HBM-write bound) and ~1.5x that read back through L2 (the GEMM is L2->SM-feed bound).  Here each CTA keeps the
// normalised, PCM_16-quantised samples of its 128 frames in shared memory as int16 (a 99 KB span: the frames overlap
// 5.3x, so 50 816 samples cover 128 frames x 2048 taps) and eight worker warps build every 128 x 64 A tile (fp16 hi and
// lo, 128-byte swizzled, exactly what TMA would have written) straight into the pipeline stage; only the DFT matrix B
// still arrives by TMA.  Per chunk the kernel reads 0.6 MB of audio instead of ~7.7 MB of operand traffic.
//
// Same mathematics, tiling and epilogue as dftf3.cu (see fold2.cu for the three-level fold); differences:
//   * an M tile never crosses a chunk (tiles of 128 frames inside a chunk, the last one partly empty), so one span and
//     one power-of-two scale serve the whole tile; a CTA pair multiplies two independent tiles;
//   * two pipeline stages (A 32 KB built here + B 20 KB by TMA) next to the span; full barrier = B bytes + one arrival
//     per worker warp of both CTAs (fence.proxy.async before the arrive: generic-proxy stores, tensor-core reads);
//   * the worker warps are also the epilogue warps: two stages into the next item they read the finished Re and Im and
//     accumulate mel power (a stage can only have been refilled after the MMAs that complete the awaited accumulators
//     retired, so the wait is free; the tensor pipe idles meanwhile, but the workers set the pace of this kernel).
// Requires quantize_pcm16 (the span holds integers); other calls use fold3_kernel + dftf3_kernel.
//
// Measured (B200, 1024 chunks per launch; fold3_kernel + dftf3_kernel = 0.94 + 1.42 = 2.36 ms):
//   first version (8 workers, I2F per sample, Re held in registers: spills)      6.9 - 8.8 ms
//   no Re hold (workers set the pace, not the tensor pipe), no spills             5.3 ms
//   16 worker warps (2 tasks per thread and stage), two-phase even-class compute  4.2 ms
//   vectorised span fill (the scalar loop exposed every global load's latency)    3.4 ms
//   shared-memory address space kept for the workers' loads / stores (LDS, not LD) 3.1 ms
//   (at 3.4 ms:) AVLD_DBG=1 (no epilogue math) 2.95 | =8 (no tile building) 1.89 | =9 (neither) 1.28 ms
// i.e. tile building (~1.5 ms), epilogue math (~0.5 ms, serial on the same warps) and the MMA / span / barrier floor
// (~1.3 ms) add up instead of overlapping: the workers issue ~1.5 instructions per cycle and share the shared-memory pipe
// with the tensor core's operand reads.  What would make it pay: dedicated epilogue warps (needs setmaxnreg to fit the
// registers), fewer instructions per sample (the unpack of 16-bit samples is half of them), a third stage.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avld {

struct DftgItem {
  int kbp;      // 64-tap K blocks per part
  int cls;      // 0 = odd bins, 1 = bins 0 mod 4, 2 = bins 2 mod 4
  int edge_im;  // the class's self-paired tap belongs to the sin part
};

struct DftgParams {
  int num_pairs, num_tiles, tiles_per_chunk, num_items;
  DftgItem item[8];
  uint32_t idesc;
  const float* x;            // [n][L] raw chunks, or
  const int16_t* x16;        // [n][L] PCM_16
  const float4* chunk_par;   // per chunk (scale, 2^s, scaled?, -) from prep_kernel
  const float* inv2;         // per chunk power un-scale (NaN for a non-finite chunk)
  const float* win;          // [N/2 + 1]
  const MelTap* taps;        // [num_items * 160]
  float* melpow;             // [classes][rows][n_mels]
  long long plane_stride;
  int F, L, hop, n_fft, n_mels;
  int dbg;
};

namespace {
constexpr int kBM = 128, kBN = 160, kBK = 64, kImCol = 256;
constexpr int kSwz = 128;
constexpr int kABytes = kBM * kSwz;              // 16 KB (one of hi / lo)
constexpr int kBBytes = (kBN / 2) * kSwz;        // 10 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
constexpr int kStages = 2;
constexpr int kWorkers = 16;                     // worker warps (A-tile producers)
constexpr int kEpiWarps = 8;                     // the first 8 of them also run the epilogue (2 per TMEM lane quarter)
constexpr int kThreads = 64 + 32 * kWorkers;
constexpr int kWarpCols = kBN / 2, kGroups = kWarpCols / 16;
constexpr int kSpanMax = 127 * 384 + 2048 + 8;   // samples (hop 384, n_fft 2048)
constexpr int kSpanBytes = ((kSpanMax * 2 + 127) / 128) * 128;
constexpr int kWinBytes = ((1025 * 4 + 127) / 128) * 128;
constexpr int kTapBytes = 8 * kBN * 16 / 2;      // 4 items x 160 x 16 B = 10 240 B
constexpr int kSmemBytes = kStages * kStageBytes + kSpanBytes + 2 * kWinBytes + kTapBytes + 512 + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget");

// rint(v * 32767) clamped, as sf.write PCM_16 does (sample.cuh::finish_sample, returning the integer)
__device__ __forceinline__ int quantise(float v, float scale, int scaled) {
  if (scaled) {
    v = __fmul_rn(v, scale);
    v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
  }
  int q = __float2int_rn(__fmul_rn(v, 32767.0f));
  return q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
}

// The span holds u = q + 32768 (uint16).  A 16-bit half dropped into the mantissa of 2^23 is the float 2^23 + u, so a
// sample costs one PRMT and no conversion-pipe instruction (I2F runs at a quarter of the FMA rate and was the ceiling of
// the first version of this kernel); differences need nothing more (the offsets cancel), sums subtract the offsets once.
constexpr float kTwo23 = 8388608.0f;
constexpr float kSumBias = 2.0f * 8388608.0f + 65536.0f;       // two mantissa offsets + two sign biases

// f[j] = 2^23 + u[i + j], j = 0..7 (i 16-byte aligned)
__device__ __forceinline__ void lds8(const uint16_t* s, int i, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(s + i);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7610));
    f[2 * j + 1] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7632));
  }
}
// f[q] = 2^23 + u[i - q], q = 0..7 (i - 8 is 16-byte aligned): the aligned vector below i, reversed, plus u[i]
__device__ __forceinline__ void lds8_down(const uint16_t* s, int i, float (&f)[8]) {
  float t[8];
  lds8(s, i - 8, t);
  f[0] = __uint_as_float(0x4B000000u | static_cast<uint32_t>(s[i]));
#pragma unroll
  for (int q = 1; q < 8; ++q) f[q] = t[8 - q];
}
// signed sums / differences of two biased samples given as 2^23 + u
__device__ __forceinline__ float ssum(float a, float b) { return (a - kSumBias) + b; }    // exact: |.| < 2^24 at each step
__device__ __forceinline__ float sdif(float a, float b) { return a - b; }
__device__ __forceinline__ float sval(float a) { return a - (kTwo23 + 32768.0f); }

__device__ __forceinline__ void load8f(const float* t, int i, float (&w)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(t + i), b = *reinterpret_cast<const float4*>(t + i + 4);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// hi = v rounded to 11 significant bits by integer arithmetic on the bit pattern (magnitude round-half-up; any rounding
// works as long as hi is fp16-representable and lo = v - hi is exact), two packed cvt per pair of values
__device__ __forceinline__ void split_sts(uint8_t* a_hi, uint8_t* a_lo, int row, int chunk, const float (&v)[8]) {
  __align__(16) __half2 h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float a = v[2 * q], b = v[2 * q + 1];
    const float ha = __int_as_float((__float_as_int(a) + 0x1000) & 0xFFFFE000);
    const float hb = __int_as_float((__float_as_int(b) + 0x1000) & 0xFFFFE000);
    h[q] = __floats2half2_rn(ha, hb);
    l[q] = __floats2half2_rn(a - ha, b - hb);
  }
  const int off = row * kSwz + ((chunk ^ (row & 7)) << 4);       // 128-byte swizzle: 16-byte chunk index xor (row % 8)
  *reinterpret_cast<uint4*>(a_hi + off) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(a_lo + off) = *reinterpret_cast<const uint4*>(l);
}
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dftg_kernel(const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const DftgParams P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (not through an integer cast): the worker warps access
  // this memory with ordinary loads / stores, and only then does the compiler know they are LDS / STS rather than generic
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint16_t* s_span = reinterpret_cast<uint16_t*>(smem + kStages * kStageBytes);   // q + 32768
  float* s_win = reinterpret_cast<float*>(smem + kStages * kStageBytes + kSpanBytes);             // w[k]
  float* s_wrv = reinterpret_cast<float*>(smem + kStages * kStageBytes + kSpanBytes + kWinBytes); // w[N/2 - k]
  MelTap* s_taps = reinterpret_cast<MelTap*>(smem + kStages * kStageBytes + kSpanBytes + 2 * kWinBytes);
  uint8_t* tail = smem + kStages * kStageBytes + kSpanBytes + 2 * kWinBytes + kTapBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);     // [2]  (used in the leader)
  uint64_t* empty_bar = full_bar + 4;                         // [2]  (per CTA)
  uint64_t* tmem_full = empty_bar + 4;                        // [2]  Re / Im complete (per CTA)
  uint64_t* tmem_empty = tmem_full + 2;                       // [2]  Re / Im drained (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_clusters = static_cast<int>(ncluster_id_x());
  const int cluster = static_cast<int>(cluster_id_x());
  const int N = P.n_fft, H = N >> 1, Q = N >> 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2 + 2 * kWorkers);   // B: leader's expect_tx arrive + the peer's arrive; A: every worker warp
      mbar_init(&empty_bar[s], 1);                 // one multicast commit
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&tmem_full[h], 1);
      mbar_init(&tmem_empty[h], 2 * kEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  for (int i = threadIdx.x; i < P.num_items * kBN; i += blockDim.x) s_taps[i] = P.taps[i];
  for (int i = threadIdx.x; i <= H; i += blockDim.x) {
    s_win[i] = P.win[i];
    s_wrv[i] = P.win[H - i];
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer of B (both CTAs, warp-uniform)
    int stage = 0;
    uint32_t phase = 0;
    for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
      for (int it = 0; it < P.num_items; ++it) {
        const int kbp = P.item[it].kbp, nkb = 2 * kbp;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 100 + stage);
          uint8_t* sb_hi = smem + stage * kStageBytes + 2 * kABytes;
          uint8_t* sb_lo = sb_hi + kBBytes;
          const int part = kb < kbp ? 0 : 1;
          const int bx = (kb - part * kbp) * kBK;
          const int by = (it * 2 + part) * kBN + static_cast<int>(rank) * (kBN / 2);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * 2 * kBBytes);      // hi + lo, both CTAs
            else mbar_arrive_cluster(&full_bar[stage], 0);
            tma_load_2d_pair(sb_hi, &tmB_hi, &full_bar[stage], bx, by);
            tma_load_2d_pair(sb_lo, &tmB_lo, &full_bar[stage], bx, by);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, warp-uniform)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
        for (int it = 0; it < P.num_items; ++it) {
          const int kbp = P.item[it].kbp, nkb = 2 * kbp;
          for (int kb = 0; kb < nkb; ++kb) {
            if (kb == 0 || kb == kbp) {
              mbar_wait(&tmem_empty[kb == 0 ? 0 : 1], acc_phase ^ 1u, 200 + (kb == 0 ? 0 : 1));
              tcgen05_fence_after();
            }
            mbar_wait(&full_bar[stage], phase, 300 + stage);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (kb < kbp ? 0u : static_cast<uint32_t>(kImCol));
            const int kb_acc = kb < kbp ? kb : kb - kbp;
            const uint32_t a_hi = smem_u32(smem + stage * kStageBytes);
            const uint32_t a_lo = a_hi + kABytes, b_hi = a_lo + kABytes, b_lo = b_hi + kBBytes;
            const uint64_t da_hi = make_smem_desc(a_hi, kSwz), da_lo = make_smem_desc(a_lo, kSwz);
            const uint64_t db_hi = make_smem_desc(b_hi, kSwz), db_lo = make_smem_desc(b_lo, kSwz);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k * 2);
                umma_f16_pair(d_tmem, da_hi + koff, db_hi + koff, P.idesc, (kb_acc | k) != 0 ? 1u : 0u);
                umma_f16_pair(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                umma_f16_pair(d_tmem, da_hi + koff, db_lo + koff, P.idesc, 1u);
              }
              umma_commit_pair(&empty_bar[stage], 0x3);
              if (kb == kbp - 1) umma_commit_pair(&tmem_full[0], 0x3);
              if (kb == nkb - 1) umma_commit_pair(&tmem_full[1], 0x3);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ workers: build A tiles, drain Re, finish Im
    const int wtid = threadIdx.x - 64;                       // 0 .. 255
    const bool epi_warp = warp < 2 + kEpiWarps;              // warps 2..9 own the accumulator read-out
    const int quarter = warp & 3, sub = ((warp - 2) >> 2) & 1; // TMEM lane quarter of this warp, half of the item's columns
    const int row_e = quarter * 32 + lane;                   // the frame this thread owns in the epilogue
    const int b0 = sub * kWarpCols;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool skip = (P.dbg & 1) != 0;
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;

    // epilogue state of the item whose accumulators are (still) in flight
    struct Epi {
      bool pending;            // Im of this item not yet processed
      bool valid;
      float s2, e_re, e_im;
      float* mrow;
      const MelTap* taps;
    } epi{false, false, 0.f, 0.f, 0.f, nullptr, nullptr};
    // Re and Im are read together, 16 columns at a time, once the whole item is complete.  Holding Re in registers from
    // the end of the cos part on (as dftf3.cu does to keep its issuer fed) is pointless here -- the workers, not the
    // tensor pipe, set the pace -- and 80 live registers across the producer loop meant spills.
    auto finish_item = [&]() {
      int mcur = epi.taps[0].first;
      float a0 = 0.f, a1 = 0.f;
      mbar_wait(&tmem_full[0], acc_phase, 400);
      mbar_wait(&tmem_full[1], acc_phase, 401);
      tcgen05_fence_after();
#pragma unroll 1
      for (int q = 0; q < kGroups; ++q) {
        uint32_t re[16], im[16];
        tmem_ld16(t_acc + b0 + q * 16, re);
        tmem_ld16(t_acc + kImCol + b0 + q * 16, im);
        tmem_ld_wait();
        if (q == kGroups - 1) {                              // last read of this warp: both column ranges may be overwritten
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) {
              mbar_arrive(&tmem_empty[0]);
              mbar_arrive(&tmem_empty[1]);
            } else {
              mbar_arrive_cluster(&tmem_empty[0], 0);
              mbar_arrive_cluster(&tmem_empty[1], 0);
            }
          }
        }
        if (!skip) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const MelTap tp = epi.taps[q * 16 + j];
            const float coef = __int_as_float(tp.pad);
            const float a = fmaf(epi.e_re, coef, __uint_as_float(re[j]));
            const float b = fmaf(epi.e_im, coef, __uint_as_float(im[j]));
            const float pw = (a * a + b * b) * epi.s2;
            if (mcur < tp.first) {
#pragma unroll 1
              while (mcur < tp.first) {
                if (epi.valid && a0 != 0.f) atomicAdd(epi.mrow + mcur, a0);
                a0 = a1;
                a1 = 0.f;
                ++mcur;
              }
            }
            a0 = fmaf(tp.w0, pw, a0);
            a1 = fmaf(tp.w1, pw, a1);
          }
        }
      }
      if (epi.valid && a0 != 0.f && mcur < P.n_mels) atomicAdd(epi.mrow + mcur, a0);
      if (epi.valid && a1 != 0.f && mcur + 1 < P.n_mels) atomicAdd(epi.mrow + mcur + 1, a1);
      acc_phase ^= 1u;
      epi.pending = false;
    };

    for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
      // ---- this CTA's tile: 128 frames of one chunk
      const int tile = pair * 2 + static_cast<int>(rank);
      const bool tile_ok = tile < P.num_tiles;
      const int chunk = tile_ok ? tile / P.tiles_per_chunk : 0;
      const int f0 = tile_ok ? (tile - chunk * P.tiles_per_chunk) * kBM : 0;
      const int rows_valid = tile_ok ? min(kBM, P.F - f0) : 0;
      const float4 par = P.chunk_par[chunk];
      const float cmul = par.y * (1.0f / 32768.0f);
      // ---- span: normalised + quantised samples of padded indices [f0 * hop, f0 * hop + 127 * hop + N)
      {
        named_barrier_sync(1, 32 * kWorkers);                // every worker is done reading the previous span
        const float scale = par.x;
        const int scaled = par.z != 0.f;
        const int p0 = f0 * P.hop, plen = P.L + N;
        const int span = (kBM - 1) * P.hop + N;
        const float* xf = P.x ? P.x + static_cast<size_t>(chunk) * P.L : nullptr;
        const int16_t* xi = P.x16 ? P.x16 + static_cast<size_t>(chunk) * P.L : nullptr;
        // 8 samples per thread and iteration: vector loads where the run is interior (no reflection, inside the chunk),
        // independent scalar loads otherwise; a sample-at-a-time loop left every iteration exposed to the full global
        // latency and was a quarter of the kernel's time
        const bool vec_ok = (P.L % 8 == 0) && (P.hop % 8 == 0) && (reinterpret_cast<uintptr_t>(P.x) % 16 == 0) &&
                            (reinterpret_cast<uintptr_t>(P.x16) % 16 == 0);
#pragma unroll 2
        for (int i = wtid * 8; i < span + 8; i += 8 * 32 * kWorkers) {
          const int p = p0 + i, sb = p - H;
          float v[8];
          uint32_t u16[8];
          if (tile_ok && vec_ok && i + 8 <= span && sb >= 0 && sb + 8 <= P.L) {
            if (xf) {
              const float4 u0 = *reinterpret_cast<const float4*>(xf + sb), u1 = *reinterpret_cast<const float4*>(xf + sb + 4);
              v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w; v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
            } else {
              const uint4 u = *reinterpret_cast<const uint4*>(xi + sb);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[2 * j] = static_cast<float>(static_cast<int>(w[j] << 16) >> 16) * (1.0f / 32768.0f);
                v[2 * j + 1] = static_cast<float>(static_cast<int>(w[j]) >> 16) * (1.0f / 32768.0f);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) u16[j] = static_cast<uint32_t>(quantise(v[j], scale, scaled) + 32768);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              int q = 0;
              if (tile_ok && i + j < span && p + j < plen) {
                int src = p + j - H;                         // np.pad(y, n_fft // 2, mode="reflect")
                if (src < 0) src = -src;
                if (src >= P.L) src = 2 * (P.L - 1) - src;
                const float r = xf ? xf[src] : static_cast<float>(xi[src]) * (1.0f / 32768.0f);
                q = quantise(r, scale, scaled);
              }
              u16[j] = static_cast<uint32_t>(q + 32768);
            }
          }
          uint4 o;
          o.x = u16[0] | (u16[1] << 16);
          o.y = u16[2] | (u16[3] << 16);
          o.z = u16[4] | (u16[5] << 16);
          o.w = u16[6] | (u16[7] << 16);
          *reinterpret_cast<uint4*>(s_span + i) = o;
        }
        named_barrier_sync(1, 32 * kWorkers);
      }
      // ---- per-row epilogue constants of this tile
      const long long g = static_cast<long long>(chunk) * P.F + f0 + row_e;
      const bool valid = tile_ok && row_e < rows_valid;
      const float s2 = valid ? P.inv2[chunk] : 0.f;
      float edge0, edge1, edge2;
      {
        const uint16_t* sr = s_span + row_e * P.hop;
        const int E = N >> 3;
        auto sv = [&](int tap) { return static_cast<float>(static_cast<int>(sr[tap]) - 32768) * cmul; };
        const float wq = s_win[Q], we = s_win[E], w3 = s_win[Q + E];
        const float xe = sv(E), x7e = sv(N - E), x3e = sv(Q + E), x5e = sv(H + E);
        edge0 = wq * (sv(Q) - sv(H + Q));                            // O[N/4]
        edge1 = we * (xe + x7e) + w3 * (x3e + x5e);                // P[N/8]
        edge2 = we * (xe - x7e) - w3 * (x3e - x5e);                // R[N/8]
      }

      for (int it = 0; it < P.num_items; ++it) {
        const int kbp = P.item[it].kbp, nkb = 2 * kbp, cls = P.item[it].cls;
        for (int kb = 0; kb < nkb; ++kb) {
          // accumulator hand-over, placed where the awaited MMAs have certainly retired (see header)
          // Stage 1 of an item reuses the buffer of the previous item's last stage: once it is free, that item's
          // accumulators are complete.  The epilogue warps then read them out and sit this stage out (they arrive on its
          // barrier first: they write nothing into it), the other eight warps build the whole stage meanwhile.  (Not at
          // stage 0 or 2: the MMAs of stage 0 wait for this very read-out, so nothing that waits for them may precede it.)
          const bool hand_over = kb == 1 && epi.pending;
          if (hand_over) {
            if (epi_warp) {
              mbar_wait(&empty_bar[stage], phase ^ 1u, 510 + stage);     // the barrier's previous phase is certainly over
              if (lane == 0) {
                if (leader) mbar_arrive(&full_bar[stage]);
                else mbar_arrive_cluster(&full_bar[stage], 0);
              }
              finish_item();
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
              continue;
            }
            acc_phase ^= 1u;
            epi.pending = false;
          }
          const int t_first = hand_over ? wtid - 32 * kEpiWarps : wtid;          // warps 10..17 only during a hand-over
          const int t_step = hand_over ? 32 * (kWorkers - kEpiWarps) : 32 * kWorkers;
          // ---- build A stage: (class, part, 64 taps) for 128 rows
          mbar_wait(&empty_bar[stage], phase ^ 1u, 500 + stage);
          uint8_t* a_hi = smem + stage * kStageBytes;
          uint8_t* a_lo = a_hi + kABytes;
          const int part = kb < kbp ? 0 : 1;
          const int kblk = (kb - part * kbp) * kBK;
#pragma unroll 1
          for (int t = (P.dbg & 8) ? kBM * 8 : t_first; t < kBM * 8; t += t_step) {      // dbg 8: barrier protocol only
            const int row = t >> 3, ch = t & 7;
            const int k0 = kblk + ch * 8;
            const uint16_t* sr = s_span + row * P.hop;
            float out[8];
            float a1[8], a2[8], a3[8], a4[8], wk[8], wh[8];
            lds8(sr, k0, a1);
            lds8_down(sr, N - k0, a2);
            lds8_down(sr, H - k0, a3);
            lds8(sr, H + k0, a4);
            load8f(s_win, k0, wk);                           // w[k]
            load8f(s_wrv, k0, wh);                           // w[N/2 - k]
            if (k0 == 0) {                                   // u[0] and u[N/2] pair with nothing: a "zero" partner sample
              a2[0] = kTwo23 + 32768.0f;
              a4[0] = kTwo23 + 32768.0f;
            }
            if (cls == 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float v = part == 0 ? wk[q] * ssum(a1[q], a2[q]) - wh[q] * ssum(a3[q], a4[q])
                                          : wk[q] * sdif(a1[q], a2[q]) + wh[q] * sdif(a3[q], a4[q]);
                out[q] = v * cmul;
              }
              if (part == 1 && k0 == 0) out[0] = 0.f;
            } else {
              // first half of the fold (taps k, N-k, N/2-k, N/2+k), then the mirrored half; live registers stay low
              float first[8];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                first[q] = part == 0 ? wk[q] * ssum(a1[q], a2[q]) + wh[q] * ssum(a3[q], a4[q])
                                     : wk[q] * sdif(a1[q], a2[q]) - wh[q] * sdif(a3[q], a4[q]);
              lds8_down(sr, Q - k0, a1);                     // a5
              lds8(sr, H + Q + k0, a2);                      // a6
              lds8(sr, Q + k0, a3);                          // a7
              lds8_down(sr, H + Q - k0, a4);                 // a8
              load8f(s_wrv, Q + k0, wk);                     // w[N/4 - k] = w[N/2 - (N/4 + k)]
              load8f(s_win, Q + k0, wh);                     // w[N/4 + k]
              if (k0 == 0) {                                 // u[N/4] and u[3N/4] are one pair, not two
                a3[0] = kTwo23 + 32768.0f;
                a4[0] = kTwo23 + 32768.0f;
              }
              const float sgn = (cls == 1) == (part == 0) ? 1.f : -1.f;     // cos: P + s P', sin: R - s R' (s = +1 for 0 mod 4)
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float second = part == 0 ? wk[q] * ssum(a1[q], a2[q]) + wh[q] * ssum(a3[q], a4[q])
                                               : wk[q] * sdif(a1[q], a2[q]) - wh[q] * sdif(a3[q], a4[q]);
                out[q] = (first[q] + sgn * second) * cmul;
              }
              if (part == 1 && k0 == 0) out[0] = 0.f;
            }
            split_sts(a_hi, a_lo, row, ch, out);
          }
          fence_proxy_async();                               // generic-proxy stores -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&full_bar[stage]);
            else mbar_arrive_cluster(&full_bar[stage], 0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // the item just queued is finished one stage into the next one
        epi.pending = true;
        epi.valid = valid;
        epi.s2 = s2;
        const float e_cls = cls == 0 ? edge0 : (cls == 1 ? edge1 : edge2);
        epi.e_re = P.item[it].edge_im ? 0.f : e_cls;
        epi.e_im = P.item[it].edge_im ? e_cls : 0.f;
        epi.mrow = P.melpow + cls * P.plane_stride + g * P.n_mels;
        epi.taps = s_taps + it * kBN + b0;
      }
    }
    if (epi.pending && epi_warp) finish_item();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
#endif
}

bool dftg_supported(const avld_ctx* c) {
  if (!c->dft_fold2 || c->f2_levels != 3) return false;
  if (c->p.n_fft != 2048 || c->p.hop != 384) return false;     // the span buffer is sized for this geometry
  if (c->f2_items > 8) return false;
  for (int it = 0; it < c->f2_items; ++it)
    if (c->f2_item[it].kbp < 1) return false;
  const char* e = getenv("AVLD_DFT_GEN");                       // opt-in: see the header of this file
  return e != nullptr && atoi(e) != 0;
}

int launch_stft_mel_gen(avld_ctx* c, const float* x, const int16_t* x16, int n, cudaStream_t st) {
  DftgParams P{};
  P.tiles_per_chunk = (c->F + kBM - 1) / kBM;
  P.num_tiles = n * P.tiles_per_chunk;
  P.num_pairs = (P.num_tiles + 1) / 2;
  P.num_items = c->f2_items;
  for (int it = 0; it < c->f2_items; ++it) P.item[it] = {c->f2_item[it].kbp, c->f2_item[it].cls, c->f2_item[it].edge_im};
  P.idesc = avld_make_idesc(0, 0, 256, kBN);
  P.x = x;
  P.x16 = x16;
  P.chunk_par = c->d_chunk_par;
  P.inv2 = c->d_inv2;
  P.win = c->d_win;
  P.taps = c->d_taps3;
  P.melpow = c->d_melpow;
  P.plane_stride = c->melpow_plane;
  P.F = c->F;
  P.L = c->L;
  P.hop = c->p.hop;
  P.n_fft = c->p.n_fft;
  P.n_mels = c->M;
  {
    const char* d = getenv("AVLD_DBG");
    P.dbg = d ? atoi(d) : 0;
  }
  static bool configured = false;
  if (!configured) {
    AVLD_CUDA(cudaFuncSetAttribute(dftg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  AVLD_CHECK(static_cast<size_t>(P.num_items) * kBN * sizeof(MelTap) <= kTapBytes, AVLD_ERR_UNSUPPORTED, "too many FFT bins");
  if (c->planes_dirty)
    AVLD_CUDA(cudaMemsetAsync(c->d_melpow, 0, static_cast<size_t>(c->melpow_plane) * c->f2_classes * sizeof(float), st));
  c->planes_dirty = true;
  const int grid = 2 * std::min(P.num_pairs, c->sm_count / 2);
  if (grid < 2) return AVLD_OK;
  LaunchScope ls(c, ST_STFT_MEL, st);
  dftg_kernel<<<grid, kThreads, kSmemBytes, st>>>(c->tm_B3_hi, c->tm_B3_lo, P);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld
