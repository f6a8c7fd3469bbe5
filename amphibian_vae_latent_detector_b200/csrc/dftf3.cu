// dftf3.cu -- M2: the three-times folded STFT as a GEMM on CTA pairs (tcgen05 cta_group::2), |X|^2 and the slaney mel
// filterbank in the epilogue (replaces librosa.feature.melspectrogram, map_detector_core.py:219-228).  Operands: folded
// frames A3 from fold3.cu (tile-major fp16 hi / lo), DFT matrix B3 (ctx.cu).
//
// The FFT bins with mel weight are split into three classes (odd | 0 mod 4 | 2 mod 4); each class has its own A columns
// (cos part | sin part, N/4 or N/8 taps each) and is covered by work items of 160 bins.  One item = 256 frames (a CTA pair)
// x 160 bins: K loop over the cos part (-> Re) and then the sin part (-> Im).  Split precision as everywhere: hi*hi + lo*hi
// + hi*lo, fp32 in TMEM.  Tensor memory: Re_A | Im | Re_B (160 columns each): the real parts alternate between two
// buffers, so the cos part of item i + 1 runs while the epilogue still reads item i; Im is single: the sin part of item
// i + 1 waits until the epilogue has pulled item i's last column (it has the whole cos part of item i + 1 to do so).
//   warp 0 (both CTAs)          TMA producer: own 128 frames of A (one 32 KB box = hi + lo tile, contiguous in memory), own
//                               half (80 rows) of the B tile
//   warp 1 (leader only)        MMA issuer: tcgen05.mma.cta_group::2 (M 256, N 160), commits multicast to both CTAs
//   warps 2..9    (both CTAs)   epilogue, once the item's Im is complete: Re and Im streamed from TMEM 16 columns at a time
//                               (the next group's loads in flight under the current group's math); edge
//                               term, |X|^2, sparse slaney mel accumulation (band changes are warp-uniform branches),
//                               un-scale, vector reductions (four mel bands per red.global.add.v4.f32) into the class's
//                               plane.
// Bit-reproducible: a warp covers 80 consecutive bins of one class (160 FFT bins), a mel filter spans <= 67 FFT bins, so
// every (frame, filter, class plane) cell receives at most two non-zero contributions (a + b is order independent; the
// zeros that pad a group of four bands do not change a sum).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace avld {

struct Dftf3Item {
  int a_kb0;    // first 64-tap K block of the item's class in A3: cos part, then sin part, kbp blocks each
  int kbp;      // 64-tap K blocks per part
  int cls;      // bin class = mel plane = component of the per-frame edge vector
  int edge_im;  // the class's self-paired tap belongs to the sin part (Im) instead of the cos part (Re)
};

struct Dftf3Params {
  int num_pairs, num_items;
  Dftf3Item item[avld_ctx::kMaxItems];
  int a_kblocks;            // K blocks per frame tile of A3 (n_fft / 64)
  uint32_t idesc;
  long long M_total;
  const float* inv2;
  const float4* edge;
  const MelTap* taps;       // [num_items * 160]; .pad holds the bit pattern of the edge coefficient
  float* melpow;            // [classes][rows][n_mels]
  long long plane_stride;
  int F, n_mels;
  int dbg;                  // AVLD_BRINGUP builds only (always 0 otherwise): 1 = skip the epilogue math, 4 = no operand
                            // loads after the first pipeline fill (MMA issue rate alone), 8 / 16 = no A / no B loads after
                            // the first fill, 32 = one MMA pass (hi x hi) instead of three, 64 = cycle counters into `prof`
  unsigned long long* prof; // [grid][8] (bring-up)
};

namespace {
#ifdef AVLD_BRINGUP
__device__ __forceinline__ int dbg_of(const Dftf3Params& P) { return P.dbg; }
#define AVLD_PROF_T(var) const long long var = clock64()
#define AVLD_PROF_ADD(acc, t0) acc += clock64() - (t0)
#else
__device__ __forceinline__ constexpr int dbg_of(const Dftf3Params&) { return 0; }   // probes compile away
#define AVLD_PROF_T(var) do { } while (0)
#define AVLD_PROF_ADD(acc, t0) do { } while (0)
#endif
constexpr int kBM = 128, kBN = 160, kBK = 64;
constexpr int kImCol = kBN, kReColB = 2 * kBN;   // TMEM columns: Re_A at 0, Im at 160, Re_B at 320
constexpr int kSwz = 128;
constexpr int kABytes = kBM * kSwz;              // one of hi / lo: 16 KB
constexpr int kBBytes = (kBN / 2) * kSwz;        // this CTA's 80 rows: 10 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // 52 KB
constexpr int kStages = 4;
constexpr int kEpiWarps = 8;                     // two per TMEM lane quarter (16 warps of 40 columns measured the same time:
                                                 // the epilogue is bound by its instruction count, not by latency)
constexpr int kWarpCols = kBN / (kEpiWarps / 4); // 80 accumulator columns per epilogue warp
constexpr int kGC = 16;                          // columns per tcgen05.ld in the column loop
constexpr int kExtra = 512 + kEpiWarps * kWarpCols * static_cast<int>(sizeof(MelTap));   // barriers + per-warp tap records
constexpr int kSmemBytes = kStages * kStageBytes + kExtra + 1024;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kGroups = kWarpCols / kGC;         // 5 loads of Re and of Im per item and warp
static_assert(kGroups % 2 == 1, "the epilogue's column loop handles pairs of groups plus one");
static_assert(kSmemBytes <= 232448, "shared memory budget");

// four consecutive mel bands of one frame row in one L2 reduction
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  // no "memory" clobber: the kernel never reads the planes back, and a clobber would pin every shared-memory load of the
  // epilogue behind the previous reduction
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}

// Mel accumulation of one frame row (= one thread) over its warp's bins of an item.  Every bin feeds the bands `first` and
// `first + 1` (slaney triangles overlap pairwise); `first` is warp uniform and monotone over the item's bins, so a running
// pair (a0, a1) = partial sums of the bands (mcur, mcur + 1) is enough.  A finished band goes through a four-deep shift
// register and leaves as an aligned group of four (VEC), or as a scalar reduction when n_mels is not a multiple of four.
// Rows past the end of the batch (tail of the last tile) are not masked: their un-scale factor s2 is 0 and the operand rows
// behind them hold finite stale values, so they add exact zeros to plane rows nobody reads in this pass.
template <bool VEC>
struct MelSink {
  float* pcur;            // address of band `mcur` in this frame row of the class's plane
  float s2;               // un-scale factor: a power of two, 0 (row past the batch) or NaN (poisoned chunk)
  int mcur, n_mels;
  float q0, q1, q2, q3;
  __device__ __forceinline__ void emit(float acc) {          // band `mcur` is complete
    const float val = acc * s2;
    if (VEC) {
      q0 = q1; q1 = q2; q2 = q3; q3 = val;
      if ((mcur & 3) == 3) red_add_v4(pcur - 3, q0, q1, q2, q3);
    } else if (val != 0.f) {
      atomicAdd(pcur, val);
    }
    ++mcur;
    ++pcur;
  }
  __device__ __forceinline__ void finish(float a0, float a1) {
    if (mcur < n_mels) emit(a0);
    if (mcur < n_mels) emit(a1);
    if (VEC)
      while (mcur & 3) emit(0.f);                            // pad the last group (n_mels % 4 == 0: never past the row)
  }
};

// kGC accumulator columns: edge term and |X|^2 (independent chains), then the mel accumulation; a band change is a warp-
// uniform branch
template <bool VEC>
__device__ __forceinline__ void mel_group(const uint32_t (&re)[kGC], const uint32_t (&im)[kGC], const MelTap* taps, float e_re,
                                          float e_im, MelSink<VEC>& sink, float& a0, float& a1) {
  MelTap tp[kGC];         // all tap records of the group first (independent 16-byte shared-memory loads)
#pragma unroll
  for (int j = 0; j < kGC; ++j) tp[j] = taps[j];
  float pw[kGC];
#pragma unroll
  for (int j = 0; j < kGC; ++j) {
    const float coef = __int_as_float(tp[j].pad);
    const float a = fmaf(e_re, coef, __uint_as_float(re[j]));
    const float b = fmaf(e_im, coef, __uint_as_float(im[j]));
    pw[j] = fmaf(a, a, b * b);
  }
#pragma unroll
  for (int j = 0; j < kGC; ++j) {
    if (sink.mcur < tp[j].first) {
#pragma unroll 1
      do {
        sink.emit(a0);
        a0 = a1;
        a1 = 0.f;
      } while (sink.mcur < tp[j].first);
    }
    a0 = fmaf(tp[j].w0, pw[j], a0);
    a1 = fmaf(tp[j].w1, pw[j], a1);
  }
}
}  // namespace

template <bool VEC>   // four mel bands per reduction (n_mels % 4 == 0)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dftf3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB_hi,
             const __grid_constant__ CUtensorMap tmB_lo, const Dftf3Params P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
  uint8_t* tail = smem + kStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);     // [8]  (used in the leader)
  uint64_t* empty_bar = full_bar + 8;                         // [8]  (per CTA)
  uint64_t* acc_full = empty_bar + 8;                         // [1]  the item's Re and Im are complete (per CTA)
  uint64_t* re_empty = acc_full + 1;                          // [2]  Re_A / Re_B drained (leader)
  uint64_t* im_empty = re_empty + 2;                          // [1]  Im drained (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(im_empty + 1);
  MelTap* s_taps = reinterpret_cast<MelTap*>(tail + 512);     // [kEpiWarps][kWarpCols]: each epilogue warp's bins of its current item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_clusters = static_cast<int>(ncluster_id_x());
  const int cluster = static_cast<int>(cluster_id_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);          // leader's expect_tx arrive + the peer producer's arrive
      mbar_init(&empty_bar[s], 1);         // one multicast commit
    }
    mbar_init(acc_full, 1);
    mbar_init(&re_empty[0], 2 * kEpiWarps);       // lane 0 of the epilogue warps of both CTAs
    mbar_init(&re_empty[1], 2 * kEpiWarps);
    mbar_init(im_empty, 2 * kEpiWarps);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers exist before anything can signal them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // whole warp on the warp-uniform schedule, one elected lane issues (same reason as for the MMA issuer below)
    int stage = 0;
    uint32_t phase = 0;
#ifdef AVLD_BRINGUP
    long long prof_wait0 = 0;
    const long long prof_t0 = clock64();
#endif
    // L2 prefetch of this CTA's A boxes, PF K blocks ahead of the loads
    constexpr int PF = 8;
    int pf_pair = cluster, pf_it = 0, pf_kb = 0;
    auto pf_step = [&]() {
      if (pf_pair < P.num_pairs) {
        const int y = ((pf_pair * 2 + static_cast<int>(rank)) * P.a_kblocks + P.item[pf_it].a_kb0 + pf_kb) * 256;
        if (elect_one()) tma_prefetch_2d(&tmA, 0, y);
        __syncwarp();
        if (++pf_kb == 2 * P.item[pf_it].kbp) {
          pf_kb = 0;
          if (++pf_it == P.num_items) { pf_it = 0; pf_pair += n_clusters; }
        }
      }
    };
    for (int i = 0; i < PF; ++i) pf_step();
    int filled = 0;
    for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
      const int a_row0 = (pair * 2 + static_cast<int>(rank)) * P.a_kblocks;      // first K block of this CTA's frame tile
      for (int it = 0; it < P.num_items; ++it) {
        const int a_kb0 = P.item[it].a_kb0, kbp = P.item[it].kbp, nkb = 2 * kbp;
        for (int kb = 0; kb < nkb; ++kb) {
          if (!(dbg_of(P) & 12)) pf_step();
          AVLD_PROF_T(tw);
          mbar_wait(&empty_bar[stage], phase ^ 1u, 100 + stage);
          AVLD_PROF_ADD(prof_wait0, tw);
          uint8_t* sa = smem + stage * kStageBytes;            // hi tile, then lo tile
          uint8_t* sb_hi = sa + 2 * kABytes;
          uint8_t* sb_lo = sb_hi + kBBytes;
          const bool probe = filled >= kStages;
          const bool no_a = (dbg_of(P) & (4 | 8)) && probe, no_b = (dbg_of(P) & (4 | 16)) && probe;
          ++filled;
          const int part = kb < kbp ? 0 : 1;
          const int bx = (kb - part * kbp) * kBK;
          const int by = (it * 2 + part) * kBN + static_cast<int>(rank) * (kBN / 2);
          if (elect_one()) {
            const uint32_t tx = 2u * ((no_a ? 0u : 2u * kABytes) + (no_b ? 0u : 2u * kBBytes));
            if (leader) {
              if (tx) mbar_arrive_expect_tx(&full_bar[stage], tx);
              else mbar_arrive(&full_bar[stage]);
            } else {
              mbar_arrive_cluster_relaxed(&full_bar[stage], 0);
            }
            // (L2 eviction-priority hints on these loads measured no difference)
            if (!no_a) tma_load_2d_pair(sa, &tmA, &full_bar[stage], 0, (a_row0 + a_kb0 + kb) * 256);
            if (!no_b) {
              tma_load_2d_pair(sb_hi, &tmB_hi, &full_bar[stage], bx, by);
              tma_load_2d_pair(sb_lo, &tmB_lo, &full_bar[stage], bx, by);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
#ifdef AVLD_BRINGUP
    if ((dbg_of(P) & 64) && lane == 0) {
      P.prof[blockIdx.x * 8 + 3] = clock64() - prof_t0;
      P.prof[blockIdx.x * 8 + 4] = prof_wait0;
    }
#endif
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The whole warp walks the (warp-uniform) schedule so that addresses and descriptors stay in uniform registers and
    // one elected lane issues.  Under `if (lane == 0)` the compiler brackets every UTCHMMA with ELECT + 5 R2UR.BROADCAST
    // (cuobjdump), which paced the issue at ~115 cycles per MMA whatever N: more than the 80 tensor cycles of an N = 160 MMA.
    if (leader) {
      int stage = 0;
      uint32_t phase = 0, n_item = 0;               // items issued so far: item n uses Re buffer n & 1
#ifdef AVLD_BRINGUP
      long long prof_wait0 = 0, prof_wait1 = 0;
      const long long prof_t0 = clock64();
      const bool one_pass = (dbg_of(P) & 32) != 0;
#else
      constexpr bool one_pass = false;
#endif
      for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
        for (int it = 0; it < P.num_items; ++it) {
          const int kbp = P.item[it].kbp, nkb = 2 * kbp;
          const uint32_t re_buf = n_item & 1u;
          for (int kb = 0; kb < nkb; ++kb) {
            if (kb == 0) {                   // this Re buffer was last used by item n - 2, Im by item n - 1
              AVLD_PROF_T(tw);
              mbar_wait(&re_empty[re_buf], ((n_item >> 1) & 1u) ^ 1u, 200 + static_cast<int>(re_buf));
              AVLD_PROF_ADD(prof_wait1, tw);
              tcgen05_fence_after();
            } else if (kb == kbp) {
              AVLD_PROF_T(tw);
              mbar_wait(im_empty, (n_item & 1u) ^ 1u, 202);
              AVLD_PROF_ADD(prof_wait1, tw);
              tcgen05_fence_after();
            }
            AVLD_PROF_T(tw);
            mbar_wait(&full_bar[stage], phase, 300 + stage);
            AVLD_PROF_ADD(prof_wait0, tw);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (kb < kbp ? re_buf * static_cast<uint32_t>(kReColB) : static_cast<uint32_t>(kImCol));
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
                if (!one_pass) {
                  umma_f16_pair(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                  umma_f16_pair(d_tmem, da_hi + koff, db_lo + koff, P.idesc, 1u);
                }
              }
              umma_commit_pair(&empty_bar[stage], 0x3);                    // stage reusable in both CTAs
              if (kb == nkb - 1) umma_commit_pair(acc_full, 0x3);          // Re and Im complete in both CTAs
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          ++n_item;
        }
      }
#ifdef AVLD_BRINGUP
      if ((dbg_of(P) & 64) && lane == 0) {
        P.prof[blockIdx.x * 8 + 0] = clock64() - prof_t0;
        P.prof[blockIdx.x * 8 + 1] = prof_wait0;
        P.prof[blockIdx.x * 8 + 2] = prof_wait1;
      }
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    // warp w may touch TMEM lanes 32 (w % 4) .. +32; the two warps of a lane quarter split the item's 160 bins
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int b0 = sub * kWarpCols;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(b0);
    const bool skip = (dbg_of(P) & 1) != 0;
    MelTap* my_taps = s_taps + (warp - 2) * kWarpCols;
    uint32_t n_item = 0;
#ifdef AVLD_BRINGUP
    long long prof_wait0 = 0, prof_wait1 = 0;
    const long long prof_t0 = clock64();
#endif
    auto release = [&](uint64_t* bar) {        // this warp has read its last column of an accumulator
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar);
        else mbar_arrive_cluster_relaxed(bar, 0);
      }
    };
    for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
      const long long g = static_cast<long long>(pair) * 2 * kBM + static_cast<long long>(rank) * kBM + row;
      const bool valid = g < P.M_total;
      const float s2 = valid ? P.inv2[g / P.F] : 0.f;
      const float4 edge = valid ? P.edge[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int it = 0; it < P.num_items; ++it, ++n_item) {
        const int cls = P.item[it].cls;
        // the class's self-paired tap joins Re or Im (coefficient +-2^10 or 0 per bin); the other part is used as it is
        const float e_cls = cls == 0 ? edge.x : (cls == 1 ? edge.y : edge.z);
        const float e_re = P.item[it].edge_im ? 0.f : e_cls, e_im = P.item[it].edge_im ? e_cls : 0.f;
        const uint32_t t_re = t_acc + (n_item & 1u) * static_cast<uint32_t>(kReColB), t_im = t_acc + static_cast<uint32_t>(kImCol);
        // this warp's tap records of the item: fetched (L2) before the wait, parked in shared memory for the column loop
        {
          const uint4* gt = reinterpret_cast<const uint4*>(P.taps + it * kBN + b0);
          static_assert(kWarpCols > 64 && kWarpCols <= 96, "tap staging below assumes three rounds");
          const uint4 tr0 = __ldg(gt + lane), tr1 = __ldg(gt + 32 + lane);
          const uint4 tr2 = lane < kWarpCols - 64 ? __ldg(gt + 64 + lane) : make_uint4(0u, 0u, 0u, 0u);
          reinterpret_cast<uint4*>(my_taps)[lane] = tr0;
          reinterpret_cast<uint4*>(my_taps)[32 + lane] = tr1;
          if (lane < kWarpCols - 64) reinterpret_cast<uint4*>(my_taps)[64 + lane] = tr2;
          __syncwarp();
        }
        {
          AVLD_PROF_T(tw);
          mbar_wait(acc_full, n_item & 1u, 400);
          AVLD_PROF_ADD(prof_wait0, tw);
        }
        tcgen05_fence_after();
        const int m_first = my_taps[0].first;
        MelSink<VEC> sink{P.melpow + cls * P.plane_stride + g * P.n_mels + m_first, s2, m_first, P.n_mels, 0.f, 0.f, 0.f, 0.f};
        float a0 = 0.f, a1 = 0.f;
        // Column loop: groups of kGC columns, the next group's Re / Im travelling from TMEM under the current group's math
        // (two register buffers, so the loop is rolled around a pair of groups plus one).  Fully unrolled -- 80 column
        // bodies, each with a taken branch over its band-change block, 80 KB of code -- the eight epilogue warps ran at
        // ~180 cycles per column, bound by instruction fetch; with three copies of the group body it is ~110, and what is
        // left is the per-column uniform branch itself (BSSY / BRA / BSYNC, ~45 cycles with two warps per scheduler;
        // profiles/r02b_dftf3_source_hotspots.txt).  A single copy in a loop of ten 8-column groups, 16 epilogue warps of
        // 40 columns, or predicated band changes with scalar reductions measured no better (1.08 / 1.02 / 1.25 ms vs 1.01).
        uint32_t re0[kGC], im0[kGC], re1[kGC], im1[kGC];
        tmem_ld16(t_re, re0);
        tmem_ld16(t_im, im0);
#pragma unroll 1
        for (int q = 0; q < kGroups - 1; q += 2) {
          tmem_ld_wait();
          tmem_ld16(t_re + (q + 1) * kGC, re1);
          tmem_ld16(t_im + (q + 1) * kGC, im1);
          if (!skip) mel_group<VEC>(re0, im0, my_taps + q * kGC, e_re, e_im, sink, a0, a1);
          tmem_ld_wait();
          tmem_ld16(t_re + (q + 2) * kGC, re0);
          tmem_ld16(t_im + (q + 2) * kGC, im0);
          if (!skip) mel_group<VEC>(re1, im1, my_taps + (q + 1) * kGC, e_re, e_im, sink, a0, a1);
        }
        tmem_ld_wait();
        release(im_empty);                   // all of this warp's columns are in registers: the issuer may overwrite Im
        release(&re_empty[n_item & 1u]);     // (next sin part) and this Re buffer (the item after the next)
        if (!skip) mel_group<VEC>(re0, im0, my_taps + (kGroups - 1) * kGC, e_re, e_im, sink, a0, a1);
        if (!skip) sink.finish(a0, a1);
        __syncwarp();                        // every lane is done with my_taps before the next item overwrites it
      }
    }
#ifdef AVLD_BRINGUP
    if ((dbg_of(P) & 64) && warp == 2 && lane == 0) {
      P.prof[blockIdx.x * 8 + 5] = clock64() - prof_t0;
      P.prof[blockIdx.x * 8 + 6] = prof_wait0;
      P.prof[blockIdx.x * 8 + 7] = prof_wait1;
    }
#endif
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer may still be reading operands / signalling our barriers
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
#endif
}

int launch_dftf3(avld_ctx* c, int n, cudaStream_t st) {
  Dftf3Params P{};
  const long long rows = static_cast<long long>(n) * c->F;
  const int m_tiles = static_cast<int>((rows + kBM - 1) / kBM);
  P.num_pairs = (m_tiles + 1) / 2;
  P.num_items = c->f2_items;
  for (int it = 0; it < c->f2_items; ++it)
    P.item[it] = {c->f2_item[it].a_col0 / kBK, c->f2_item[it].kbp, c->f2_item[it].cls, c->f2_item[it].edge_im};
  P.a_kblocks = c->a3_kblocks;
  P.idesc = avld_make_idesc(0, 0, 256, kBN);
  P.M_total = rows;
  P.inv2 = c->d_inv2;
  P.edge = c->d_edge;
  P.taps = c->d_taps3;
  P.melpow = c->d_melpow;
  P.plane_stride = c->melpow_plane;
  P.F = c->F;
  P.n_mels = c->M;
  P.dbg = 0;
  P.prof = nullptr;
#ifdef AVLD_BRINGUP
  static unsigned long long* d_prof = nullptr;
  {
    const char* d = getenv("AVLD_DBG");
    P.dbg = d ? atoi(d) : 0;
    const char* dn = getenv("AVLD_DBG_N");      // with AVLD_DBG=4: MMA N override, results are garbage
    if (dn && (P.dbg & 4)) P.idesc = avld_make_idesc(0, 0, 256, atoi(dn));
    if (P.dbg & 64) {
      if (!d_prof) AVLD_CUDA(cudaMalloc(&d_prof, 1024 * 8 * sizeof(unsigned long long)));
      AVLD_CUDA(cudaMemsetAsync(d_prof, 0, 1024 * 8 * sizeof(unsigned long long), st));
      P.prof = d_prof;
    }
  }
#endif
  // the epilogue accumulates mel outputs with reductions, one plane per bin class; the planes are zero on entry: cleared at
  // context creation and again by logmel_post_kernel as it reads them (the two always run as a pair)
  if (c->planes_dirty)     // only after a pass that failed between the two kernels
    AVLD_CUDA(cudaMemsetAsync(c->d_melpow, 0, static_cast<size_t>(c->melpow_plane) * c->f2_classes * sizeof(float), st));
  c->planes_dirty = true;
  const int grid = 2 * std::min(P.num_pairs, c->sm_count / 2);
  if (grid < 2) return AVLD_OK;
  const bool vec = (c->M % 4 == 0) && (c->melpow_plane % 4 == 0);
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(dftf3_kernel<true>), kSmemBytes));
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(dftf3_kernel<false>), kSmemBytes));
  {
    LaunchScope ls(c, ST_STFT_MEL, st);
    if (vec) dftf3_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(c->tm_A3, c->tm_B3_hi, c->tm_B3_lo, P);
    else dftf3_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(c->tm_A3, c->tm_B3_hi, c->tm_B3_lo, P);
  }
  AVLD_CUDA(cudaGetLastError());
#ifdef AVLD_BRINGUP
  if (P.prof && getenv("AVLD_PROF_DUMP")) {     // synchronous: bring-up runs only
    std::vector<unsigned long long> h(static_cast<size_t>(grid) * 8);
    AVLD_CUDA(cudaStreamSynchronize(st));
    AVLD_CUDA(cudaMemcpy(h.data(), d_prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < grid; ++b)
      for (int i = 0; i < 8; ++i) s[i] += static_cast<double>(h[b * 8 + i]);
    const double nl = grid / 2.0, na = grid;
    fprintf(stderr,
            "dftf3 prof (cycles, mean per CTA): issuer total %.0f wait_full %.0f wait_tmem_empty %.0f | producer total %.0f "
            "wait_empty %.0f | epilogue(w2) total %.0f wait_re %.0f wait_im %.0f\n",
            s[0] / nl, s[1] / nl, s[2] / nl, s[3] / na, s[4] / na, s[5] / na, s[6] / na, s[7] / na);
  }
#endif
  return AVLD_OK;
}

}  // namespace avld
