// dftf4.cu -- dftf3.cu with both 160-bin tiles of the odd bin class fed from ONE pass over that class's A columns.
// EXPERIMENTAL, opt-in (AVLD_DFT_DUAL=1); written at the end of round 1 after the GPU budget was spent: it compiles for
// sm_100a but HAS NOT RUN ON A B200 YET (tests/test_gpu_features.py::test_dual_tile_kernel_matches_default is skipped
// unless AVLD_TEST_DUAL=1).  The default path is dftf3.cu and does not touch this file.
//
// Why: dftf3_kernel is fed at the practical L2 -> SM rate and streams the odd class's A columns twice, once
// per tile of that class (TMEM holds one Re/Im accumulator set).  Here a K block of A is loaded once and multiplied into
// both tiles, which takes the operand bytes per frame pair and CTA from 48 x 52 KB to 16 x 72 + 16 x 52 KB (-20 %).
//
//   work per frame pair  group 0 (dual): odd tiles 0 and 1, K loop cos part then sin part, per K block: A once, B twice
//                        group 1, 2    : bins = 0 / 2 mod 4, one tile each, as in dftf3
//   rings                A: 4 slots x (128 x 64 hi + lo) = 32 KB; B: 4 slots x (80 x 64 hi + lo) = 20 KB; separate full /
//                        empty barriers, so that a K block of the dual group takes one A slot and two B slots
//   TMEM                 three regions of 160 columns (R0 = 0, R1 = 160, R2 = 320); the dual group wants four accumulators
//                        (Re0 Re1 Im0 Im1), so Im1 goes where Re0 was once the epilogue has pulled Re0 into registers:
//                            dual : Re0 -> R0, Re1 -> R1, Im0 -> R2, Im1 -> R0      bins 0 mod 4 : Re -> R2, Im -> R1
//                            bins 2 mod 4 : Re -> R0, Im -> R2
//                        every region has its own full / empty barrier pair; issuer and epilogue walk the same fixed
//                        sequence of uses per region (R0: 3 per frame pair, R1: 2, R2: 3) and keep one parity bit each
//   warps                0 = TMA producer (both CTAs), 1 = MMA issuer (leader), 2..9 = epilogue (both CTAs), as in dftf3
// Accumulation order per output element is that of dftf3 (same K blocks, same three passes), the mel atomics see the same
// values, so features are expected to be bit-identical to the default path.
//
// AVLD_DFT_DUAL=2 additionally shares B between the two CTA pairs of a 4-CTA cluster (template parameter kPairs = 2): the
// pairs work on neighbouring frame pairs in step, every CTA loads only a quarter of a B tile (40 rows) and TMA multicast
// delivers it to the CTA of the same parity in the other pair as well, so each SM pulls half of its B bytes from L2; a B
// slot is free again once BOTH issuers have committed (b_empty counts kPairs commits, multicast to the whole cluster).
// The instruction form is the one of CUTLASS' SM100_TMA_2SM_LOAD_MULTICAST (cta_group::2 + .multicast::cluster, barrier
// address masked to the pair leader of every destination).  Same status: compiles, protocol modelled, never run.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avld {

struct Dftf4Group {
  int a_col0;   // first A column of the group's class (cos part | sin part)
  int kbp;      // 64-tap K blocks per part
  int item0;    // first work item (= B row block, tap table block) of the group
  int tiles;    // 1 or 2
};

struct Dftf4Params {
  int num_pairs, num_groups;
  Dftf4Group group[3];
  int cls[4], edge_im[4];   // per work item
  uint32_t idesc;
  long long M_total;
  const float* inv2;
  const float4* edge;
  const MelTap* taps;       // [4 * 160]
  float* melpow;
  long long plane_stride;
  int F, n_mels;
};

namespace {
constexpr int kBM = 128, kBN = 160, kBK = 64;
constexpr int kSwz = 128;
constexpr int kAHalf = kBM * kSwz;               // hi or lo of an A block: 16 KB
constexpr int kBHalf = (kBN / 2) * kSwz;         // hi or lo of this CTA's 80 B rows: 10 KB
constexpr int kASlot = 2 * kAHalf, kBSlot = 2 * kBHalf;
constexpr int kSA = 4, kSB = 4;
constexpr int kRingBytes = kSA * kASlot + kSB * kBSlot;      // 208 KB
constexpr int kExtra = 12288;
constexpr int kSmemBytes = kRingBytes + kExtra + 1024;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kWarpCols = kBN / 2;
constexpr int kGroups = kWarpCols / 16;
constexpr int kRegions = 3;
static_assert(kSmemBytes <= 232448, "shared memory budget");
static_assert(kRegions * kBN <= 512, "TMEM columns");

// TMEM region of (group, tile, part); see the table in the header
__device__ __forceinline__ int region_of(int g, int t, int part) {
  if (g == 0) return t == 0 ? (part == 0 ? 0 : 2) : (part == 0 ? 1 : 0);
  if (g == 1) return part == 0 ? 2 : 1;
  return part == 0 ? 0 : 2;
}
}  // namespace

// multicast TMA load of a CTA pair member: the box lands at the same offset in every CTA of `mask`, the bytes are counted
// on the barrier at this offset of each destination's pair leader
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                                    uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// kPairs CTA pairs per cluster (launched with a runtime cluster dimension of 2 * kPairs); tmB_* have 80-row boxes for
// kPairs = 1 and 40-row boxes for kPairs = 2
template <int kPairs>
__global__ void __launch_bounds__(kThreads, 1)
dftf4_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const Dftf4Params P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* ring_a = smem;
  uint8_t* ring_b = smem + kSA * kASlot;
  uint8_t* tail = smem + kRingBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);       // [kSA] leader: expect_tx arrive + the peer producer's arrive
  uint64_t* a_empty = a_full + kSA;                           // [kSA] per CTA: one multicast commit
  uint64_t* b_full = a_empty + kSA;                           // [kSB]
  uint64_t* b_empty = b_full + kSB;                           // [kSB]
  uint64_t* r_full = b_empty + kSB;                           // [kRegions] per CTA: accumulator complete
  uint64_t* r_empty = r_full + kRegions;                      // [kRegions] leader: columns drained by all epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + kRegions);
  MelTap* s_taps = reinterpret_cast<MelTap*>(tail + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();                    // rank in the cluster
  const uint32_t rank = crank & 1u;                            // rank in the CTA pair
  const uint32_t cpair = crank >> 1;                           // which pair of the cluster
  const uint32_t lead = crank & ~1u;                           // cluster rank of this pair's leader
  const bool leader = rank == 0;
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * cpair));
  const uint16_t all_mask = static_cast<uint16_t>((1u << (2 * kPairs)) - 1u);
  const int n_clusters = static_cast<int>(ncluster_id_x());
  const int cluster = static_cast<int>(cluster_id_x());
  // frame pairs: the cluster takes kPairs neighbouring ones per round; a pair past the end still walks the schedule (its A
  // rows are out of range = zero fill, its epilogue rows are invalid) because the B slots are shared with the other pair
  const int first_fp = cluster * kPairs + static_cast<int>(cpair), fp_step = n_clusters * kPairs;
  const int num_rounds = (P.num_pairs + kPairs - 1) / kPairs;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < kSA; ++s) {
      mbar_init(&a_full[s], 2);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kSB; ++s) {
      mbar_init(&b_full[s], 2);
      mbar_init(&b_empty[s], kPairs);      // one multicast commit per issuer of the cluster
    }
    for (int r = 0; r < kRegions; ++r) {
      mbar_init(&r_full[r], 1);
      mbar_init(&r_empty[r], 2 * kEpiWarps);   // lane 0 of the epilogue warps of both CTAs
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  for (int i = threadIdx.x; i < 4 * kBN; i += blockDim.x) s_taps[i] = P.taps[i];
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    // L2 prefetch of this CTA's A rows, PF K blocks ahead of the loads
    constexpr int PF = 8;
    int pf_pair = first_fp, pf_g = 0, pf_kb = 0;
    auto pf_step = [&]() {
      if (pf_pair < P.num_pairs) {
        const int y = pf_pair * 2 * kBM + static_cast<int>(rank) * kBM;
        const int x = P.group[pf_g].a_col0 + pf_kb * kBK;
        if (elect_one()) {
          tma_prefetch_2d(&tmA_hi, x, y);
          tma_prefetch_2d(&tmA_lo, x, y);
        }
        __syncwarp();
        if (++pf_kb == 2 * P.group[pf_g].kbp) {
          pf_kb = 0;
          if (++pf_g == P.num_groups) { pf_g = 0; pf_pair += fp_step; }
        }
      }
    };
    for (int i = 0; i < PF; ++i) pf_step();
    for (int round = cluster; round < num_rounds; round += n_clusters) {
      const int pair = round * kPairs + static_cast<int>(cpair);
      const int ay = pair * 2 * kBM + static_cast<int>(rank) * kBM;
      for (int g = 0; g < P.num_groups; ++g) {
        const int a_col0 = P.group[g].a_col0, kbp = P.group[g].kbp, item0 = P.group[g].item0, tiles = P.group[g].tiles;
        for (int kb = 0; kb < 2 * kbp; ++kb) {
          pf_step();
          const int part = kb < kbp ? 0 : 1;
          mbar_wait(&a_empty[sa], pa ^ 1u, 100 + sa);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * kASlot);
            else mbar_arrive_cluster(&a_full[sa], lead);
            uint8_t* da = ring_a + sa * kASlot;
            tma_load_2d_pair(da, &tmA_hi, &a_full[sa], a_col0 + kb * kBK, ay);
            tma_load_2d_pair(da + kAHalf, &tmA_lo, &a_full[sa], a_col0 + kb * kBK, ay);
          }
          __syncwarp();
          if (++sa == kSA) { sa = 0; pa ^= 1u; }
          const int bx = (kb - part * kbp) * kBK;
          for (int t = 0; t < tiles; ++t) {
            const int by = ((item0 + t) * 2 + part) * kBN + static_cast<int>(rank) * (kBN / 2);
            mbar_wait(&b_empty[sb], pb ^ 1u, 110 + sb);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * kBSlot);
              else mbar_arrive_cluster(&b_full[sb], lead);
              uint8_t* db = ring_b + sb * kBSlot;
              if (kPairs == 1) {
                tma_load_2d_pair(db, &tmB_hi, &b_full[sb], bx, by);
                tma_load_2d_pair(db + kBHalf, &tmB_lo, &b_full[sb], bx, by);
              } else {                     // this CTA's share of the 80 rows, delivered to the same-parity CTA of every pair
                constexpr int kShare = (kBN / 2) / kPairs;
                const int sy = by + static_cast<int>(cpair) * kShare;
                uint8_t* ds = db + static_cast<int>(cpair) * kShare * kSwz;
                uint16_t mc = 0;
#pragma unroll
                for (int q = 0; q < kPairs; ++q) mc |= static_cast<uint16_t>(1u << (2 * q + static_cast<int>(rank)));
                tma_load_2d_pair_mc(ds, &tmB_hi, &b_full[sb], bx, sy, mc);
                tma_load_2d_pair_mc(ds + kBHalf, &tmB_lo, &b_full[sb], bx, sy, mc);
              }
            }
            __syncwarp();
            if (++sb == kSB) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // whole warp on the warp-uniform schedule, one elected lane around the tcgen05 instructions (see dftf3.cu)
    if (leader) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t used = 0;                       // bit r = parity of the completed uses of TMEM region r
      for (int round = cluster; round < num_rounds; round += n_clusters) {
        for (int g = 0; g < P.num_groups; ++g) {
          const int kbp = P.group[g].kbp, tiles = P.group[g].tiles;
          for (int part = 0; part < 2; ++part) {
            for (int kb = 0; kb < kbp; ++kb) {
              mbar_wait(&a_full[sa], pa, 300 + sa);
              tcgen05_fence_after();
              const uint32_t a_hi = smem_u32(ring_a + sa * kASlot), a_lo = a_hi + kAHalf;
              const uint64_t da_hi = make_smem_desc(a_hi, kSwz), da_lo = make_smem_desc(a_lo, kSwz);
              for (int t = 0; t < tiles; ++t) {
                const int r = region_of(g, t, part);
                if (kb == 0) {                 // the region's previous accumulator must have been drained
                  mbar_wait(&r_empty[r], ((used >> r) & 1u) ^ 1u, 200 + r);
                  tcgen05_fence_after();
                }
                mbar_wait(&b_full[sb], pb, 310 + sb);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(r * kBN);
                const uint32_t b_hi = smem_u32(ring_b + sb * kBSlot), b_lo = b_hi + kBHalf;
                const uint64_t db_hi = make_smem_desc(b_hi, kSwz), db_lo = make_smem_desc(b_lo, kSwz);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t koff = static_cast<uint64_t>(k * 2);
                    umma_f16_pair(d_tmem, da_hi + koff, db_hi + koff, P.idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_f16_pair(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                    umma_f16_pair(d_tmem, da_hi + koff, db_lo + koff, P.idesc, 1u);
                  }
                  umma_commit_pair(&b_empty[sb], all_mask);                  // this pair is done with the B slot (all CTAs hear it)
                  if (kb == kbp - 1) umma_commit_pair(&r_full[r], pair_mask);   // accumulator complete in both CTAs
                }
                __syncwarp();
                if (kb == kbp - 1) used ^= 1u << r;
                if (++sb == kSB) { sb = 0; pb ^= 1u; }
              }
              if (elect_one()) umma_commit_pair(&a_empty[sa], pair_mask);    // A slot reusable in both CTAs
              __syncwarp();
              if (++sa == kSA) { sa = 0; pa ^= 1u; }
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int b0 = sub * kWarpCols;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t used = 0;                         // same bookkeeping as the issuer's
    auto release = [&](int r) {                // this warp is done with region r
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&r_empty[r]);
        else mbar_arrive_cluster(&r_empty[r], lead);
      }
      used ^= 1u << r;
    };
    for (int round = cluster; round < num_rounds; round += n_clusters) {
      const int pair = round * kPairs + static_cast<int>(cpair);
      const long long gr = static_cast<long long>(pair) * 2 * kBM + static_cast<long long>(rank) * kBM + row;
      const bool valid = gr < P.M_total;
      const float s2 = valid ? P.inv2[gr / P.F] : 0.f;
      const float4 edge = valid ? P.edge[gr] : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < P.num_groups; ++g) {
        for (int t = 0; t < P.group[g].tiles; ++t) {
          const int it = P.group[g].item0 + t;
          const int cls = P.cls[it];
          const int r_re = region_of(g, t, 0), r_im = region_of(g, t, 1);
          float* mrow = P.melpow + cls * P.plane_stride + gr * P.n_mels;
          const float e_cls = cls == 0 ? edge.x : (cls == 1 ? edge.y : edge.z);
          const float e_re = P.edge_im[it] ? 0.f : e_cls, e_im = P.edge_im[it] ? e_cls : 0.f;
          // ---- Re: into registers
          uint32_t re[kGroups][16];
          mbar_wait(&r_full[r_re], (used >> r_re) & 1u, 400 + r_re);
          tcgen05_fence_after();
#pragma unroll
          for (int q = 0; q < kGroups; ++q) tmem_ld16(t_acc + static_cast<uint32_t>(r_re * kBN + b0 + q * 16), re[q]);
          tmem_ld_wait();
          release(r_re);
          // ---- Im: streamed, combined with the held Re
          const MelTap* item_taps = s_taps + it * kBN + b0;
          int mcur = item_taps[0].first;
          float a0 = 0.f, a1 = 0.f;
          mbar_wait(&r_full[r_im], (used >> r_im) & 1u, 410 + r_im);
          tcgen05_fence_after();
#pragma unroll
          for (int q = 0; q < kGroups; ++q) {
            uint32_t im[16];
            tmem_ld16(t_acc + static_cast<uint32_t>(r_im * kBN + b0 + q * 16), im);
            tmem_ld_wait();
            if (q == kGroups - 1) release(r_im);   // last Im read of this warp: the columns may be overwritten
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const MelTap tp = item_taps[q * 16 + j];
              const float coef = __int_as_float(tp.pad);
              const float a = fmaf(e_re, coef, __uint_as_float(re[q][j]));
              const float b = fmaf(e_im, coef, __uint_as_float(im[j]));
              const float pw = (a * a + b * b) * s2;
              if (mcur < tp.first) {
#pragma unroll 1
                while (mcur < tp.first) {
                  if (valid && a0 != 0.f) atomicAdd(mrow + mcur, a0);
                  a0 = a1;
                  a1 = 0.f;
                  ++mcur;
                }
              }
              a0 = fmaf(tp.w0, pw, a0);
              a1 = fmaf(tp.w1, pw, a1);
            }
          }
          if (valid && a0 != 0.f && mcur < P.n_mels) atomicAdd(mrow + mcur, a0);
          if (valid && a1 != 0.f && mcur + 1 < P.n_mels) atomicAdd(mrow + mcur + 1, a1);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer may still be reading operands / signalling our barriers
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
#endif
}

// Opt-in, and only for the item pattern the TMEM schedule above is written for: two tiles of one class, then two
// single-tile classes (the three-level fold at n_fft = 2048 with the reference's mel band).
// AVLD_DFT_DUAL = 1: CTA pairs on their own; 2: B shared by multicast between the two pairs of a 4-CTA cluster.
static int dftf4_mode(const avld_ctx* c) {
  const char* e = getenv("AVLD_DFT_DUAL");                        // read per pass, like AVLD_DFT_GEN: A/B in one process
  const int mode = e != nullptr ? atoi(e) : 0;
  if (mode < 1 || mode > 2 || !c->dft_fold2 || c->f2_levels != 3 || c->f2_items != 4) return 0;
  const auto* it = c->f2_item;
  const bool ok = it[0].a_col0 == it[1].a_col0 && it[0].cls == it[1].cls && it[0].kbp == it[1].kbp && it[2].cls != it[0].cls &&
                  it[3].cls != it[2].cls && it[3].cls != it[0].cls && it[0].kbp >= 1 && it[2].kbp >= 1 && it[3].kbp >= 1;
  if (!ok) return 0;
  if (mode == 2 && c->sm_count % 4 != 0) return 1;
  return mode;
}

bool dftf4_supported(const avld_ctx* c) { return dftf4_mode(c) != 0; }

template <int kPairs>
static int launch_dual(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const Dftf4Params& P, int sm_count, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    AVLD_CUDA(cudaFuncSetAttribute(dftf4_kernel<kPairs>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    AVLD_CUDA(cudaFuncSetAttribute(dftf4_kernel<kPairs>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  constexpr int kCluster = 2 * kPairs;
  const int rounds = (P.num_pairs + kPairs - 1) / kPairs;
  const int grid = kCluster * std::min(rounds, sm_count / kCluster);
  if (grid < kCluster) return AVLD_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AVLD_CUDA(cudaLaunchKernelEx(&cfg, dftf4_kernel<kPairs>, a_hi, a_lo, b_hi, b_lo, P));
  return AVLD_OK;
}

int launch_stft_mel_fold2_dual(avld_ctx* c, int n, cudaStream_t st) {
  const int mode = dftf4_mode(c);
  Dftf4Params P{};
  const long long rows = static_cast<long long>(n) * c->F;
  const int m_tiles = static_cast<int>((rows + kBM - 1) / kBM);
  P.num_pairs = (m_tiles + 1) / 2;
  P.num_groups = 3;
  P.group[0] = {c->f2_item[0].a_col0, c->f2_item[0].kbp, 0, 2};
  P.group[1] = {c->f2_item[2].a_col0, c->f2_item[2].kbp, 2, 1};
  P.group[2] = {c->f2_item[3].a_col0, c->f2_item[3].kbp, 3, 1};
  for (int it = 0; it < 4; ++it) {
    P.cls[it] = c->f2_item[it].cls;
    P.edge_im[it] = c->f2_item[it].edge_im;
  }
  P.idesc = avld_make_idesc(0, 0, 256, kBN);
  P.M_total = rows;
  P.inv2 = c->d_inv2;
  P.edge = c->d_edge;
  P.taps = c->d_taps3;
  P.melpow = c->d_melpow;
  P.plane_stride = c->melpow_plane;
  P.F = c->F;
  P.n_mels = c->M;
  // 40-row boxes of the DFT matrix for the multicast variant: encoded on first use only, so that the default path's
  // context creation is untouched by this experiment
  static const avld_ctx* tm_owner = nullptr;
  static CUtensorMap tm_q_hi, tm_q_lo;
  if (mode == 2 && tm_owner != c) {
    const int Q = c->p.n_fft / 4;
    const uint64_t rows3 = static_cast<uint64_t>(c->f2_items) * 2 * kBN;
    AVLD_TRY(encode_tmap_2d(&tm_q_hi, c->d_B3hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Q, rows3, static_cast<uint64_t>(Q) * 2, 64, kBN / 4, 128));
    AVLD_TRY(encode_tmap_2d(&tm_q_lo, c->d_B3lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, Q, rows3, static_cast<uint64_t>(Q) * 2, 64, kBN / 4, 128));
    tm_owner = c;
  }
  if (c->planes_dirty)     // only after a pass that failed between the GEMM and logmel_post_kernel (see dftf3.cu)
    AVLD_CUDA(cudaMemsetAsync(c->d_melpow, 0, static_cast<size_t>(c->melpow_plane) * c->f2_classes * sizeof(float), st));
  c->planes_dirty = true;
  LaunchScope ls(c, ST_STFT_MEL, st);
  if (mode == 2) return launch_dual<2>(c->tm_A2_hi, c->tm_A2_lo, tm_q_hi, tm_q_lo, P, c->sm_count, st);
  return launch_dual<1>(c->tm_A2_hi, c->tm_A2_lo, c->tm_B3_hi, c->tm_B3_lo, P, c->sm_count, st);
}

}  // namespace avld
