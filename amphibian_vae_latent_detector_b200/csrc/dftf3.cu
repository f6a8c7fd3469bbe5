// dftf3.cu -- twice-folded STFT GEMM on CTA pairs (tcgen05 cta_group::2); operands from fold2.cu.
//
// The FFT bins with mel weight are split into classes (odd | 0 mod 4 | 2 mod 4, or even | odd in the two-level form);
// each class has its own A columns (cos part | sin part, N/4 or N/8 taps each) and is covered by work items of 160
// bins.  One item = 256 frames (a CTA pair) x 160 bins: K loop over the cos part (-> Re, TMEM columns 0..159) and then
// the sin part (-> Im, columns 256..415), N/4 taps each -- a quarter of the taps of the plain DFT GEMM and half of the
// once-folded one (dftf2.cu), for the same bins.  Split precision as everywhere: hi*hi + lo*hi + hi*lo, fp32 in TMEM.
//   warp 0 (both CTAs)          TMA producer: own 128 frames of A, own half (80 rows) of the B tile
//   warp 1 (leader only)        MMA issuer: tcgen05.mma.cta_group::2 (M 256, N 160), commits multicast to both CTAs
//   warps 2..9    (both CTAs)   epilogue: Re into registers as soon as the cos part is done (the issuer moves on to the
//                               sin part and, after it, straight to the next item's cos part), then Im streamed from
//                               TMEM; edge term, |X|^2, un-scale, sparse slaney mel, atomicAdd into the class's plane.
// Bit-reproducible: a warp covers 80 consecutive bins of one class (160 FFT bins), a mel filter spans <= 67 FFT bins, so
// every (frame, filter, class plane) cell receives at most two atomic contributions (a + b is order independent).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace avld {

struct Dftf3Item {
  int a_col0;   // first A column of the item's class: cos part | sin part, kbp * 64 taps each
  int kbp;      // 64-tap K blocks per part
  int cls;      // bin class = mel plane = component of the per-frame edge vector
  int edge_im;  // the class's self-paired tap belongs to the sin part (Im) instead of the cos part (Re)
};

struct Dftf3Params {
  int num_pairs, num_items;
  Dftf3Item item[8];
  uint32_t idesc;
  long long M_total;
  const float* inv2;
  const float4* edge;
  const MelTap* taps;       // [num_items * 160]; .pad holds the bit pattern of the edge coefficient
  float* melpow;            // [2 planes][rows][n_mels]
  long long plane_stride;
  int F, n_mels;
  int dbg;                  // bring-up: 1 = skip the epilogue math, 2 = A rows fixed (always L2 resident),
                            // 4 = no operand loads after the first pipeline fill (MMA issue rate alone)
};

namespace {
constexpr int kBM = 128, kBN = 160, kBK = 64, kImCol = 256;
constexpr int kSwz = 128;
constexpr int kABytes = kBM * kSwz;              // one of hi / lo: 16 KB
constexpr int kBBytes = (kBN / 2) * kSwz;        // this CTA's 80 rows: 10 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // 52 KB
constexpr int kStages = 4;
constexpr int kExtra = 12288;
constexpr int kSmemBytes = kStages * kStageBytes + kExtra + 1024;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kWarpCols = kBN / 2;               // 80 accumulator columns per epilogue warp
constexpr int kGroups = kWarpCols / 16;          // 5 tcgen05.ld x16 per half
static_assert(kSmemBytes <= 232448, "shared memory budget");
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dftf3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const Dftf3Params P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
  uint8_t* tail = smem + kStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);     // [8]  (used in the leader)
  uint64_t* empty_bar = full_bar + 8;                         // [8]  (per CTA)
  uint64_t* tmem_full = empty_bar + 8;                        // [2]  Re / Im complete (per CTA)
  uint64_t* tmem_empty = tmem_full + 2;                       // [2]  Re / Im columns drained (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  MelTap* s_taps = reinterpret_cast<MelTap*>(tail + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_clusters = static_cast<int>(ncluster_id_x());
  const int cluster = static_cast<int>(cluster_id_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);          // leader's expect_tx arrive + the peer producer's arrive
      mbar_init(&empty_bar[s], 1);         // one multicast commit
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&tmem_full[h], 1);
      mbar_init(&tmem_empty[h], 2 * kEpiWarps);   // lane 0 of the epilogue warps of both CTAs
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  for (int i = threadIdx.x; i < P.num_items * kBN; i += blockDim.x) s_taps[i] = P.taps[i];
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers exist before anything can signal them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // whole warp on the warp-uniform schedule, one elected lane issues (same reason as for the MMA issuer below)
    {
      int stage = 0;
      uint32_t phase = 0;
      // L2 prefetch of this CTA's A rows, PF K blocks ahead of the loads
      constexpr int PF = 8;
      int pf_pair = cluster, pf_it = 0, pf_kb = 0;
      auto pf_step = [&]() {
        if (pf_pair < P.num_pairs) {
          const int y = pf_pair * 2 * kBM + static_cast<int>(rank) * kBM;
          const int x = P.item[pf_it].a_col0 + pf_kb * kBK;
          if (elect_one()) {
            tma_prefetch_2d(&tmA_hi, x, y);
            tma_prefetch_2d(&tmA_lo, x, y);
          }
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
        const int ay = (P.dbg & 2) ? (cluster * 2 * kBM + static_cast<int>(rank) * kBM)
                                   : pair * 2 * kBM + static_cast<int>(rank) * kBM;
        for (int it = 0; it < P.num_items; ++it) {
          const int a_col0 = P.item[it].a_col0, kbp = P.item[it].kbp, nkb = 2 * kbp;
          for (int kb = 0; kb < nkb; ++kb) {
            if (!(P.dbg & 6)) pf_step();
            mbar_wait(&empty_bar[stage], phase ^ 1u, 100 + stage);
            uint8_t* sa_hi = smem + stage * kStageBytes;
            uint8_t* sa_lo = sa_hi + kABytes;
            uint8_t* sb_hi = sa_lo + kABytes;
            uint8_t* sb_lo = sb_hi + kBBytes;
            const bool stale = (P.dbg & 4) && filled >= kStages;     // bring-up: stale operands, barrier protocol only
            ++filled;
            const int part = kb < kbp ? 0 : 1;
            const int bx = (kb - part * kbp) * kBK;
            const int by = (it * 2 + part) * kBN + static_cast<int>(rank) * (kBN / 2);
            if (elect_one()) {
              if (stale) {
                if (leader) mbar_arrive(&full_bar[stage]);
                else mbar_arrive_cluster(&full_bar[stage], 0);
              } else {
                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
                else mbar_arrive_cluster(&full_bar[stage], 0);
                // (L2 eviction-priority hints on these loads -- evict_last for B and for A until its class's last
                // tile -- measured no difference: the K loop is bound by L2 -> SM throughput, not by misses)
                tma_load_2d_pair(sa_hi, &tmA_hi, &full_bar[stage], a_col0 + kb * kBK, ay);
                tma_load_2d_pair(sa_lo, &tmA_lo, &full_bar[stage], a_col0 + kb * kBK, ay);
                tma_load_2d_pair(sb_hi, &tmB_hi, &full_bar[stage], bx, by);
                tma_load_2d_pair(sb_lo, &tmB_lo, &full_bar[stage], bx, by);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The whole warp walks the (warp-uniform) schedule so that addresses and descriptors stay in uniform registers and
    // one elected lane issues.  Under `if (lane == 0)` the compiler brackets every UTCHMMA with ELECT + 5 R2UR.BROADCAST
    // (cuobjdump), which paced the issue at ~115 cycles per MMA whatever N (measured with AVLD_DBG=5 and N = 64..192):
    // more than the 80 tensor cycles of an N = 160 MMA.
    if (leader) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
        for (int it = 0; it < P.num_items; ++it) {
          const int kbp = P.item[it].kbp, nkb = 2 * kbp;
          for (int kb = 0; kb < nkb; ++kb) {
            if (kb == 0 || kb == kbp) {      // the part's accumulator columns must have been drained
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
              umma_commit_pair(&empty_bar[stage], 0x3);                    // stage reusable in both CTAs
              if (kb == kbp - 1) umma_commit_pair(&tmem_full[0], 0x3);     // Re complete in both CTAs
              if (kb == nkb - 1) umma_commit_pair(&tmem_full[1], 0x3);     // Im complete in both CTAs
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    // warp w may touch TMEM lanes 32 (w % 4) .. +32; the two warps of a lane quarter split the item's 160 bins
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int b0 = sub * kWarpCols;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool skip = (P.dbg & 1) != 0;
    uint32_t acc_phase = 0;
    for (int pair = cluster; pair < P.num_pairs; pair += n_clusters) {
      const long long g = static_cast<long long>(pair) * 2 * kBM + static_cast<long long>(rank) * kBM + row;
      const bool valid = g < P.M_total;
      const float s2 = valid ? P.inv2[g / P.F] : 0.f;
      const float4 edge = valid ? P.edge[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int it = 0; it < P.num_items; ++it) {
        const int cls = P.item[it].cls;
        float* mrow = P.melpow + cls * P.plane_stride + g * P.n_mels;
        const float e_cls = cls == 0 ? edge.x : (cls == 1 ? edge.y : edge.z);
        const float e_re = P.item[it].edge_im ? 0.f : e_cls, e_im = P.item[it].edge_im ? e_cls : 0.f;
        // ---- Re: into registers while the sin part is still being multiplied
        uint32_t re[kGroups][16];
        mbar_wait(&tmem_full[0], acc_phase, 400);
        tcgen05_fence_after();
#pragma unroll
        for (int q = 0; q < kGroups; ++q) tmem_ld16(t_acc + b0 + q * 16, re[q]);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tmem_empty[0]);
          else mbar_arrive_cluster(&tmem_empty[0], 0);
        }
        // ---- Im: streamed, combined with the held Re
        const MelTap* item_taps = s_taps + it * kBN + b0;
        int mcur = item_taps[0].first;
        float a0 = 0.f, a1 = 0.f;
        mbar_wait(&tmem_full[1], acc_phase, 401);
        tcgen05_fence_after();
#pragma unroll
        for (int q = 0; q < kGroups; ++q) {
          uint32_t im[16];
          tmem_ld16(t_acc + kImCol + b0 + q * 16, im);
          tmem_ld_wait();
          if (q == kGroups - 1) {            // last Im read of this warp: the columns may be overwritten
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) mbar_arrive(&tmem_empty[1]);
              else mbar_arrive_cluster(&tmem_empty[1], 0);
            }
          }
          if (!skip) {
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
        }
        if (valid && a0 != 0.f && mcur < P.n_mels) atomicAdd(mrow + mcur, a0);
        if (valid && a1 != 0.f && mcur + 1 < P.n_mels) atomicAdd(mrow + mcur + 1, a1);
        acc_phase ^= 1u;
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

int launch_stft_mel_fold2(avld_ctx* c, int n, cudaStream_t st) {
  Dftf3Params P{};
  const long long rows = static_cast<long long>(n) * c->F;
  const int m_tiles = static_cast<int>((rows + kBM - 1) / kBM);
  P.num_pairs = (m_tiles + 1) / 2;
  P.num_items = c->f2_items;
  for (int it = 0; it < c->f2_items; ++it)
    P.item[it] = {c->f2_item[it].a_col0, c->f2_item[it].kbp, c->f2_item[it].cls, c->f2_item[it].edge_im};
  P.idesc = avld_make_idesc(0, 0, 256, kBN);
  P.M_total = rows;
  P.inv2 = c->d_inv2;
  P.edge = c->d_edge;
  P.taps = c->d_taps3;
  P.melpow = c->d_melpow;
  P.plane_stride = c->melpow_plane;
  P.F = c->F;
  P.n_mels = c->M;
  {
    const char* d = getenv("AVLD_DBG");
    P.dbg = d ? atoi(d) : 0;
    const char* dn = getenv("AVLD_DBG_N");      // bring-up (with AVLD_DBG=5): MMA N override, results are garbage
    if (dn && (P.dbg & 4)) P.idesc = avld_make_idesc(0, 0, 256, atoi(dn));
  }
  static bool configured = false;
  if (!configured) {
    AVLD_CUDA(cudaFuncSetAttribute(dftf3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  AVLD_CHECK(static_cast<size_t>(P.num_items) * kBN * sizeof(MelTap) <= kExtra - 512, AVLD_ERR_UNSUPPORTED, "too many FFT bins");
  // the epilogue accumulates mel outputs with atomicAdd, one plane per bin class; the planes are zero on entry: cleared at
  // context creation and again by logmel_post_kernel as it reads them (the two always run as a pair)
  if (c->planes_dirty)     // only after a pass that failed between the two kernels
    AVLD_CUDA(cudaMemsetAsync(c->d_melpow, 0, static_cast<size_t>(c->melpow_plane) * c->f2_classes * sizeof(float), st));
  c->planes_dirty = true;
  const int grid = 2 * std::min(P.num_pairs, c->sm_count / 2);
  if (grid < 2) return AVLD_OK;
  LaunchScope ls(c, ST_STFT_MEL, st);
  dftf3_kernel<<<grid, kThreads, kSmemBytes, st>>>(c->tm_A2_hi, c->tm_A2_lo, c->tm_B3_hi, c->tm_B3_lo, P);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld
