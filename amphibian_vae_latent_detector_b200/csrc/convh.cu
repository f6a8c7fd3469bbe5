// convh.cu -- 3x3 "same" convolution as an implicit GEMM on tcgen05 with HALO REUSE.
//
// The first implicit-GEMM version (gemm3_kernel<EPI_CONV>) loaded one 128-pixel activation box per filter tap, i.e.
// every input pixel nine times, and was bound by the TMA unit's row rate (~2.3 cycles per <=128-byte row per SM).
// Here the (16+2) x (8+2) input halo of a 16 x 8 output tile is loaded ONCE per channel block; the nine taps are nine
// shared-memory descriptors over the same tile: start shifted by (kh * 10 + kw) rows, 8-row groups (= 8 pixels of one
// output row) strided by the halo pitch of 10 rows.  That works because the UMMA swizzle is a function of the absolute
// shared-memory address (bring-up probe tools/probe_shift.py: shifted descriptors read correctly with base_offset 0),
// so any row of a TMA-written swizzled tile can be addressed.  Filter weights stay resident in shared memory when all
// taps fit (<= 80 KB, e.g. C_in = 32, C_out = 64), otherwise they stream through a ring, one (tap, channel block) per slot.
//
//   warp 0 lane 0  TMA producer: halo tiles (4-D box, out-of-bounds zero fill = the padding) and weight blocks
//   warp 1 lane 0  MMA issuer: per tap and 16-channel step three kind::f16 MMAs (hi*hi, lo*hi, hi*lo), bf16 operands
//   warps 2..5     epilogue: bias + ReLU + 2x2 max-pool (warp shuffles) + bf16 hi/lo split, NHWC stores
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avld {

struct ConvhParams {
  int n_tiles;                 // images * tiles_h * tiles_w
  int tiles_w, tiles_h;
  int cblocks;                 // channel blocks of CBLK input channels
  int resident;                // weights resident in shared memory (loaded once)
  uint32_t idesc;
  uint32_t idesc2;              // BN = 64: the N = 2 BN form over the adjacent [W_hi | W_lo] blocks
  int H, W, Cout, relu, pool;  // conv output size (= input size), before pooling
  int pool_avg;                // the 2x2 pooling averages (after bias and ReLU) instead of taking the maximum
  int dbg;                     // AVLD_BRINGUP builds only: timing probe (AVLD_CONVH_DBG = 2, wrong results): hi*hi pass only
  const float* bias;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  const __nv_bfloat16* res_hi;  // optional residual (same NHWC shape as the output, no pooling): out = act(conv + bias + res)
  const __nv_bfloat16* res_lo;
};

namespace {
constexpr int kTW = 8, kTH = 16;                    // output tile: 16 rows x 8 columns = 128 pixels
constexpr int kHW = kTW + 2, kHH = kTH + 2;         // halo
constexpr int kHaloRows = kHW * kHH;                // 180 pixels

template <int BN, int CBLK>
struct ConvhCfg {
  static constexpr int ROWB = CBLK * 2;                                   // bytes per pixel row = swizzle span (64 / 128)
  static constexpr int HALO_BYTES = ((kHaloRows * ROWB + 1023) / 1024) * 1024;   // one of hi / lo, 1 KB aligned
  static constexpr int HSTAGE = 2 * HALO_BYTES;
  static constexpr int WBLK = BN * ROWB;                                  // one (tap, cblk) weight block, one of hi / lo
  static constexpr int WSTAGE = 2 * WBLK;
  static constexpr int EXTRA = 8192;
  static constexpr int BUDGET = 227 * 1024 - EXTRA - 1024;
  static constexpr int RES_BYTES = 9 * WSTAGE;                            // all taps of ONE channel block
  static constexpr int HSTAGES_RES = (BUDGET - RES_BYTES) / HSTAGE > 6 ? 6 : (BUDGET - RES_BYTES) / HSTAGE;
  static constexpr int HSTAGES_STR = 2;
  static constexpr int WSTAGES_STR = (BUDGET - HSTAGES_STR * HSTAGE) / WSTAGE > 8 ? 8 : (BUDGET - HSTAGES_STR * HSTAGE) / WSTAGE;
  // BN = 64: an N = 64 MMA reads 6 KB of shared memory for 32 cycles of tensor work and is bound by that read.  The hi and
  // lo weight blocks of a tap are adjacent in shared memory, so A_hi x [W_hi | W_lo] is ONE N = 128 MMA into two column
  // ranges of the accumulator (A is read once for both), A_lo x W_hi a second one into the first range; the epilogue adds
  // the two ranges.  Two MMAs and 14 KB per K step instead of three and 18 KB.
  static constexpr bool CONCAT = BN == 64 || BN == 128;   // BN = 128: 20 KB instead of 24 KB of operand reads per K step (N = 256 + N = 128)
  static constexpr int ACC_COLS = CONCAT ? 2 * BN : BN;                   // TMEM columns of one accumulator
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = BUDGET + EXTRA + 1024;
  // epilogue warps: one per TMEM lane quarter (two per quarter, half the columns each, measured no faster with BN = 64:
  // 280.5 vs 282.3 us per 1024-chunk launch -- the epilogue is not what paces that layer)
  static constexpr int EPI_WARPS = 4;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};
}  // namespace

template <int BN, int CBLK>
__global__ void __launch_bounds__(ConvhCfg<BN, CBLK>::THREADS, 1)
convh_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const ConvhParams P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  using Cfg = ConvhCfg<BN, CBLK>;
  constexpr int ROWB = Cfg::ROWB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
  const bool resident = P.resident != 0;
  const int hstages = resident ? Cfg::HSTAGES_RES : Cfg::HSTAGES_STR;
  const int wstages = resident ? 0 : Cfg::WSTAGES_STR;
  uint8_t* s_halo = smem;                                             // [hstages][hi | lo]
  uint8_t* s_w = smem + hstages * Cfg::HSTAGE;                        // resident: [9][hi | lo]; streamed: [wstages][hi | lo]
  uint8_t* tail = smem + Cfg::BUDGET;
  uint64_t* h_full = reinterpret_cast<uint64_t*>(tail);              // [8]
  uint64_t* h_empty = h_full + 8;                                     // [8]
  uint64_t* w_full = h_empty + 8;                                     // [8]  (resident: w_full[0] only)
  uint64_t* w_empty = w_full + 8;                                     // [8]
  uint64_t* tmem_full = w_empty + 8;                                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 512);               // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmW_hi);
    tma_prefetch_desc(&tmW_lo);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&h_full[s], 1);
      mbar_init(&h_empty[s], 1);
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 32 * Cfg::EPI_WARPS);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  for (int i = threadIdx.x; i < BN; i += blockDim.x) s_bias[i] = i < P.Cout ? P.bias[i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = P.tiles_w * P.tiles_h;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (resident) {           // all nine taps of the single channel block, once
        mbar_arrive_expect_tx(&w_full[0], 9 * Cfg::WSTAGE);
        for (int tap = 0; tap < 9; ++tap) {
          tma_load_2d(s_w + tap * Cfg::WSTAGE, &tmW_hi, &w_full[0], tap * CBLK, 0);
          tma_load_2d(s_w + tap * Cfg::WSTAGE + Cfg::WBLK, &tmW_lo, &w_full[0], tap * CBLK, 0);
        }
      }
      int hs = 0, ws = 0;
      uint32_t hphase = 0, wphase = 0;
      for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const int img = t / per_img, r = t - img * per_img;
        const int h0 = (r / P.tiles_w) * kTH, w0 = (r % P.tiles_w) * kTW;
        for (int cb = 0; cb < P.cblocks; ++cb) {
          mbar_wait(&h_empty[hs], hphase ^ 1u, 100 + hs);
          uint8_t* dst = s_halo + hs * Cfg::HSTAGE;
          mbar_arrive_expect_tx(&h_full[hs], 2 * kHaloRows * ROWB);
          tma_load_4d(dst, &tmA_hi, &h_full[hs], cb * CBLK, w0 - 1, h0 - 1, img);
          tma_load_4d(dst + Cfg::HALO_BYTES, &tmA_lo, &h_full[hs], cb * CBLK, w0 - 1, h0 - 1, img);
          if (++hs == hstages) { hs = 0; hphase ^= 1u; }
          if (!resident) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&w_empty[ws], wphase ^ 1u, 150 + ws);
              uint8_t* wd = s_w + ws * Cfg::WSTAGE;
              mbar_arrive_expect_tx(&w_full[ws], Cfg::WSTAGE);
              const int kx = (tap * P.cblocks + cb) * CBLK;
              tma_load_2d(wd, &tmW_hi, &w_full[ws], kx, 0);
              tma_load_2d(wd + Cfg::WBLK, &tmW_lo, &w_full[ws], kx, 0);
              if (++ws == wstages) { ws = 0; wphase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the (warp-uniform) schedule so that addresses and descriptors live in uniform registers;
    // one elected lane issues.  With `if (lane == 0)` around the loop every UTCHMMA paid ~75 cycles of R2UR traffic,
    // which made the small N = 64 MMAs (32 tensor cycles each) issue bound.
    {
      int hs = 0, ws = 0, acc = 0;
      uint32_t hphase = 0, wphase = 0, acc_phase = 0;
      if (resident) {
        mbar_wait(&w_full[0], 0, 250);
        tcgen05_fence_after();
      }
      const uint64_t sbo_fix = (static_cast<uint64_t>((kHW * ROWB) >> 4) << 32) - (static_cast<uint64_t>((8 * ROWB) >> 4) << 32);
      for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, 200 + acc);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * Cfg::ACC_COLS);
        for (int cb = 0; cb < P.cblocks; ++cb) {
          mbar_wait(&h_full[hs], hphase, 300 + hs);
          tcgen05_fence_after();
          const uint32_t halo_hi = smem_u32(s_halo + hs * Cfg::HSTAGE), halo_lo = halo_hi + Cfg::HALO_BYTES;
          const uint64_t dh_hi = make_smem_desc(halo_hi, ROWB) + sbo_fix, dh_lo = make_smem_desc(halo_lo, ROWB) + sbo_fix;
          // Taps fully unrolled: the tap shift of A ((kh * pitch + kw) rows) and, with resident weights, the weight block
          // address are compile-time offsets of descriptors that live in uniform registers.  With a rolled tap loop the
          // compiler rebuilt both descriptors in vector registers and moved them over with a dozen R2UR per tap; the issuing
          // warp was then busy all the time while the tensor pipe idled two thirds of it (ncu + AVLD_CONVH_DBG=2: a third of
          // the MMAs took only 14 - 26 % off the kernel).
          const bool lo_passes = !(P.dbg & 2);
          if (resident) {
            const uint64_t dw_hi = make_smem_desc(smem_u32(s_w), ROWB);
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint64_t shift = static_cast<uint64_t>((((tap / 3) * kHW + (tap % 3)) * ROWB) >> 4);
                const uint64_t da_hi = dh_hi + shift, da_lo = dh_lo + shift;
                const uint64_t db_hi = dw_hi + static_cast<uint64_t>((tap * Cfg::WSTAGE) >> 4);
                const uint64_t db_lo = db_hi + static_cast<uint64_t>(Cfg::WBLK >> 4);
#pragma unroll
                for (int k = 0; k < CBLK / 16; ++k) {
                  const uint64_t koff = static_cast<uint64_t>(k * 2);
                  if constexpr (Cfg::CONCAT) {
                    umma_f16(d_tmem, da_hi + koff, db_hi + koff, P.idesc2, (cb | tap | k) != 0 ? 1u : 0u);   // [hi*hi | hi*lo]
                    umma_f16(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                  } else {
                    umma_f16(d_tmem, da_hi + koff, db_hi + koff, P.idesc, (cb | tap | k) != 0 ? 1u : 0u);
                    if (lo_passes) {
                      umma_f16(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                      umma_f16(d_tmem, da_hi + koff, db_lo + koff, P.idesc, 1u);
                    }
                  }
                }
              }
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&w_full[ws], wphase, 350 + ws);
              tcgen05_fence_after();
              const uint64_t shift = static_cast<uint64_t>((((tap / 3) * kHW + (tap % 3)) * ROWB) >> 4);
              const uint64_t da_hi = dh_hi + shift, da_lo = dh_lo + shift;
              const uint64_t db_hi = make_smem_desc(smem_u32(s_w + ws * Cfg::WSTAGE), ROWB);
              const uint64_t db_lo = db_hi + static_cast<uint64_t>(Cfg::WBLK >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < CBLK / 16; ++k) {
                  const uint64_t koff = static_cast<uint64_t>(k * 2);
                  if constexpr (Cfg::CONCAT) {
                    umma_f16(d_tmem, da_hi + koff, db_hi + koff, P.idesc2, (cb | tap | k) != 0 ? 1u : 0u);   // [hi*hi | hi*lo]
                    umma_f16(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                  } else {
                    umma_f16(d_tmem, da_hi + koff, db_hi + koff, P.idesc, (cb | tap | k) != 0 ? 1u : 0u);
                    if (lo_passes) {
                      umma_f16(d_tmem, da_lo + koff, db_hi + koff, P.idesc, 1u);
                      umma_f16(d_tmem, da_hi + koff, db_lo + koff, P.idesc, 1u);
                    }
                  }
                }
                umma_commit(&w_empty[ws]);
              }
              __syncwarp();
              if (++ws == wstages) { ws = 0; wphase ^= 1u; }
            }
          }
          if (elect_one()) umma_commit(&h_empty[hs]);
          __syncwarp();
          if (++hs == hstages) { hs = 0; hphase ^= 1u; }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5, with BN = 64 also 6..9)
    const int quarter = warp & 3;
    constexpr int kColsPerWarp = BN / (Cfg::EPI_WARPS / 4);
    const int col_lo = ((warp - 2) >> 2) * kColsPerWarp;
    const int row = quarter * 32 + lane;                       // pixel (row / 8, row % 8) of the tile
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int OH = P.H / P.pool, OW = P.W / P.pool;
    for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
      const int img = t / per_img, r = t - img * per_img;
      const int h = (r / P.tiles_w) * kTH + row / kTW, w = (r % P.tiles_w) * kTW + row % kTW;
      const bool inb = (h < P.H) && (w < P.W);
      const bool writer = inb && (P.pool == 1 || (((h | w) & 1) == 0));
      const size_t opix = (static_cast<size_t>(img) * OH + h / P.pool) * OW + w / P.pool;
      // the residual operand of the first column group is requested before the wait for the accumulator
      const bool has_res = P.res_hi != nullptr && writer && P.pool == 1;
      Res16 res_next;
      if (has_res) res16_load(res_next, P.res_hi, P.res_lo, opix * P.Cout + col_lo);
      mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + lane_base + static_cast<uint32_t>(acc * Cfg::ACC_COLS);
#pragma unroll 1
      for (int c0 = col_lo; c0 < col_lo + kColsPerWarp; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_acc + c0, v);
        if constexpr (Cfg::CONCAT) {
          uint32_t v2[16];
          tmem_ld16(t_acc + BN + c0, v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          tmem_ld_wait();
        }
        if (P.pool == 2) {
          // 2x2 max-pool as a reduce-scatter over the four lanes of a window (partners lane ^ 1 (w) and lane ^ 8 (h), same
          // warp): each exchange halves the columns a lane keeps, so a lane ends with 4 of the 16 columns, pooled, and
          // does bias / ReLU / split / store for those only.  12 shuffles instead of 32 and a quarter of the conversion work
          // per lane; bias and ReLU commute with the max (both monotonic, fl(x + b) is monotonic in x), so the bits are the
          // same as pooling after them.
          const bool odd_w = (lane & 1) != 0, odd_h = (lane & kTW) != 0;
          const bool avg = P.pool_avg != 0;
          if (avg) {                          // the average does not commute with bias + ReLU: apply them first
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float t = __uint_as_float(v[j + u]) + bb[u];
                if (P.relu) t = relu_nan(t);
                v[j + u] = __float_as_uint(t);
              }
            }
          }
          float keep[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float lo8 = __uint_as_float(v[j]), hi8 = __uint_as_float(v[j + 8]);
            const float send = odd_w ? lo8 : hi8;
            const float mine = odd_w ? hi8 : lo8, other = __shfl_xor_sync(0xffffffffu, send, 1);
            keep[j] = avg ? mine + other : max_nan(mine, other);
          }
          float q4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float send = odd_h ? keep[j] : keep[j + 4];
            const float mine = odd_h ? keep[j + 4] : keep[j], other = __shfl_xor_sync(0xffffffffu, send, kTW);
            q4[j] = avg ? (mine + other) * 0.25f : max_nan(mine, other);
          }
          const int cq = c0 + (odd_w ? 8 : 0) + (odd_h ? 4 : 0);            // first of this lane's four output channels
          if (((h | 1) < P.H) && ((w | 1) < P.W) && cq < P.Cout) {         // the whole window lies inside the image
            const float4 b4 = avg ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(s_bias + cq);
            float o4[4] = {q4[0] + b4.x, q4[1] + b4.y, q4[2] + b4.z, q4[3] + b4.w};
            if (P.relu && !avg) {
#pragma unroll
              for (int j = 0; j < 4; ++j) o4[j] = relu_nan(o4[j]);
            }
            uint32_t hw[2], lw[2];                       // packed conversions (F2FP), not four XU-pipe F2F per plane
            split_bf16x2(o4[0], o4[1], hw[0], lw[0]);
            split_bf16x2(o4[2], o4[3], hw[1], lw[1]);
            *reinterpret_cast<uint2*>(P.out_hi + opix * P.Cout + cq) = make_uint2(hw[0], hw[1]);
            *reinterpret_cast<uint2*>(P.out_lo + opix * P.Cout + cq) = make_uint2(lw[0], lw[1]);
          }
          continue;
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
          o[j] = __uint_as_float(v[j]) + b4.x;
          o[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
          o[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
          o[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
        }
        if (has_res) {                          // fused residual add, before the activation; the next group's loads go out first
          const Res16 cur = res_next;
          if (c0 + 16 < col_lo + kColsPerWarp) res16_load(res_next, P.res_hi, P.res_lo, opix * P.Cout + c0 + 16);
          res16_add(o, cur);
        }
        if (P.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = relu_nan(o[j]);
        }
        if (writer && c0 < P.Cout) {
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) split_bf16x2(o[j], o[j + 1], hw[j >> 1], lw[j >> 1]);
          uint4* dh = reinterpret_cast<uint4*>(P.out_hi + opix * P.Cout + c0);
          uint4* dl = reinterpret_cast<uint4*>(P.out_lo + opix * P.Cout + c0);
          dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
#endif
}

template <int BN, int CBLK>
static int launch_one(avld_ctx* c, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                      const CUtensorMap& w_lo, ConvhParams P, cudaStream_t st) {
  const int sm_count = c->sm_count;
  using Cfg = ConvhCfg<BN, CBLK>;
  auto kfn = convh_kernel<BN, CBLK>;
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(kfn), Cfg::SMEM_BYTES));
  P.resident = (P.cblocks == 1 && Cfg::RES_BYTES <= 80 * 1024 && Cfg::HSTAGES_RES >= 2) ? 1 : 0;
  if (!P.resident && Cfg::WSTAGES_STR < 2) {
    set_error("convh: tile does not fit shared memory (BN=%d, CBLK=%d)", BN, CBLK);
    return AVLD_ERR_UNSUPPORTED;
  }
  const int grid = std::min(P.n_tiles, sm_count);
  if (grid < 1) return AVLD_OK;
  kfn<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(a_hi, a_lo, w_hi, w_lo, P);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

// halo tensor map of the layer input: [n][H][W][C] bf16, box {cblk, 10, 18, 1}
int convh_encode_input_map(CUtensorMap* out, const void* base, int n, int H, int W, int C, int cblk) {
  const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(n)};
  const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2, static_cast<uint64_t>(H) * W * C * 2};
  const uint32_t box[4] = {static_cast<uint32_t>(cblk), static_cast<uint32_t>(kHW), static_cast<uint32_t>(kHH), 1};
  return encode_tmap_4d(out, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, strides, box, static_cast<uint32_t>(cblk * 2));
}

bool convh_supported(int c_in, int c_out, int ksize, int w) {
  return ksize == 3 && (c_in == 32 || c_in % 64 == 0) && (c_out == 64 || c_out == 128) && (w % kTW == 0);
}

int launch_convh(avld_ctx* c, const OpDev& L, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int n, __nv_bfloat16* out_hi,
                 __nv_bfloat16* out_lo, const __nv_bfloat16* res_hi, const __nv_bfloat16* res_lo, cudaStream_t st) {
  ConvhParams P{};
  P.tiles_w = L.in_w / kTW;
  P.tiles_h = (L.in_h + kTH - 1) / kTH;
  P.n_tiles = n * P.tiles_w * P.tiles_h;
  P.cblocks = L.cblocks;
  P.idesc = avld_make_idesc(1, 1, 128, L.c_out);
  P.idesc2 = avld_make_idesc(1, 1, 128, 2 * L.c_out);
  P.H = L.in_h; P.W = L.in_w; P.Cout = L.c_out; P.relu = L.relu; P.pool = L.pool; P.pool_avg = L.pool_avg;
  P.bias = L.bias;
#ifdef AVLD_BRINGUP
  static const int dbg_env = std::getenv("AVLD_CONVH_DBG") ? std::atoi(std::getenv("AVLD_CONVH_DBG")) : 0;
  P.dbg = dbg_env;
#endif
  P.out_hi = out_hi;
  P.out_lo = out_lo;
  P.res_hi = res_hi;
  P.res_lo = res_lo;
  if (L.c_out == 64 && L.cblk == 32) return launch_one<64, 32>(c, a_hi, a_lo, L.tm_w_hi, L.tm_w_lo, P, st);
  if (L.c_out == 64 && L.cblk == 64) return launch_one<64, 64>(c, a_hi, a_lo, L.tm_w_hi, L.tm_w_lo, P, st);
  if (L.c_out == 128 && L.cblk == 32) return launch_one<128, 32>(c, a_hi, a_lo, L.tm_w_hi, L.tm_w_lo, P, st);
  if (L.c_out == 128 && L.cblk == 64) return launch_one<128, 64>(c, a_hi, a_lo, L.tm_w_hi, L.tm_w_lo, P, st);
  set_error("convh: unsupported shape C_out=%d cblk=%d", L.c_out, L.cblk);
  return AVLD_ERR_UNSUPPORTED;
}

}  // namespace avld
