// gemm3.cuh -- split-precision GEMM core on tcgen05 / TMEM (sm_100a) for the encoder's dense layers and the
// implicit-GEMM convolutions that convh.cu does not cover (the STFT has its own CTA-pair kernel, dftf3.cu).
//
//   D[128 x BN] (fp32, TMEM) += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo       per 16-wide K step
//
// Every fp32 operand is pre-split into a 16-bit "hi" part and a bf16 "lo" remainder, so three
// kind::f16 MMAs reproduce an fp32-accurate product (error ~2^-20) at the 16-bit tensor rate.
// The reference computes this path in float64 FFT / float32 torch, and the parity target is
// 1e-3 on latents behind an 80 dB log floor: one 16-bit pass is not enough.
//
// Structure (one CTA per SM, persistent over M tiles, 192 threads):
//   warp 0 / lane 0 : TMA producer  -- A (hi, lo) and B (hi, lo) boxes into a STAGES-deep smem ring
//   warp 1 / lane 0 : MMA issuer    -- tcgen05.mma, tcgen05.commit frees smem slots / publishes TMEM
//   warps 2..5      : epilogue      -- tcgen05.ld from a double-buffered TMEM accumulator
// A tiles are addressed two ways (a_mode):
//   0  plain row-major [M, K]                      box (k0, m0)
//   2  NHWC activations [N, H, W, C]: one box per filter tap, shifted by (kh - pad, kw - pad);
//      TMA out-of-bounds zero fill *is* the convolution's zero padding, the tensor map's element
//      strides *are* the convolution's stride
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace avld {

enum { EPI_PLAIN = 0, EPI_CONV = 2 };

struct Gemm3Params {
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int dbg_shift, dbg_baseoff;   // bring-up probe: A descriptor start shifted by dbg_shift rows, descriptor base-offset field
  int split_n;  // 1: a work item is one (m tile, n tile) pair (dense layers with few m tiles); 0: one m tile, all n tiles
  uint32_t idesc_hh, idesc_lh, idesc_hl;  // (A_hi,B_hi) (A_lo,B_hi) (A_hi,B_lo)
  int a_mode;
  // a_mode 2 geometry
  int tiles_w, tiles_h, tw, th, ksize, cblocks, cblk, pad, stride;   // stride: of the convolution (the tensor map traverses with it)
  // common epilogue
  long long M_total;
  int N_total;
  const float* bias;
  int relu;
  float* out_f32;           // PLAIN: C[M_total][ldc]
  int ldc;
  __nv_bfloat16* out_hi;    // PLAIN/CONV: split output for the next tensor-core layer
  __nv_bfloat16* out_lo;
  // CONV epilogue
  int H, W, Cout, pool;     // H, W: conv output size before pooling
  int pool_avg;             // the fused 2x2 pooling averages (after the activation) instead of taking the maximum
  const __nv_bfloat16* res_hi;   // optional residual of the CONV epilogue (output shape, no pooling): out = act(conv + bias + res)
  const __nv_bfloat16* res_lo;
};

// non-template dispatcher (all instantiations live in gemm3.cu)
int run_gemm3(avld_ctx* c, int bn, int swz, int epi, const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo,
              const CUtensorMap& tmB_hi, const CUtensorMap& tmB_lo, const Gemm3Params& P, cudaStream_t st);

#ifdef AVLD_GEMM3_IMPL
template <int BN, int SWZ>
struct Gemm3Cfg {
  static constexpr int BM = 128;
  static constexpr int BK = SWZ / 2;                       // 16-bit elements per swizzle row
  static constexpr int A_BYTES = BM * SWZ;                 // one of hi / lo
  static constexpr int B_BYTES = BN * SWZ;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int EXTRA = 16384;                      // barriers, taps, alignment slack
  static constexpr int STAGES_RAW = (227 * 1024 - EXTRA) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EXTRA;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(STAGES >= 2, "tile too large for shared memory");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
};

// threads per CTA: TMA warp + MMA warp + 4 epilogue warps
// The CONV epilogue moves an NHWC row per pixel (output, optionally a residual operand): with four warps only 128 threads carry
// all of a CTA's memory traffic and layers with little MMA work per tile (1x1 shortcuts) starve -- eight warps, two per TMEM
// lane quarter with half the columns each.
template <int EPI>
struct Gemm3Threads { static constexpr int epi_warps = EPI == EPI_CONV ? 8 : 4; static constexpr int value = 64 + 32 * epi_warps; };

template <int BN, int SWZ, int EPI>
__global__ void __launch_bounds__(Gemm3Threads<EPI>::value, 1)
gemm3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
             const Gemm3Params P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  using Cfg = Gemm3Cfg<BN, SWZ>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NACC = 2;                       // TMEM accumulator buffers
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
  uint8_t* tail = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                           // [STAGES]
  uint64_t* tmem_full = empty_bar + STAGES;                          // [2]
  uint64_t* tmem_empty = tmem_full + 2;                              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 512);              // bias[N_total], zero padded

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 32 * Gemm3Threads<EPI>::epi_warps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  {
    const int nb = P.num_n_tiles * BN;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_bias[i] = (P.bias != nullptr && i < P.N_total) ? P.bias[i] : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkb = P.num_k_blocks;
  const int n_items = P.split_n ? P.num_m_tiles * P.num_n_tiles : P.num_m_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = P.split_n ? item / P.num_n_tiles : item;
        const int nt_begin = P.split_n ? item - mt * P.num_n_tiles : 0;
        const int nt_end = P.split_n ? nt_begin + 1 : P.num_n_tiles;
        int img = 0, h0 = 0, w0 = 0;
        if (P.a_mode == 2) {
          const int per_img = P.tiles_w * P.tiles_h;
          img = mt / per_img;
          const int r = mt - img * per_img;
          h0 = (r / P.tiles_w) * P.th;
          w0 = (r % P.tiles_w) * P.tw;
        }
        for (int nt = nt_begin; nt < nt_end; ++nt) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u, 100 + stage);
            uint8_t* sa_hi = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sa_lo = sa_hi + Cfg::A_BYTES;
            uint8_t* sb_hi = sa_lo + Cfg::A_BYTES;
            uint8_t* sb_lo = sb_hi + Cfg::B_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (P.a_mode == 0) {
              tma_load_2d(sa_hi, &tmA_hi, &full_bar[stage], kb * Cfg::BK, mt * Cfg::BM);
              tma_load_2d(sa_lo, &tmA_lo, &full_bar[stage], kb * Cfg::BK, mt * Cfg::BM);
            } else {
              const int tap = kb / P.cblocks;
              const int cb = kb - tap * P.cblocks;
              const int kh = tap / P.ksize, kw = tap - kh * P.ksize;
              tma_load_4d(sa_hi, &tmA_hi, &full_bar[stage], cb * P.cblk, w0 * P.stride + kw - P.pad, h0 * P.stride + kh - P.pad, img);
              tma_load_4d(sa_lo, &tmA_lo, &full_bar[stage], cb * P.cblk, w0 * P.stride + kw - P.pad, h0 * P.stride + kh - P.pad, img);
            }
            tma_load_2d(sb_hi, &tmB_hi, &full_bar[stage], kb * Cfg::BK, nt * BN);
            tma_load_2d(sb_lo, &tmB_lo, &full_bar[stage], kb * Cfg::BK, nt * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // the whole warp walks the warp-uniform schedule (descriptors stay in uniform registers), one elected lane issues:
    // under `if (lane == 0)` every UTCHMMA is bracketed by ELECT + R2UR.BROADCAST moves (~115 cycles per MMA, measured)
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n_sub = P.split_n ? 1 : P.num_n_tiles;
        for (int sub = 0; sub < n_sub; ++sub) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, 200 + acc);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
          const uint32_t id_hh = P.idesc_hh, id_lh = P.idesc_lh, id_hl = P.idesc_hl;
          for (int kb = 0; kb < nkb; ++kb) {
            const int kb_acc = kb;
            mbar_wait(&full_bar[stage], phase, 300 + stage);
            tcgen05_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t a_lo = a_hi + Cfg::A_BYTES;
            const uint32_t b_hi = a_lo + Cfg::A_BYTES;
            const uint32_t b_lo = b_hi + Cfg::B_BYTES;
            const uint64_t dbg_bits = (static_cast<uint64_t>(P.dbg_baseoff & 7) << 49) + static_cast<uint64_t>((P.dbg_shift * SWZ) >> 4);
            const uint64_t da_hi = make_smem_desc(a_hi, SWZ) + dbg_bits, da_lo = make_smem_desc(a_lo, SWZ) + dbg_bits;
            const uint64_t db_hi = make_smem_desc(b_hi, SWZ), db_lo = make_smem_desc(b_lo, SWZ);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < Cfg::BK / 16; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k * 2);  // 16 elements * 2 B = 32 B, in 16 B units
                umma_f16(d_tmem, da_hi + koff, db_hi + koff, id_hh, (kb_acc | k) != 0 ? 1u : 0u);
                umma_f16(d_tmem, da_lo + koff, db_hi + koff, id_lh, 1u);
                umma_f16(d_tmem, da_hi + koff, db_lo + koff, id_hl, 1u);
              }
              umma_commit(&empty_bar[stage]);                    // smem slot reusable once these MMAs retire
              if (kb == nkb - 1) umma_commit(&tmem_full[acc]);   // accumulator complete
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter + 32) belong to this warp
    constexpr int kConvCols = BN / (Gemm3Threads<EPI>::epi_warps / 4);   // CONV: columns per epilogue warp (BN >= 32)
    const int conv_col0 = ((warp - 2) >> 2) * kConvCols;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = P.split_n ? item / P.num_n_tiles : item;
      const int nt_begin = P.split_n ? item - mt * P.num_n_tiles : 0;
      const int nt_end = P.split_n ? nt_begin + 1 : P.num_n_tiles;
      for (int nt = nt_begin; nt < nt_end; ++nt) {
        // EPI_CONV: row = pixel (h_local * tw + w_local) of a th x tw output tile, tw in {8, 16}: the 2x2 pooling partners
        // are lane ^ 1 (w) and lane ^ tw (h) of the same warp.  The residual operand of the first column group is
        // requested before the wait for the accumulator: it does not depend on it.
        int img = 0, h = 0, w = 0, OH = 1, OW = 1;
        bool inb = false, writer = false, has_res = false;
        size_t opix = 0;
        Res16 res[4];                   // ring of four column groups: a whole 64-column tile row, or the next 64 columns
        if (EPI == EPI_CONV) {
          const int per_img = P.tiles_w * P.tiles_h;
          img = mt / per_img;
          const int r = mt - img * per_img;
          h = (r / P.tiles_w) * P.th + row / P.tw;
          w = (r % P.tiles_w) * P.tw + row % P.tw;
          inb = (h < P.H) && (w < P.W);
          OH = P.H / P.pool; OW = P.W / P.pool;
          writer = inb && (P.pool == 1 || (((h | w) & 1) == 0));
          opix = (static_cast<size_t>(img) * OH + h / P.pool) * OW + w / P.pool;
          has_res = P.res_hi != nullptr && writer && (nt + 1) * BN <= P.Cout;
          if (has_res) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (16 * g < kConvCols) res16_load(res[g], P.res_hi, P.res_lo, opix * P.Cout + nt * BN + conv_col0 + 16 * g);
          }
        }
        mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
        tcgen05_fence_after();
        const uint32_t t_acc = tmem_base + lane_base + static_cast<uint32_t>(acc * BN);

        if (EPI == EPI_PLAIN) {
          const long long m = static_cast<long long>(mt) * Cfg::BM + row;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_acc + c0, v);
            tmem_ld_wait();
            const int n0 = nt * BN + c0;
            if (m < P.M_total && n0 < P.N_total) {
              float o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float t = __uint_as_float(v[j]) + s_bias[n0 + j];
                if (P.relu) t = relu_nan(t);
                o[j] = t;
              }
              if (P.out_f32 != nullptr) {
                float* dst = P.out_f32 + m * P.ldc + n0;
                if (n0 + 16 <= P.N_total && (P.ldc & 3) == 0) {
#pragma unroll
                  for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                } else {
                  for (int j = 0; j < 16 && n0 + j < P.N_total; ++j) dst[j] = o[j];
                }
              }
              if (P.out_hi != nullptr) {
                uint32_t hw[8], lw[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) split_bf16x2(o[j], o[j + 1], hw[j >> 1], lw[j >> 1]);
                __nv_bfloat16* dh = P.out_hi + m * P.ldc + n0;
                __nv_bfloat16* dl = P.out_lo + m * P.ldc + n0;
                if (n0 + 16 <= P.N_total && (P.ldc & 7) == 0) {
                  reinterpret_cast<uint4*>(dh)[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                  reinterpret_cast<uint4*>(dh)[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                  reinterpret_cast<uint4*>(dl)[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                  reinterpret_cast<uint4*>(dl)[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                } else {
                  const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(hw);
                  const __nv_bfloat16* ls = reinterpret_cast<const __nv_bfloat16*>(lw);
                  for (int j = 0; j < 16 && n0 + j < P.N_total; ++j) { dh[j] = hs[j]; dl[j] = ls[j]; }
                }
              }
            }
          }
        } else {
          // one 16-column group; `slot` holds its residual operand (requested a whole 64-column row ahead: a group ahead was
          // not enough to cover the DRAM latency -- ncu: the adds waited ~1 000 cycles per group, 4.3 us per 128-pixel tile)
          auto do_group = [&](const int c0, Res16& slot) {
            uint32_t v[16];
            tmem_ld16(t_acc + c0, v);
            tmem_ld_wait();
            const int n0 = nt * BN + c0;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n0 + j);
              o[j] = __uint_as_float(v[j]) + b4.x;
              o[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              o[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
              o[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
            if (has_res) {                      // fused residual add, before the activation; this slot is refilled four groups ahead
              const Res16 cur = slot;
              if (c0 + 64 < conv_col0 + kConvCols) res16_load(slot, P.res_hi, P.res_lo, opix * P.Cout + n0 + 64);
              res16_add(o, cur);
            }
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = relu_nan(o[j]);
            }
            if (P.pool == 2) {   // 16 independent shuffles in flight per stage
              float u[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) u[j] = __shfl_xor_sync(0xffffffffu, o[j], 1);
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = P.pool_avg ? o[j] + u[j] : max_nan(o[j], u[j]);
#pragma unroll
              for (int j = 0; j < 16; ++j) u[j] = __shfl_xor_sync(0xffffffffu, o[j], P.tw);
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = P.pool_avg ? (o[j] + u[j]) * 0.25f : max_nan(o[j], u[j]);
            }
            if (writer && n0 < P.Cout) {
              // packed conversions (one full-rate F2FP per pair): the scalar form is 32 XU-pipe F2F per column group, which
              // paced layers with little MMA work per tile (a 1x1 shortcut: 4.4 us per 128-pixel tile)
              uint32_t hw[8], lw[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) split_bf16x2(o[j], o[j + 1], hw[j >> 1], lw[j >> 1]);
              __nv_bfloat16* dh = P.out_hi + opix * P.Cout + n0;
              __nv_bfloat16* dl = P.out_lo + opix * P.Cout + n0;
              if (n0 + 16 <= P.Cout) {
                reinterpret_cast<uint4*>(dh)[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                reinterpret_cast<uint4*>(dh)[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                reinterpret_cast<uint4*>(dl)[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                reinterpret_cast<uint4*>(dl)[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
              } else {
                const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(hw);
                const __nv_bfloat16* ls = reinterpret_cast<const __nv_bfloat16*>(lw);
                for (int j = 0; j < 16 && n0 + j < P.Cout; ++j) { dh[j] = hs[j]; dl[j] = ls[j]; }
              }
            }
          };
#pragma unroll 1
          for (int c64 = 0; c64 < kConvCols; c64 += 64) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (16 * g < kConvCols) do_group(conv_col0 + c64 + 16 * g, res[g]);
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
#endif
}

template <int BN, int SWZ, int EPI>
int launch_gemm3(avld_ctx* c, const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi,
                 const CUtensorMap& tmB_lo, const Gemm3Params& P, cudaStream_t st) {
  const int sm_count = c->sm_count;
  using Cfg = Gemm3Cfg<BN, SWZ>;
  auto kfn = gemm3_kernel<BN, SWZ, EPI>;
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(kfn), Cfg::SMEM_BYTES));
  const int items = P.split_n ? P.num_m_tiles * P.num_n_tiles : P.num_m_tiles;
  int grid = items < sm_count ? items : sm_count;
  if (grid < 1) return AVLD_OK;
  kfn<<<grid, Gemm3Threads<EPI>::value, Cfg::SMEM_BYTES, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, P);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

#endif  // AVLD_GEMM3_IMPL

}  // namespace avld
