// gemm3.cuh -- split-precision GEMM core on tcgen05 / TMEM (sm_100a), shared by the STFT
// (windowed DFT as a GEMM), the encoder's implicit-GEMM convolutions and its dense layers.
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
// A tiles are addressed three ways (a_mode):
//   0  plain row-major [M, K]                      box (k0, m0)
//   1  audio rows [n_rows, hop] (frame g, tap k lives at row g + k / hop, column k % hop: the STFT's
//      im2col is free -- frames are overlapping windows, so no frame matrix is ever materialised)
//   2  NHWC activations [N, H, W, C]: one box per filter tap, shifted by (kh - pad, kw - pad);
//      TMA out-of-bounds zero fill *is* the convolution's zero padding
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace avld {

enum { EPI_PLAIN = 0, EPI_DFT = 1, EPI_CONV = 2, EPI_DFTF = 3 };
// EPI_DFTF = folded STFT: per 256-bin N tile the first half of the K blocks (E | cos) accumulates the real parts
// into TMEM columns [0, 256), the second half (O | -sin) the imaginary parts into [256, 512): one 512-column
// accumulator, single buffered (the epilogue of a tile is not overlapped with the next tile's MMAs).

struct Gemm3Params {
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int dbg_shift, dbg_baseoff;   // bring-up probe: A descriptor start shifted by dbg_shift rows, descriptor base-offset field
  int split_n;  // 1: a work item is one (m tile, n tile) pair (dense layers with few m tiles); 0: one m tile, all n tiles
  uint32_t idesc_hh, idesc_lh, idesc_hl;  // (A_hi,B_hi) (A_lo,B_hi) (A_hi,B_lo)
  uint32_t idesc_last;                    // EPI_DFTF: descriptor of the last N tile when it holds only last_bins bins
  int last_bins;
  int a_mode;
  int hpb;  // a_mode 1: 64-sample blocks per hop
  // a_mode 2 geometry
  int tiles_w, tiles_h, tw, th, ksize, cblocks, cblk, pad;
  // common epilogue
  long long M_total;
  int N_total;
  const float* bias;
  int relu;
  float* out_f32;           // PLAIN: C[M_total][ldc]
  int ldc;
  __nv_bfloat16* out_hi;    // PLAIN/CONV: split output for the next tensor-core layer
  __nv_bfloat16* out_lo;
  // DFT epilogue
  const void* a_hi_ptr;     // EPI_DFTF: base pointers / row pitch (bytes) of the A operand, for the L2-prefetch warp
  const void* a_lo_ptr;
  long long a_pitch;
  const float* inv2;        // per chunk 2^(-2 s)
  const MelTap* taps;       // [nbins_pad]
  float* melpow;            // [rows][n_mels]
  int R, F, n_mels, nbins_pad;
  // CONV epilogue
  int H, W, Cout, pool;     // H, W: conv output size before pooling
};

// non-template dispatcher (all instantiations live in gemm3.cu)
int run_gemm3(int bn, int swz, int epi, const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi,
              const CUtensorMap& tmB_lo, const Gemm3Params& P, int sm_count, cudaStream_t st);

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

// threads per CTA: TMA warp + MMA warp + 4 epilogue warps (8 for the folded STFT, whose epilogue is not overlapped)
template <int EPI>
struct Gemm3Threads { static constexpr int value = (EPI == EPI_DFTF) ? 352 : 192; };   // DFTF: + 4 epilogue warps + 1 L2-prefetch warp

template <int BN, int SWZ, int EPI>
__global__ void __launch_bounds__(Gemm3Threads<EPI>::value, 1)
gemm3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
             const Gemm3Params P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  using Cfg = Gemm3Cfg<BN, SWZ>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool FOLD = (EPI == EPI_DFTF);
  constexpr int NACC = FOLD ? 1 : 2;            // TMEM accumulator buffers
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
  uint8_t* tail = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                           // [STAGES]
  uint64_t* tmem_full = empty_bar + STAGES;                          // [2]
  uint64_t* tmem_empty = tmem_full + 2;                              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  MelTap* s_taps = reinterpret_cast<MelTap*>(tail + 512);            // EPI_DFT only (<= 12 KB)
  float* s_bias = reinterpret_cast<float*>(tail + 512);              // other epilogues: bias[N_total], zero padded

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
      mbar_init(&tmem_empty[a], FOLD ? 256 : 128);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (EPI == EPI_DFT || EPI == EPI_DFTF) {
    for (int i = threadIdx.x; i < P.nbins_pad; i += blockDim.x) s_taps[i] = P.taps[i];
  } else {
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
      // FOLD: the A operand (folded frames) streams from HBM and is re-read once per N tile; a second iterator runs
      // PF K blocks ahead of the loads and prefetches those boxes into L2
      constexpr int PF = 512 / Cfg::BK;   // 8 blocks of 64 / 16 blocks of 32 taps ahead
      [[maybe_unused]] int pf_item = blockIdx.x, pf_nt = 0, pf_kb = 0;
      [[maybe_unused]] auto pf_step = [&]() {
        if (pf_item < n_items) {
          tma_prefetch_2d(&tmA_hi, pf_kb * Cfg::BK, pf_item * Cfg::BM);
          tma_prefetch_2d(&tmA_lo, pf_kb * Cfg::BK, pf_item * Cfg::BM);
          if (++pf_kb == nkb) {
            pf_kb = 0;
            if (++pf_nt == P.num_n_tiles) { pf_nt = 0; pf_item += gridDim.x; }
          }
        }
      };
      (void)pf_step;
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
            } else if (P.a_mode == 1) {
              const int x = (kb % P.hpb) * Cfg::BK;
              const int y = mt * Cfg::BM + kb / P.hpb;
              tma_load_2d(sa_hi, &tmA_hi, &full_bar[stage], x, y);
              tma_load_2d(sa_lo, &tmA_lo, &full_bar[stage], x, y);
            } else {
              const int tap = kb / P.cblocks;
              const int cb = kb - tap * P.cblocks;
              const int kh = tap / P.ksize, kw = tap - kh * P.ksize;
              tma_load_4d(sa_hi, &tmA_hi, &full_bar[stage], cb * P.cblk, w0 + kw - P.pad, h0 + kh - P.pad, img);
              tma_load_4d(sa_lo, &tmA_lo, &full_bar[stage], cb * P.cblk, w0 + kw - P.pad, h0 + kh - P.pad, img);
            }
            if (FOLD) {   // B2 rows: tile nt holds BN cos rows then BN (-sin) rows, each K/2 wide
              const int hk = nkb >> 1;
              const int bx = (kb < hk ? kb : kb - hk) * Cfg::BK, by = nt * 2 * BN + (kb < hk ? 0 : BN);
              tma_load_2d(sb_hi, &tmB_hi, &full_bar[stage], bx, by);
              tma_load_2d(sb_lo, &tmB_lo, &full_bar[stage], bx, by);
            } else {
              tma_load_2d(sb_hi, &tmB_hi, &full_bar[stage], kb * Cfg::BK, nt * BN);
              tma_load_2d(sb_lo, &tmB_lo, &full_bar[stage], kb * Cfg::BK, nt * BN);
            }
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
          const uint32_t d_tmem0 = tmem_base + static_cast<uint32_t>(acc * BN);
          [[maybe_unused]] const bool last_tile = FOLD && (sub == n_sub - 1) && P.last_bins != BN;
          const uint32_t id_hh = last_tile ? P.idesc_last : P.idesc_hh;
          const uint32_t id_lh = last_tile ? P.idesc_last : P.idesc_lh;
          const uint32_t id_hl = last_tile ? P.idesc_last : P.idesc_hl;
          for (int kb = 0; kb < nkb; ++kb) {
            const int hk = nkb >> 1;
            const uint32_t d_tmem = FOLD ? d_tmem0 + (kb < hk ? 0u : static_cast<uint32_t>(BN)) : d_tmem0;
            const int kb_acc = FOLD ? (kb < hk ? kb : kb - hk) : kb;
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
  } else if (FOLD && warp == 10) {
    // ------------------------------------------------------------------ L2 prefetch warp (folded STFT only)
    // The A operand (folded frames) streams from HBM and is re-read once per N tile.  This warp walks the producer's
    // K-block sequence PFD blocks ahead, paced by the same "stage free" barriers, and pulls the rows of that block into
    // L2 with plain prefetch instructions (LSU path): the TMA unit's row rate is the scarce resource of this kernel.
    constexpr int PFD = 6;
    int stage = 0;
    uint32_t phase = 0;
    int pf_item = blockIdx.x, pf_nt = 0, pf_kb = 0;
    auto issue = [&]() {
      if (pf_item < n_items) {
        const long long row0 = static_cast<long long>(pf_item) * Cfg::BM;
        const char* ph = static_cast<const char*>(P.a_hi_ptr) + row0 * P.a_pitch + static_cast<long long>(pf_kb) * SWZ;
        const char* pl = static_cast<const char*>(P.a_lo_ptr) + row0 * P.a_pitch + static_cast<long long>(pf_kb) * SWZ;
#pragma unroll
        for (int q = 0; q < Cfg::BM / 32; ++q) {
          const long long off = static_cast<long long>(lane + 32 * q) * P.a_pitch;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ph + off));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + off));
        }
        if (++pf_kb == nkb) {
          pf_kb = 0;
          if (++pf_nt == P.num_n_tiles) { pf_nt = 0; pf_item += gridDim.x; }
        }
      }
    };
    for (int i = 0; i < PFD; ++i) issue();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int nt = 0; nt < P.num_n_tiles; ++nt) {
        for (int kb = 0; kb < nkb; ++kb) {
          if (lane == 0) mbar_wait(&empty_bar[stage], phase ^ 1u, 500 + stage);
          __syncwarp();
          issue();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5; 2..9 for the folded STFT)
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter + 32) belong to this warp
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = P.split_n ? item / P.num_n_tiles : item;
      const int nt_begin = P.split_n ? item - mt * P.num_n_tiles : 0;
      const int nt_end = P.split_n ? nt_begin + 1 : P.num_n_tiles;
      // ---- per M-tile state
      [[maybe_unused]] int cur = 0;
      [[maybe_unused]] float acc0 = 0.f, acc1 = 0.f, s2 = 0.f;
      [[maybe_unused]] bool valid = false;
      [[maybe_unused]] long long g = 0;
      if (EPI == EPI_DFT || EPI == EPI_DFTF) {
        g = static_cast<long long>(mt) * Cfg::BM + row;
        const long long c = g / P.R;                    // R = rows per chunk (direct: R incl. junk frames; folded: F)
        const int f = static_cast<int>(g - c * P.R);
        valid = (g < P.M_total) && (f < P.F);
        s2 = valid ? P.inv2[c] : 0.f;
      }
      for (int nt = nt_begin; nt < nt_end; ++nt) {
        mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
        tcgen05_fence_after();
        const uint32_t t_acc = tmem_base + lane_base + static_cast<uint32_t>(acc * BN);

        if (EPI == EPI_DFTF) {
          // Folded STFT: [0, BN) = Re, [BN, 2 BN) = Im of the tile's BN bins.  Two warps per TMEM lane quarter, each
          // taking half of the tile's bins; every mel output is accumulated with atomicAdd onto a zeroed buffer (a
          // filter is narrower than half a tile, so it receives at most two partial sums: 0 + a + b is order independent).
          const int half_id = (warp - 2) >> 2;
          const int nb_tile = (nt == P.num_n_tiles - 1) ? P.last_bins : BN;
          const int hbins = nb_tile >> 1;
          const int b0 = half_id * hbins;
          const MelTap* tile_taps = s_taps + nt * BN;
          float* mrow = P.melpow + g * P.n_mels;
          int mcur = tile_taps[b0].first;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll 1
          for (int c0 = b0; c0 < b0 + hbins; c0 += 16) {
            uint32_t re[16], im[16];
            tmem_ld16(t_acc + c0, re);
            tmem_ld16(t_acc + BN + c0, im);
            tmem_ld_wait();
            float pw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = __uint_as_float(re[j]), b = __uint_as_float(im[j]);
              pw[j] = (a * a + b * b) * s2;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const MelTap tp = tile_taps[c0 + j];
              if (mcur < tp.first) {                 // warp-uniform (taps do not depend on the row), rare: 64 times per row
#pragma unroll 1
                while (mcur < tp.first) {
                  if (valid && a0 != 0.f) atomicAdd(mrow + mcur, a0);
                  a0 = a1;
                  a1 = 0.f;
                  ++mcur;
                }
              }
              a0 = fmaf(tp.w0, pw[j], a0);
              a1 = fmaf(tp.w1, pw[j], a1);
            }
          }
          if (valid && a0 != 0.f && mcur < P.n_mels) atomicAdd(mrow + mcur, a0);
          if (valid && a1 != 0.f && mcur + 1 < P.n_mels) atomicAdd(mrow + mcur + 1, a1);
        } else if (EPI == EPI_PLAIN) {
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
                for (int j = 0; j < 16 && n0 + j < P.N_total; ++j) {
                  const __nv_bfloat16 hi = __float2bfloat16_rn(o[j]);
                  P.out_hi[m * P.ldc + n0 + j] = hi;
                  P.out_lo[m * P.ldc + n0 + j] = __float2bfloat16_rn(o[j] - __bfloat162float(hi));
                }
              }
            }
          }
        } else if (EPI == EPI_DFT) {
          // columns [0, BN/2) = Re, [BN/2, BN) = Im of the same BN/2 bins
          constexpr int HB = BN / 2;
#pragma unroll 1
          for (int c0 = 0; c0 < HB; c0 += 16) {
            uint32_t re[16], im[16];
            tmem_ld16(t_acc + c0, re);
            tmem_ld16(t_acc + HB + c0, im);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const MelTap tp = s_taps[nt * HB + c0 + j];
              const float a = __uint_as_float(re[j]), b = __uint_as_float(im[j]);
              const float pw = (a * a + b * b) * s2;
              while (cur < tp.first) {               // warp-uniform: taps and cur do not depend on the row
                if (valid) P.melpow[g * P.n_mels + cur] = acc0;
                acc0 = acc1;
                acc1 = 0.f;
                ++cur;
              }
              acc0 = fmaf(tp.w0, pw, acc0);
              acc1 = fmaf(tp.w1, pw, acc1);
            }
          }
          if (nt == P.num_n_tiles - 1) {
            while (cur < P.n_mels) {
              if (valid) P.melpow[g * P.n_mels + cur] = acc0;
              acc0 = acc1;
              acc1 = 0.f;
              ++cur;
            }
          }
        } else {
          // EPI_CONV: row = pixel (h_local * tw + w_local) of a th x tw output tile, tw in {8, 16}:
          // the 2x2 pooling partners are lane ^ 1 (w) and lane ^ tw (h) of the same warp.
          const int per_img = P.tiles_w * P.tiles_h;
          const int img = mt / per_img;
          const int r = mt - img * per_img;
          const int h = (r / P.tiles_w) * P.th + row / P.tw;
          const int w = (r % P.tiles_w) * P.tw + row % P.tw;
          const bool inb = (h < P.H) && (w < P.W);
          const int OH = P.H / P.pool, OW = P.W / P.pool;
          const bool writer = inb && (P.pool == 1 || (((h | w) & 1) == 0));
          const size_t opix = (static_cast<size_t>(img) * OH + h / P.pool) * OW + w / P.pool;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
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
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = relu_nan(o[j]);
            }
            if (P.pool == 2) {   // 16 independent shuffles in flight per stage
              float u[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) u[j] = __shfl_xor_sync(0xffffffffu, o[j], 1);
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = max_nan(o[j], u[j]);
#pragma unroll
              for (int j = 0; j < 16; ++j) u[j] = __shfl_xor_sync(0xffffffffu, o[j], P.tw);
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = max_nan(o[j], u[j]);
            }
            if (writer && n0 < P.Cout) {
              __align__(16) __nv_bfloat16 hi[16];
              __align__(16) __nv_bfloat16 lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                hi[j] = __float2bfloat16_rn(o[j]);
                lo[j] = __float2bfloat16_rn(o[j] - __bfloat162float(hi[j]));
              }
              __nv_bfloat16* dh = P.out_hi + opix * P.Cout + n0;
              __nv_bfloat16* dl = P.out_lo + opix * P.Cout + n0;
              if (n0 + 16 <= P.Cout) {
                reinterpret_cast<uint4*>(dh)[0] = reinterpret_cast<const uint4*>(hi)[0];
                reinterpret_cast<uint4*>(dh)[1] = reinterpret_cast<const uint4*>(hi)[1];
                reinterpret_cast<uint4*>(dl)[0] = reinterpret_cast<const uint4*>(lo)[0];
                reinterpret_cast<uint4*>(dl)[1] = reinterpret_cast<const uint4*>(lo)[1];
              } else {
                for (int j = 0; j < 16 && n0 + j < P.Cout; ++j) { dh[j] = hi[j]; dl[j] = lo[j]; }
              }
            }
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
int launch_gemm3(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi,
                 const CUtensorMap& tmB_lo, const Gemm3Params& P, int sm_count, cudaStream_t st) {
  using Cfg = Gemm3Cfg<BN, SWZ>;
  static bool configured = false;
  auto kfn = gemm3_kernel<BN, SWZ, EPI>;
  if (!configured) {
    AVLD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int items = P.split_n ? P.num_m_tiles * P.num_n_tiles : P.num_m_tiles;
  int grid = items < sm_count ? items : sm_count;
  if (grid < 1) return AVLD_OK;
  kfn<<<grid, Gemm3Threads<EPI>::value, Cfg::SMEM_BYTES, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, P);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

#endif  // AVLD_GEMM3_IMPL

}  // namespace avld
