// conv1t.cu -- the encoder's first convolution (3x3, stride 1, one input channel: the fp32 feature image) on tcgen05.
//
// K = 9 is no tensor-core shape, so the operand is built on the fly (im2col, K padded 9 -> 16): per 16 x 8 output tile
//   warps 6..9      builders, one thread per output pixel: nine taps of the fp32 image through the read-only L1 path (the
//                   49 KB image is L1 / L2 resident; borders are predicated to zero = the convolution's padding; the next
//                   tile's taps are in flight while this tile's row is written), bf16 hi / lo split (packed cvt.rn.bf16x2),
//                   one 32-byte row of the A_hi and of the A_lo tile each (SWIZZLE_32B layout, ordinary stores +
//                   fence.proxy.async).  (A TMA box of the fp32 halo raised `illegal instruction` on this part in every
//                   form tried -- rank 3 / 4, 48- / 64-byte rows -- so the halo is not staged.)
//   warp 1          MMA issuer: A_hi x [W_hi | W_lo] as one N = 64 MMA into two column ranges of the accumulator and
//                   A_lo x W_hi as an N = 32 one (same split-precision scheme and same N-concatenation as convh.cu)
//   warps 2..5      epilogue: the two ranges added, bias + ReLU + 2x2 max / average pool (reduce-scatter over the window's four
//                   lanes, as in convh.cu) + bf16 hi / lo split, NHWC stores
// The weights (32 x 16 hi and lo, 2 KB) are split once per CTA.  Shapes this kernel does not take (other kernel sizes,
// strides, more than 32 output channels) stay on the CUDA-core kernels of encoder.cu.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avld {

struct Conv1tParams {
  const float* in;             // [n][H][W] fp32 feature images
  const float* w;              // [32][9] fp32, BatchNorm folded, rows past Cout zero
  const float* bias;           // [32]
  __nv_bfloat16* out_hi;       // [n][OH][OW][32]
  __nv_bfloat16* out_lo;
  int n_tiles, tiles_w, tiles_h;
  int H, W, relu, pool, pool_avg;
  uint32_t idesc64, idesc32;
  int dbg;                     // AVLD_BRINGUP builds: AVLD_CONV1T_DBG bit 0 no MMAs, bit 2 no TMEM loads, bit 3 no N = 32 MMA
};

namespace {
constexpr int kTW = 8, kTH = 16;                  // output tile: 16 rows x 8 columns = 128 pixels = the 128 TMEM lanes
constexpr int kAStage = 2 * 128 * 32;             // A_hi + A_lo tiles, 32-byte rows
constexpr int kAStages = 4;
constexpr int kBN = 32, kAccCols = 2 * kBN;       // [hi*hi + lo*hi | hi*lo]
constexpr int kThreads = 320;
constexpr int kSmem = kAStages * kAStage + 2 * kBN * 32 + 1024 /* barriers, bias */ + 1024 /* alignment */;

// byte offset of 16-byte chunk `c` (0 / 1) of row `r` in a SWIZZLE_32B K-major tile (Swizzle<1,4,3>: bit 4 ^= bit 7)
__device__ __forceinline__ uint32_t sw32(int r, int c) { return static_cast<uint32_t>(r * 32 + ((c ^ ((r >> 2) & 1)) << 4)); }

// nine values -> one 32-byte bf16 hi row and one lo row (K = 16: taps 0..8, then zeros)
__device__ __forceinline__ void put_row(uint8_t* t_hi, uint8_t* t_lo, int r, const float (&v)[9]) {
  uint32_t h[5], l[5];
#pragma unroll
  for (int k = 0; k < 4; ++k) split_bf16x2(v[2 * k], v[2 * k + 1], h[k], l[k]);
  split_bf16x2(v[8], 0.f, h[4], l[4]);
  *reinterpret_cast<uint4*>(t_hi + sw32(r, 0)) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(t_hi + sw32(r, 1)) = make_uint4(h[4], 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(t_lo + sw32(r, 0)) = make_uint4(l[0], l[1], l[2], l[3]);
  *reinterpret_cast<uint4*>(t_lo + sw32(r, 1)) = make_uint4(l[4], 0u, 0u, 0u);
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 2) conv1t_kernel(const Conv1tParams P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;                                             // [kAStages][hi 4 KB | lo 4 KB]
  uint8_t* s_w = s_a + kAStages * kAStage;                         // [W_hi 32 rows | W_lo 32 rows] x 32 B
  uint8_t* tail = s_w + 2 * kBN * 32;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);           // [4]
  uint64_t* a_empty = a_full + 4;                                  // [4]
  uint64_t* tmem_full = a_empty + 4;                               // [2]
  uint64_t* tmem_empty = tmem_full + 2;                            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 512);            // [32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&a_full[s], 4);
      mbar_init(&a_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccCols);
  if (threadIdx.x < kBN) {                        // weights: one output channel per thread, rows 0..31 hi, 32..63 lo
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = P.w[threadIdx.x * 9 + t];
    uint32_t h[5], l[5];
#pragma unroll
    for (int k = 0; k < 4; ++k) split_bf16x2(v[2 * k], v[2 * k + 1], h[k], l[k]);
    split_bf16x2(v[8], 0.f, h[4], l[4]);
    const int r = threadIdx.x;
    *reinterpret_cast<uint4*>(s_w + sw32(r, 0)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(s_w + sw32(r, 1)) = make_uint4(h[4], 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s_w + sw32(kBN + r, 0)) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(s_w + sw32(kBN + r, 1)) = make_uint4(l[4], 0u, 0u, 0u);
    s_bias[threadIdx.x] = P.bias[threadIdx.x];
    fence_proxy_async();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = P.tiles_w * P.tiles_h;

  if (warp == 0) {
    // (idle)
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp walks, one lane issues)
    int as = 0, acc = 0;
    uint32_t aph = 0, acc_phase = 0;
    const uint64_t dw = make_smem_desc(smem_u32(s_w), 32);
    for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u, 520 + acc);
      mbar_wait(&a_full[as], aph, 530 + as);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kAccCols);
      const uint32_t a_hi = smem_u32(s_a + as * kAStage);
      const uint64_t da_hi = make_smem_desc(a_hi, 32), da_lo = make_smem_desc(a_hi + 128 * 32, 32);
      if (elect_one()) {
        if (!(P.dbg & 1)) umma_f16(d_tmem, da_hi, dw, P.idesc64, 0u);          // [hi*hi | hi*lo]
        if (!(P.dbg & 9)) umma_f16(d_tmem, da_lo, dw, P.idesc32, 1u);          // + lo*hi into the first range
        umma_commit(&a_empty[as]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (++as == kAStages) { as = 0; aph ^= 1u; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 6) {
    // ------------------------------------------------------------------ builders: im2col rows, one pixel per thread
    const int p = threadIdx.x - 192;              // 0..127: pixel (p / 8, p % 8) of the tile = TMEM lane p
    const int py = p >> 3, px = p & 7;
    auto fetch = [&](int t, float (&v)[9]) {
      const int img = t / per_img, r = t - img * per_img;
      const int h = (r / P.tiles_w) * kTH + py, w = (r % P.tiles_w) * kTW + px;
      const float* __restrict__ src = P.in + (static_cast<size_t>(img) * P.H + h) * P.W + w;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int hh = h + kh - 1, ww = w + kw - 1;
          v[kh * 3 + kw] = (hh >= 0 && hh < P.H && ww >= 0 && ww < P.W) ? __ldg(src + (kh - 1) * P.W + (kw - 1)) : 0.f;
        }
    };
    int as = 0;
    uint32_t aph = 0;
    float cur[9], nxt[9];
    if (static_cast<int>(blockIdx.x) < P.n_tiles) fetch(blockIdx.x, cur);
    for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
      const int tn = t + gridDim.x;
      if (tn < P.n_tiles) fetch(tn, nxt);         // in flight while this tile's row is written
      mbar_wait(&a_empty[as], aph ^ 1u, 550 + as);
      uint8_t* t_hi = s_a + as * kAStage;
      put_row(t_hi, t_hi + 128 * 32, p, cur);
      fence_proxy_async();                        // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[as]);
      if (++as == kAStages) { as = 0; aph ^= 1u; }
#pragma unroll
      for (int k = 0; k < 9; ++k) cur[k] = nxt[k];
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5), as convh.cu with BN = 32
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int OH = P.H / P.pool, OW = P.W / P.pool;
    for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
      const int img = t / per_img, r = t - img * per_img;
      const int h = (r / P.tiles_w) * kTH + row / kTW, w = (r % P.tiles_w) * kTW + row % kTW;
      const bool inb = (h < P.H) && (w < P.W);
      const size_t opix = (static_cast<size_t>(img) * OH + h / P.pool) * OW + w / P.pool;
      mbar_wait(&tmem_full[acc], acc_phase, 560 + acc);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + lane_base + static_cast<uint32_t>(acc * kAccCols);
#pragma unroll 1
      for (int c0 = 0; c0 < kBN; c0 += 16) {
        uint32_t v[16], v2[16];
        if (!(P.dbg & 4)) {
          tmem_ld16(t_acc + c0, v);
          tmem_ld16(t_acc + kBN + c0, v2);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v2[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        if (P.pool == 2) {
          // reduce-scatter over the window's four lanes (partners lane ^ 1 and lane ^ 8): a lane ends with 4 of the 16
          // channels, pooled; bias and ReLU commute with the max, the average takes them first
          const bool odd_w = (lane & 1) != 0, odd_h = (lane & kTW) != 0;
          const bool avg = P.pool_avg != 0;
          if (avg) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float tv = __uint_as_float(v[j]) + s_bias[c0 + j];
              if (P.relu) tv = relu_nan(tv);
              v[j] = __float_as_uint(tv);
            }
          }
          float keep[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float lo8 = __uint_as_float(v[j]), hi8 = __uint_as_float(v[j + 8]);
            const float mine = odd_w ? hi8 : lo8, other = __shfl_xor_sync(0xffffffffu, odd_w ? lo8 : hi8, 1);
            keep[j] = avg ? mine + other : max_nan(mine, other);
          }
          float q4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float mine = odd_h ? keep[j + 4] : keep[j], other = __shfl_xor_sync(0xffffffffu, odd_h ? keep[j] : keep[j + 4], kTW);
            q4[j] = avg ? (mine + other) * 0.25f : max_nan(mine, other);
          }
          const int cq = c0 + (odd_w ? 8 : 0) + (odd_h ? 4 : 0);
          if (((h | 1) < P.H) && ((w | 1) < P.W)) {
            float o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o4[j] = avg ? q4[j] : q4[j] + s_bias[cq + j];
              if (P.relu && !avg) o4[j] = relu_nan(o4[j]);
            }
            uint32_t hw[2], lw[2];
            split_bf16x2(o4[0], o4[1], hw[0], lw[0]);
            split_bf16x2(o4[2], o4[3], hw[1], lw[1]);
            *reinterpret_cast<uint2*>(P.out_hi + opix * kBN + cq) = make_uint2(hw[0], hw[1]);
            *reinterpret_cast<uint2*>(P.out_lo + opix * kBN + cq) = make_uint2(lw[0], lw[1]);
          }
        } else if (inb) {
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float o0 = __uint_as_float(v[j]) + s_bias[c0 + j], o1 = __uint_as_float(v[j + 1]) + s_bias[c0 + j + 1];
            if (P.relu) { o0 = relu_nan(o0); o1 = relu_nan(o1); }
            split_bf16x2(o0, o1, hw[j >> 1], lw[j >> 1]);
          }
          uint4* dh = reinterpret_cast<uint4*>(P.out_hi + opix * kBN + c0);
          uint4* dl = reinterpret_cast<uint4*>(P.out_lo + opix * kBN + c0);
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
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kAccCols);
#endif
}

bool conv1t_supported(int ksize, int stride, int pad, int c_in, int c_out_pad, int h, int w, int pool) {
  return ksize == 3 && stride == 1 && pad == 1 && c_in == 1 && c_out_pad == kBN && w % kTW == 0 && (pool == 1 || (h % 2 == 0 && w % 2 == 0));
}

// feat: fp32 [n][H][W] (16-byte aligned), w: [32][9], bias: [32]; out: NHWC bf16 hi / lo with 32 channels
int launch_conv1t(avld_ctx* c, const float* feat, const OpDev& L, int n, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t st) {
  Conv1tParams P{};
  P.in = feat;
  P.w = L.w_f32;
  P.bias = L.bias;
  P.out_hi = out_hi;
  P.out_lo = out_lo;
  P.tiles_w = L.in_w / kTW;
  P.tiles_h = (L.in_h + kTH - 1) / kTH;
  P.n_tiles = n * P.tiles_w * P.tiles_h;
  P.H = L.in_h; P.W = L.in_w; P.relu = L.relu; P.pool = L.pool; P.pool_avg = L.pool_avg;
  P.idesc64 = avld_make_idesc(1, 1, 128, 2 * kBN);
  P.idesc32 = avld_make_idesc(1, 1, 128, kBN);
#ifdef AVLD_BRINGUP
  if (const char* e = getenv("AVLD_CONV1T_DBG")) P.dbg = atoi(e);
#endif
  AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(conv1t_kernel), kSmem));
  const int grid = std::min(P.n_tiles, 2 * c->sm_count);
  if (grid < 1) return AVLD_OK;
  conv1t_kernel<<<grid, kThreads, kSmem, st>>>(P);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld
