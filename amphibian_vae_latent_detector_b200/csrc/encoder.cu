// encoder.cu -- L1/E1/E2: the VAE encoder forward to the latent mean (map_detector_core.py:270-300) as a small dataflow
// program over numbered tensors (encoder.py::export_program walks the nn.Module): Conv2d (+ folded BatchNorm) (+ ReLU)
// (+ 2x2 max / average pooling), Linear (+ ReLU), residual add (+ ReLU), per-channel affine (+ ReLU), Max / AvgPool2d,
// global average pool; a time axis split into segments whose latents are averaged (core:292-293), and a feature-map
// latent flattened in NCHW order (core:294-295).
//
//   first conv (C_in = 1)           conv1_kernel / conv_direct_kernel on CUDA cores: fp32 feature image in, NHWC bf16 hi / lo
//                                   activations out (K = 9 is no tensor-core shape)
//   3x3 stride-1 convs              convh.cu: implicit GEMM on tcgen05 with halo reuse
//   other convs (1x1, stride 2 ...) gemm3.cuh: implicit GEMM on tcgen05, one TMA box of the NHWC activation per filter tap
//                                   (zero padding = TMA out-of-bounds fill, stride = the tensor map's element strides)
//   linears                         plain GEMM on tcgen05 over the flattened NHWC activations
//   add / affine / pool / GAP       element-wise kernels on the hi / lo planes (HBM bound, a few per residual block)
// Activations and weights are bf16 hi + bf16 lo pairs (3 MMAs per K step, ~2^-17 relative error).  Channel counts are
// padded to 32 or a multiple of 64 with zero filters, so every tensor has a tcgen05-friendly layout whatever the model's
// widths; pad channels hold exact zeros through every operation.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm3.cuh"

namespace avld {

__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
__global__ void split_f16_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                 size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = src[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}
int launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t st) {
  if (n == 0) return AVLD_OK;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 8));
  split_bf16_kernel<<<grid, 256, 0, st>>>(src, hi, lo, n);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}
int launch_split_f16(const float* src, __half* hi, __half* lo, size_t n, cudaStream_t st) {
  if (n == 0) return AVLD_OK;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 8));
  split_f16_kernel<<<grid, 256, 0, st>>>(src, hi, lo, n);
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

// ------------------------------------------------------------------------------------------------
// direct convolution for the first layer (fp32 single-channel feature image)
// one thread = one (pooled) output pixel, all output channels in groups of 8
// ------------------------------------------------------------------------------------------------
struct DirectConvParams {
  const float* in;          // [n][H][W][Cin] fp32
  const float* w;           // [Cout][k][k][Cin]
  const float* bias;        // [Cout]
  __nv_bfloat16* out_hi;    // [n][OH][OW][Cout]
  __nv_bfloat16* out_lo;
  int H, W, Cin, Cout, k, pad, stride, relu, pool, pool_avg, OH, OW, rows_per_block;
};

__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectConvParams P) {
  extern __shared__ float s_mem[];
  const int in_rows = (P.rows_per_block * P.pool - 1) * P.stride + P.k;
  const int in_cols = (P.OW * P.pool - 1) * P.stride + P.k;
  float* s_in = s_mem;                                           // [in_rows][in_cols][Cin]
  float* s_w = s_in + in_rows * in_cols * P.Cin;                 // [Cout][k*k*Cin]
  float* s_b = s_w + P.Cout * P.k * P.k * P.Cin;                 // [Cout]
  const int img = blockIdx.y;
  const int orow0 = blockIdx.x * P.rows_per_block;               // first (pooled) output row of this block
  const int irow0 = orow0 * P.pool * P.stride - P.pad;
  const float* __restrict__ src = P.in + static_cast<size_t>(img) * P.H * P.W * P.Cin;
  for (int i = threadIdx.x; i < in_rows * in_cols * P.Cin; i += blockDim.x) {
    const int ci = i % P.Cin, cc = (i / P.Cin) % in_cols, rr = i / (P.Cin * in_cols);
    const int h = irow0 + rr, w = cc - P.pad;
    s_in[i] = (h >= 0 && h < P.H && w >= 0 && w < P.W) ? src[(static_cast<size_t>(h) * P.W + w) * P.Cin + ci] : 0.f;
  }
  const int kk = P.k * P.k * P.Cin;
  for (int i = threadIdx.x; i < P.Cout * kk; i += blockDim.x) s_w[i] = P.w[i];
  for (int i = threadIdx.x; i < P.Cout; i += blockDim.x) s_b[i] = P.bias[i];
  __syncthreads();

  const int npix = P.rows_per_block * P.OW;
  for (int pix = threadIdx.x; pix < npix; pix += blockDim.x) {
    const int orl = pix / P.OW, oc = pix - orl * P.OW;
    const int orow = orow0 + orl;
    if (orow >= P.OH) continue;
    const size_t obase = ((static_cast<size_t>(img) * P.OH + orow) * P.OW + oc) * P.Cout;
    for (int co0 = 0; co0 < P.Cout; co0 += 8) {
      float best[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) best[q] = P.pool_avg ? 0.f : -INFINITY;
      for (int ph = 0; ph < P.pool; ++ph) {
        for (int pw = 0; pw < P.pool; ++pw) {
          float a[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = (co0 + q < P.Cout) ? s_b[co0 + q] : 0.f;
          const int r0 = (orl * P.pool + ph) * P.stride, c0 = (oc * P.pool + pw) * P.stride;
          for (int kh = 0; kh < P.k; ++kh)
            for (int kw = 0; kw < P.k; ++kw)
              for (int ci = 0; ci < P.Cin; ++ci) {
                const float v = s_in[((r0 + kh) * in_cols + (c0 + kw)) * P.Cin + ci];
                const int wi = (kh * P.k + kw) * P.Cin + ci;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (co0 + q < P.Cout) a[q] = fmaf(v, s_w[(co0 + q) * kk + wi], a[q]);
              }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (P.pool_avg) best[q] += P.relu ? relu_nan(a[q]) : a[q];     // average of the activated values
            else best[q] = max_nan(best[q], a[q]);
          }
        }
      }
      for (int q = 0; q < 8 && co0 + q < P.Cout; ++q) {
        float t = best[q];
        if (P.pool_avg) t *= 1.0f / static_cast<float>(P.pool * P.pool);
        else if (P.relu) t = relu_nan(t);
        const __nv_bfloat16 h = __float2bfloat16_rn(t);
        P.out_hi[obase + co0 + q] = h;
        P.out_lo[obase + co0 + q] = __float2bfloat16_rn(t - __bfloat162float(h));
      }
    }
  }
}

// Fast path for the usual first layer: C_in == 1, 3x3, stride 1, 'same', C_out % 8 == 0, max pooling or none.
// thread = (pooled output pixel, group of 8 output channels); the 72 filter taps of the group live in
// registers for the whole kernel, the fp32 input strip (+ halo, zero padded) in shared memory; four
// consecutive threads cover 32 channels of one pixel, so a warp stores 512 contiguous bytes of the NHWC
// hi and lo planes.
template <int POOL>
__global__ void __launch_bounds__(256) conv1_kernel(const DirectConvParams P) {
  extern __shared__ float s_in[];                               // [in_rows][in_cols]
  const int in_rows = P.rows_per_block * POOL + 2;
  const int in_cols = P.W + 2;
  const int img = blockIdx.y;
  const int orow0 = blockIdx.x * P.rows_per_block;
  const int irow0 = orow0 * POOL - 1;
  const float* __restrict__ src = P.in + static_cast<size_t>(img) * P.H * P.W;
  if ((P.W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // strip fill with eight independent 16-byte loads in flight per thread (the scalar loop below spent a third of the
    // kernel waiting on one dependent load per iteration: ncu, STS behind LDG)
    const int w4 = P.W >> 2, nvec = in_rows * w4;
    for (int base = threadIdx.x; base < nvec; base += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * blockDim.x;
        const int rr = idx / w4, c4 = idx - rr * w4;
        const int h = irow0 + rr;
        v[u] = (idx < nvec && h >= 0 && h < P.H) ? __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(h) * P.W) + c4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + u * blockDim.x;
        if (idx < nvec) {
          const int rr = idx / w4, c4 = idx - rr * w4;
          float* d = s_in + rr * in_cols + 1 + 4 * c4;
          d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
        }
      }
    }
    for (int rr = threadIdx.x; rr < in_rows; rr += blockDim.x) {      // zero padding left and right
      s_in[rr * in_cols] = 0.f;
      s_in[rr * in_cols + P.W + 1] = 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < in_rows * in_cols; i += blockDim.x) {
      const int rr = i / in_cols, cc = i - rr * in_cols;
      const int h = irow0 + rr, w = cc - 1;
      s_in[i] = (h >= 0 && h < P.H && w >= 0 && w < P.W) ? src[static_cast<size_t>(h) * P.W + w] : 0.f;
    }
  }
  const int groups = P.Cout >> 3;
  const int cg = threadIdx.x % groups;
  const int co0 = cg * 8;
  float wreg[8][9], breg[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    breg[q] = P.bias[co0 + q];
#pragma unroll
    for (int t = 0; t < 9; ++t) wreg[q][t] = P.w[(co0 + q) * 9 + t];
  }
  __syncthreads();
  const int npix = P.rows_per_block * P.OW;
  for (int pix = threadIdx.x / groups; pix < npix; pix += blockDim.x / groups) {
    const int orl = pix / P.OW, oc = pix - orl * P.OW;
    const int orow = orow0 + orl;
    if (orow >= P.OH) break;
    float patch[POOL + 2][POOL + 2];
#pragma unroll
    for (int r = 0; r < POOL + 2; ++r)
#pragma unroll
      for (int cc = 0; cc < POOL + 2; ++cc) patch[r][cc] = s_in[(orl * POOL + r) * in_cols + oc * POOL + cc];
    float best[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) best[q] = -INFINITY;
#pragma unroll
    for (int ph = 0; ph < POOL; ++ph)
#pragma unroll
      for (int pw = 0; pw < POOL; ++pw)
#pragma unroll
        for (int q = 0; q < 8; q += 2) {                  // two output channels per FFMA2
          float2 a = make_float2(breg[q], breg[q + 1]);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float pv = patch[ph + kh][pw + kw];
              a = fma2(make_float2(pv, pv), make_float2(wreg[q][kh * 3 + kw], wreg[q + 1][kh * 3 + kw]), a);
            }
          best[q] = max_nan(best[q], a.x);
          best[q + 1] = max_nan(best[q + 1], a.y);
        }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      float t0 = best[q], t1 = best[q + 1];
      if (P.relu) { t0 = relu_nan(t0); t1 = relu_nan(t1); }
      split_bf16x2(t0, t1, hi[q >> 1], lo[q >> 1]);
    }
    const size_t obase = ((static_cast<size_t>(img) * P.OH + orow) * P.OW + oc) * P.Cout + co0;
    *reinterpret_cast<uint4*>(P.out_hi + obase) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(P.out_lo + obase) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// element-wise operations on bf16 hi / lo planes (value = hi + lo, fp32 arithmetic, re-split)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float hl(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t i) {
  return __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
__device__ __forceinline__ void put_hl(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t i, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// out = a + b (+ ReLU): the residual connection.  8 channels (16 bytes of each plane) per thread.
__global__ void __launch_bounds__(256) ew_add_kernel(const uint4* __restrict__ a_hi, const uint4* __restrict__ a_lo,
                                                     const uint4* __restrict__ b_hi, const uint4* __restrict__ b_lo,
                                                     uint4* __restrict__ o_hi, uint4* __restrict__ o_lo, size_t n8, int relu) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 ah = a_hi[i], al = a_lo[i], bh = b_hi[i], bl = b_lo[i];
    const __nv_bfloat16* pah = reinterpret_cast<const __nv_bfloat16*>(&ah);
    const __nv_bfloat16* pal = reinterpret_cast<const __nv_bfloat16*>(&al);
    const __nv_bfloat16* pbh = reinterpret_cast<const __nv_bfloat16*>(&bh);
    const __nv_bfloat16* pbl = reinterpret_cast<const __nv_bfloat16*>(&bl);
    __align__(16) __nv_bfloat16 oh[8], ol[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = (__bfloat162float(pah[q]) + __bfloat162float(pal[q])) + (__bfloat162float(pbh[q]) + __bfloat162float(pbl[q]));
      if (relu) v = relu_nan(v);
      oh[q] = __float2bfloat16_rn(v);
      ol[q] = __float2bfloat16_rn(v - __bfloat162float(oh[q]));
    }
    o_hi[i] = *reinterpret_cast<const uint4*>(oh);
    o_lo[i] = *reinterpret_cast<const uint4*>(ol);
  }
}

// out = in * scale[c] + shift[c] (+ ReLU): a BatchNorm that could not be folded into a convolution, a lone ReLU
__global__ void __launch_bounds__(256) ew_affine_kernel(const __nv_bfloat16* __restrict__ i_hi, const __nv_bfloat16* __restrict__ i_lo,
                                                        __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo,
                                                        const float* __restrict__ scale, const float* __restrict__ shift, int C,
                                                        size_t n, int relu) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C);
    float v = fmaf(hl(i_hi, i_lo, i), scale[ch], shift[ch]);
    if (relu) v = relu_nan(v);
    put_hl(o_hi, o_lo, i, v);
  }
}

// Max / AvgPool2d(k, stride) without padding on NHWC planes; one thread per output element
__global__ void __launch_bounds__(256) pool_kernel(const __nv_bfloat16* __restrict__ i_hi, const __nv_bfloat16* __restrict__ i_lo,
                                                   __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, int H, int W,
                                                   int C, int k, int stride, int OH, int OW, int avg, size_t n_out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_out; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C);
    size_t r = i / C;
    const int ow = static_cast<int>(r % OW);
    r /= OW;
    const int oh = static_cast<int>(r % OH);
    const size_t img = r / OH;
    float acc = avg ? 0.f : -INFINITY;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const size_t at = ((img * H + static_cast<size_t>(oh * stride + kh)) * W + (ow * stride + kw)) * C + ch;
        const float v = hl(i_hi, i_lo, at);
        acc = avg ? acc + v : max_nan(acc, v);
      }
    if (avg) acc *= 1.0f / static_cast<float>(k * k);
    put_hl(o_hi, o_lo, i, acc);
  }
}

// global average pool: [N][HW][C] -> [N][C]; one warp per (image, 32 channels), lanes = channels
__global__ void __launch_bounds__(256) gap_kernel(const __nv_bfloat16* __restrict__ i_hi, const __nv_bfloat16* __restrict__ i_lo,
                                                  __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, int HW, int C,
                                                  size_t n_out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_out; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % C);
    const size_t img = i / C;
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += hl(i_hi, i_lo, (img * HW + p) * C + ch);
    put_hl(o_hi, o_lo, i, acc * (1.0f / static_cast<float>(HW)));
  }
}

// the latent as fp32 rows: from a vector tensor [N][c_pad] or a feature map [N][H][W][c_pad] in NCHW flatten order
__global__ void __launch_bounds__(256) latent_export_kernel(const __nv_bfloat16* __restrict__ i_hi, const __nv_bfloat16* __restrict__ i_lo,
                                                            float* __restrict__ out, int C, int c_pad, int H, int W, size_t n_out) {
  const int D = C * H * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_out; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t img = i / D;
    const int d = static_cast<int>(i - img * D);
    const int ch = d / (H * W), p = d - ch * (H * W);          // NCHW order of the flattened latent (core:294-295)
    out[i] = hl(i_hi, i_lo, (img * H * W + p) * c_pad + ch);
  }
}

// mu[i] = mean over the chunk's segments (core:292-293)
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ lat, float* __restrict__ mu, int n_seg, int D,
                                                           size_t n_out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_out; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t chunk = i / D;
    const int d = static_cast<int>(i - chunk * D);
    float acc = 0.f;
    for (int s = 0; s < n_seg; ++s) acc += lat[(chunk * n_seg + s) * D + d];
    mu[i] = acc / static_cast<float>(n_seg);
  }
}

static int ew_grid(size_t n, int sm_count) {
  const size_t want = (n + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count) * 16;
  return static_cast<int>(want < cap ? (want ? want : 1) : cap);
}

static int pick_bn(int cout) { return cout <= 64 ? 64 : (cout <= 128 ? 128 : 256); }

int launch_encoder(avld_ctx* c, const float* feat, float* mu, int n, cudaStream_t st) {
  AVLD_CHECK(!c->ops.empty(), AVLD_ERR_STATE, "avld_encoder_load has not been called");
  if (n <= 0) return AVLD_OK;
  const int N = n * c->n_seg;                      // images: the segments of a chunk are consecutive rows of its feature image
  auto hi_of = [&](int t) { return c->d_slot_hi[c->tensors[t].slot]; };
  auto lo_of = [&](int t) { return c->d_slot_lo[c->tensors[t].slot]; };
  const bool direct_out = c->n_seg == 1 && !c->enc_out_nchw && c->ops.back().kind == OP_LINEAR && c->ops.back().dst == c->enc_out;
  for (size_t li = 0; li < c->ops.size(); ++li) {
    const OpDev& L = c->ops[li];
    const TensorDev& T = c->tensors[L.dst];
    const bool last_linear_out = L.kind == OP_LINEAR && L.dst == c->enc_out && li + 1 == c->ops.size();
#ifdef AVLD_BRINGUP
    // tensor-core form of the first convolution (conv1t.cu: im2col rows built in shared memory, K = 9 -> 16): correct
    // (2.3e-5 against torch) but 0.42 ms against 0.20 ms for the CUDA-core kernel below, so it is not in the product build
    if (L.kind == OP_CONV_FIRST && c->conv1_tensor && conv1t_supported(L.ksize, L.stride, L.pad, 1, L.c_out, L.in_h, L.in_w, L.pool)) {
      LaunchScope ls(c, ST_CONV_DIRECT, st);
      AVLD_TRY(launch_conv1t(c, feat, L, N, hi_of(L.dst), lo_of(L.dst), st));
      continue;
    }
#endif
    if (L.kind == OP_CONV_FIRST) {
      DirectConvParams P{};
      P.in = feat;
      P.w = L.w_f32;
      P.bias = L.bias;
      P.out_hi = hi_of(L.dst);
      P.out_lo = lo_of(L.dst);
      P.H = L.in_h; P.W = L.in_w; P.Cin = 1; P.Cout = L.c_out; P.k = L.ksize; P.pad = L.pad; P.stride = L.stride;
      P.relu = L.relu; P.pool = L.pool; P.pool_avg = L.pool_avg; P.OH = L.out_h; P.OW = L.out_w;
      const int groups = P.Cout / 8;
      const bool fast = P.k == 3 && P.pad == 1 && P.stride == 1 && !P.pool_avg && P.Cout % 8 == 0 && groups >= 1 &&
                        256 % groups == 0 && static_cast<size_t>(16 * P.pool + 2) * (P.W + 2) * sizeof(float) <= 48 * 1024;
      if (fast) {
        // taller strips amortise the per-block prologue (72 filter taps per thread, the input strip) once the grid is large anyway
        P.rows_per_block = 16;
        for (int rpb : {48, 32})
          if (L.out_h % rpb == 0 && static_cast<size_t>(rpb * P.pool + 2) * (P.W + 2) * sizeof(float) <= 48 * 1024 &&
              static_cast<long long>(L.out_h / rpb) * N >= 8ll * c->sm_count) {
            P.rows_per_block = rpb;
            break;
          }
        const size_t smem = static_cast<size_t>(P.rows_per_block * P.pool + 2) * (P.W + 2) * sizeof(float);
        dim3 grid((L.out_h + P.rows_per_block - 1) / P.rows_per_block, N);
        LaunchScope ls(c, ST_CONV_DIRECT, st);
        if (P.pool == 2) conv1_kernel<2><<<grid, 256, smem, st>>>(P);
        else conv1_kernel<1><<<grid, 256, smem, st>>>(P);
      } else {
        P.rows_per_block = std::max(1, 256 / L.out_w);
        const int in_rows = (P.rows_per_block * P.pool - 1) * P.stride + P.k, in_cols = (P.OW * P.pool - 1) * P.stride + P.k;
        const size_t smem = (static_cast<size_t>(in_rows) * in_cols + static_cast<size_t>(P.Cout) * P.k * P.k + P.Cout) * sizeof(float);
        AVLD_CHECK(smem <= 48 * 1024, AVLD_ERR_UNSUPPORTED, "first-layer direct convolution tile does not fit shared memory");
        dim3 grid((L.out_h + P.rows_per_block - 1) / P.rows_per_block, N);
        LaunchScope ls(c, ST_CONV_DIRECT, st);
        conv_direct_kernel<<<grid, 256, smem, st>>>(P);
      }
      AVLD_CUDA(cudaGetLastError());
    } else if (L.kind == OP_CONV_HALO) {
      LaunchScope ls(c, ST_CONV_GEMM, st);
      AVLD_TRY(launch_convh(c, L, L.tm_in_hi, L.tm_in_lo, N, hi_of(L.dst), lo_of(L.dst), L.src2 >= 0 ? hi_of(L.src2) : nullptr,
                            L.src2 >= 0 ? lo_of(L.src2) : nullptr, st));
    } else if (L.kind == OP_CONV_GEMM) {
      Gemm3Params P{};
      P.num_m_tiles = N * L.tiles_w * L.tiles_h;
      P.num_n_tiles = (L.c_out + L.bn - 1) / L.bn;
      P.num_k_blocks = L.ksize * L.ksize * L.cblocks;
      P.idesc_hh = P.idesc_lh = P.idesc_hl = avld_make_idesc(1, 1, 128, L.bn);
      P.a_mode = 2;
      P.tiles_w = L.tiles_w; P.tiles_h = L.tiles_h; P.tw = L.tw; P.th = L.th; P.ksize = L.ksize;
      P.cblocks = L.cblocks; P.cblk = L.cblk; P.pad = L.pad; P.stride = L.stride;
      P.M_total = static_cast<long long>(P.num_m_tiles) * 128;
      P.N_total = L.c_out;
      P.bias = L.bias;
      P.relu = L.relu;
      P.out_hi = hi_of(L.dst);
      P.out_lo = lo_of(L.dst);
      P.H = L.conv_h; P.W = L.conv_w; P.Cout = L.c_out; P.pool = L.pool; P.pool_avg = L.pool_avg;
      P.res_hi = L.src2 >= 0 ? hi_of(L.src2) : nullptr;
      P.res_lo = L.src2 >= 0 ? lo_of(L.src2) : nullptr;
      LaunchScope ls(c, ST_CONV_GEMM, st);
      AVLD_TRY(run_gemm3(c, L.bn, L.swz, EPI_CONV, L.tm_in_hi, L.tm_in_lo, L.tm_w_hi, L.tm_w_lo, P, st));
    } else if (L.kind == OP_LINEAR) {
      Gemm3Params P{};
      P.num_m_tiles = (N + 127) / 128;
      P.num_n_tiles = (L.c_out + L.bn - 1) / L.bn;
      P.split_n = 1;   // few m tiles (128 images each): spread (m, n) pairs over the SMs
      P.num_k_blocks = static_cast<int>(L.K / 64);
      P.idesc_hh = P.idesc_lh = P.idesc_hl = avld_make_idesc(1, 1, 128, L.bn);
      P.a_mode = 0;
      P.M_total = N;
      P.bias = L.bias;
      P.relu = L.relu;
      if (last_linear_out) {                    // fp32 latents straight from the epilogue: logical width, no pad columns
        P.N_total = c->latent_dim;
        P.ldc = c->latent_dim;
        P.out_f32 = direct_out ? mu : c->d_lat;
      } else {
        P.N_total = L.c_out;
        P.ldc = L.c_out;
        P.out_hi = hi_of(L.dst);
        P.out_lo = lo_of(L.dst);
      }
      LaunchScope ls(c, ST_DENSE_GEMM, st);
      AVLD_TRY(run_gemm3(c, L.bn, 128, EPI_PLAIN, L.tm_in_hi, L.tm_in_lo, L.tm_w_hi, L.tm_w_lo, P, st));
    } else if (L.kind == OP_ADD) {
      const size_t n8 = T.elems() * N / 8;
      LaunchScope ls(c, ST_ELEMENTWISE, st);
      ew_add_kernel<<<ew_grid(n8, c->sm_count), 256, 0, st>>>(
          reinterpret_cast<const uint4*>(hi_of(L.src)), reinterpret_cast<const uint4*>(lo_of(L.src)),
          reinterpret_cast<const uint4*>(hi_of(L.src2)), reinterpret_cast<const uint4*>(lo_of(L.src2)),
          reinterpret_cast<uint4*>(hi_of(L.dst)), reinterpret_cast<uint4*>(lo_of(L.dst)), n8, L.relu);
      AVLD_CUDA(cudaGetLastError());
    } else if (L.kind == OP_AFFINE) {
      const size_t ne = T.elems() * N;
      LaunchScope ls(c, ST_ELEMENTWISE, st);
      ew_affine_kernel<<<ew_grid(ne, c->sm_count), 256, 0, st>>>(hi_of(L.src), lo_of(L.src), hi_of(L.dst), lo_of(L.dst), L.w_f32,
                                                                 L.bias, T.c_pad, ne, L.relu);
      AVLD_CUDA(cudaGetLastError());
    } else if (L.kind == OP_POOL) {
      const size_t ne = T.elems() * N;
      LaunchScope ls(c, ST_ELEMENTWISE, st);
      pool_kernel<<<ew_grid(ne, c->sm_count), 256, 0, st>>>(hi_of(L.src), lo_of(L.src), hi_of(L.dst), lo_of(L.dst), L.in_h, L.in_w,
                                                            T.c_pad, L.ksize, L.stride, L.out_h, L.out_w, L.pool_avg, ne);
      AVLD_CUDA(cudaGetLastError());
    } else if (L.kind == OP_GAP) {
      const size_t ne = T.elems() * N;
      LaunchScope ls(c, ST_ELEMENTWISE, st);
      gap_kernel<<<ew_grid(ne, c->sm_count), 256, 0, st>>>(hi_of(L.src), lo_of(L.src), hi_of(L.dst), lo_of(L.dst), L.in_h * L.in_w,
                                                           T.c_pad, ne);
      AVLD_CUDA(cudaGetLastError());
    } else {
      AVLD_CHECK(false, AVLD_ERR_STATE, "corrupt encoder program (op kind %d)", L.kind);
    }
  }
  const bool head_linear = c->ops.back().kind == OP_LINEAR && c->ops.back().dst == c->enc_out;
  if (!head_linear) {      // the latent is a pooled vector or a feature map: export it as fp32 rows
    const TensorDev& T = c->tensors[c->enc_out];
    const size_t ne = static_cast<size_t>(N) * c->latent_dim;
    LaunchScope ls(c, ST_ELEMENTWISE, st);
    latent_export_kernel<<<ew_grid(ne, c->sm_count), 256, 0, st>>>(hi_of(c->enc_out), lo_of(c->enc_out), c->n_seg == 1 ? mu : c->d_lat,
                                                                   T.c, T.c_pad, T.h, T.w, ne);
    AVLD_CUDA(cudaGetLastError());
  }
  if (c->n_seg > 1) {
    const size_t ne = static_cast<size_t>(n) * c->latent_dim;
    LaunchScope ls(c, ST_ELEMENTWISE, st);
    segment_mean_kernel<<<ew_grid(ne, c->sm_count), 256, 0, st>>>(c->d_lat, mu, c->n_seg, c->latent_dim, ne);
    AVLD_CUDA(cudaGetLastError());
  }
  return AVLD_OK;
}

}  // namespace avld

using namespace avld;

namespace {
int pad_channels(int ch) { return ch <= 32 ? 32 : (ch + 63) / 64 * 64; }

// host [rows][cols] -> device bf16 hi / lo + fp32 copy helpers
int upload_f32(float** dst, const std::vector<float>& v) {
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), std::max<size_t>(v.size(), 1) * sizeof(float)));
  if (!v.empty()) AVLD_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return AVLD_OK;
}
int upload_split(OpDev& L, const std::vector<float>& w) {
  float* tmp = nullptr;
  AVLD_TRY(upload_f32(&tmp, w));
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&L.w_hi), w.size() * 2));
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&L.w_lo), w.size() * 2));
  const int r = launch_split_bf16(tmp, L.w_hi, L.w_lo, w.size(), nullptr);
  cudaDeviceSynchronize();
  cudaFree(tmp);
  return r;
}
}  // namespace

extern "C" int avld_encoder_load_program(avld_ctx* c, const avld_op* ops, int32_t n_ops, int32_t out_tensor, int32_t out_is_map,
                                         int32_t seg_frames, int32_t n_seg) {
  AVLD_ENTER(c);
  AVLD_CHECK(ops && n_ops > 0, AVLD_ERR_INVALID, "NULL / empty encoder program");
  AVLD_CHECK(c->ops.empty(), AVLD_ERR_STATE, "an encoder is already loaded into this context");
  AVLD_CHECK(n_seg >= 1 && seg_frames >= 1 && seg_frames * n_seg == c->T, AVLD_ERR_INVALID,
             "segments: %d x %d frames do not tile target_frames = %d", n_seg, seg_frames, c->T);
  const long long images = static_cast<long long>(c->max_batch) * n_seg;
  AVLD_CHECK(images < (1ll << 30), AVLD_ERR_UNSUPPORTED, "max_batch x segments too large");
  std::vector<TensorDev> tens(1);
  tens[0].c = 1; tens[0].h = seg_frames; tens[0].w = c->M; tens[0].c_pad = 1;
  auto tensor_at = [&](int id) -> TensorDev* { return (id >= 0 && id < static_cast<int>(tens.size()) && tens[id].h > 0) ? &tens[id] : nullptr; };
  auto define = [&](int id, int ch, int h, int w, bool vec) -> int {
    AVLD_CHECK(id >= 1 && id < 4096, AVLD_ERR_INVALID, "tensor id %d out of range", id);
    if (id >= static_cast<int>(tens.size())) tens.resize(id + 1);
    AVLD_CHECK(tens[id].h == 0, AVLD_ERR_INVALID, "tensor %d is written twice", id);
    tens[id].c = ch; tens[id].h = h; tens[id].w = w;
    tens[id].c_pad = vec ? (ch + 63) / 64 * 64 : pad_channels(ch);
    return AVLD_OK;
  };
  std::vector<OpDev> out;
  // a failure half way frees what was uploaded so far
  struct Guard {
    avld_ctx* c;
    std::vector<OpDev>* o;
    bool ok = false;
    ~Guard() {
      if (ok) return;
      for (OpDev& l : *o) {
        cudaFree(l.w_f32); cudaFree(l.bias); cudaFree(l.w_hi); cudaFree(l.w_lo);
      }
      for (auto* p : c->d_slot_hi) cudaFree(p);
      for (auto* p : c->d_slot_lo) cudaFree(p);
      c->d_slot_hi.clear();
      c->d_slot_lo.clear();
      if (c->d_lat) { cudaFree(c->d_lat); c->d_lat = nullptr; }
      if (c->d_mu) { cudaFree(c->d_mu); c->d_mu = nullptr; }
    }
  } guard{c, &out};

  for (int i = 0; i < n_ops; ++i) {
    const avld_op& s = ops[i];
    OpDev L{};
    L.src = s.in0; L.src2 = s.in1; L.dst = s.out;
    L.ksize = s.ksize; L.stride = s.stride; L.pad = s.pad; L.relu = s.relu != 0;
    AVLD_CHECK(tensor_at(s.in0) != nullptr, AVLD_ERR_INVALID, "op %d reads tensor %d before it is written", i, s.in0);
    const TensorDev in_copy = *tensor_at(s.in0);      // by value: define() below may grow `tens` and move its elements
    const TensorDev* in = &in_copy;
    if (s.kind == AVLD_OP_CONV) {
      AVLD_CHECK(s.weight && s.bias, AVLD_ERR_INVALID, "op %d: NULL weights", i);
      AVLD_CHECK(in->w > 1 || in->h > 1 || s.ksize == 1, AVLD_ERR_INVALID, "op %d: convolution on a vector", i);
      AVLD_CHECK(s.c_in == in->c && s.in_h == in->h && s.in_w == in->w, AVLD_ERR_INVALID,
                 "op %d: expects input %dx%dx%d but tensor %d is %dx%dx%d", i, s.in_h, s.in_w, s.c_in, s.in0, in->h, in->w, in->c);
      AVLD_CHECK(s.ksize >= 1 && s.ksize <= 7 && (s.stride == 1 || s.stride == 2) && s.pad >= 0 && s.pad < s.ksize, AVLD_ERR_UNSUPPORTED,
                 "op %d: convolution k=%d stride=%d pad=%d is not implemented", i, s.ksize, s.stride, s.pad);
      AVLD_CHECK(s.pool == 0 || s.pool == 1 || s.pool == 2, AVLD_ERR_INVALID, "op %d: pool must be 0 (none), 1 (max) or 2 (average)", i);
      const int ch = (in->h + 2 * s.pad - s.ksize) / s.stride + 1, cw = (in->w + 2 * s.pad - s.ksize) / s.stride + 1;
      AVLD_CHECK(ch >= 1 && cw >= 1, AVLD_ERR_INVALID, "op %d: empty convolution output", i);
      L.pool = s.pool ? 2 : 1;
      L.pool_avg = s.pool == 2;
      AVLD_CHECK(L.pool == 1 || (ch % 2 == 0 && cw % 2 == 0), AVLD_ERR_UNSUPPORTED, "op %d: 2x2 pooling of an odd map", i);
      L.in_h = in->h; L.in_w = in->w; L.conv_h = ch; L.conv_w = cw; L.out_h = ch / L.pool; L.out_w = cw / L.pool;
      AVLD_TRY(define(s.out, s.c_out, L.out_h, L.out_w, false));
      if (s.in1 >= 0) {            // fused residual add: out = act(conv(in0) + bias + in1)
        const TensorDev* res = tensor_at(s.in1);
        AVLD_CHECK(res != nullptr && s.in1 != 0 && s.in1 != s.out, AVLD_ERR_INVALID, "op %d: residual tensor %d is undefined", i, s.in1);
        AVLD_CHECK(s.in0 != 0 && s.pool == 0, AVLD_ERR_UNSUPPORTED, "op %d: a residual input needs a tensor-core convolution without pooling", i);
        AVLD_CHECK(res->c == s.c_out && res->h == L.out_h && res->w == L.out_w && res->c_pad == tens[s.out].c_pad, AVLD_ERR_INVALID,
                   "op %d: residual tensor %d has another shape than the convolution's output", i, s.in1);
      } else {
        L.src2 = -1;
      }
      const int co_pad = tens[s.out].c_pad, ci_pad = s.in0 == 0 ? 1 : in->c_pad;
      L.c_in = ci_pad; L.c_out = co_pad;
      L.K = static_cast<int64_t>(s.ksize) * s.ksize * ci_pad;
      std::vector<float> w(static_cast<size_t>(co_pad) * L.K, 0.f), b(co_pad, 0.f);
      for (int o = 0; o < s.c_out; ++o) {
        b[o] = s.bias[o];
        for (int t = 0; t < s.ksize * s.ksize; ++t)
          for (int ci = 0; ci < s.c_in; ++ci)
            w[(static_cast<size_t>(o) * s.ksize * s.ksize + t) * ci_pad + ci] = s.weight[(static_cast<size_t>(o) * s.ksize * s.ksize + t) * s.c_in + ci];
      }
      AVLD_TRY(upload_f32(&L.bias, b));
      if (s.in0 == 0) {
        AVLD_CHECK(s.c_in == 1, AVLD_ERR_UNSUPPORTED, "op %d: the feature image has one channel", i);
        L.kind = OP_CONV_FIRST;
        AVLD_TRY(upload_f32(&L.w_f32, w));
      } else {
        L.cblk = ci_pad % 64 == 0 ? 64 : 32;
        L.swz = L.cblk * 2;
        L.cblocks = ci_pad / L.cblk;
        L.bn = pick_bn(co_pad);
        AVLD_CHECK(co_pad <= 2048, AVLD_ERR_UNSUPPORTED, "op %d: C_out > 2048", i);
        // output tile: 128 pixels = th rows x tw columns (powers of two); the shape that wastes the fewest out-of-image
        // pixels wins (they are zero-filled on load and masked on store); fused pooling pairs lanes, so it needs tw >= 2
        long long best_area = -1;
        for (int tw : {16, 8, 4, 2, 1}) {
          if (tw == 1 && L.pool != 1) continue;
          const int th = 128 / tw;
          const long long area = static_cast<long long>((cw + tw - 1) / tw) * tw * ((ch + th - 1) / th) * th;
          if (best_area < 0 || area < best_area) { best_area = area; L.tw = tw; }
        }
        L.th = 128 / L.tw;
        L.tiles_w = (cw + L.tw - 1) / L.tw;
        L.tiles_h = (ch + L.th - 1) / L.th;
        AVLD_TRY(upload_split(L, w));
        AVLD_TRY(encode_tmap_2d(&L.tm_w_hi, L.w_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.K, co_pad, L.K * 2, L.cblk, L.bn, L.swz));
        AVLD_TRY(encode_tmap_2d(&L.tm_w_lo, L.w_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.K, co_pad, L.K * 2, L.cblk, L.bn, L.swz));
        L.kind = OP_CONV_GEMM;
        if (s.stride == 1 && s.pad == 1 && convh_supported(ci_pad, co_pad, s.ksize, in->w)) L.kind = OP_CONV_HALO;
      }
    } else if (s.kind == AVLD_OP_LINEAR) {
      AVLD_CHECK(s.weight && s.bias, AVLD_ERR_INVALID, "op %d: NULL weights", i);
      AVLD_CHECK(s.in0 != 0, AVLD_ERR_UNSUPPORTED, "op %d: a linear layer cannot read the feature image directly", i);
      const int pix = in->h * in->w;
      AVLD_CHECK(s.c_in == pix * in->c, AVLD_ERR_INVALID, "op %d: linear expects %d inputs, tensor %d has %d", i, s.c_in, s.in0, pix * in->c);
      L.kind = OP_LINEAR;
      L.K = static_cast<int64_t>(pix) * in->c_pad;
      AVLD_CHECK(L.K % 64 == 0, AVLD_ERR_UNSUPPORTED, "op %d: padded linear input width %lld is not a multiple of 64", i, static_cast<long long>(L.K));
      AVLD_TRY(define(s.out, s.c_out, 1, 1, true));
      const int co_pad = tens[s.out].c_pad;
      AVLD_CHECK(co_pad <= 4096, AVLD_ERR_UNSUPPORTED, "op %d: linear out_features > 4096", i);
      L.c_in = static_cast<int>(L.K); L.c_out = co_pad;
      // (m, n)-tile work items: 64-column tiles leave most SMs idle on a full pass when the layer is narrow (1024 rows x 512
      // columns = 64 CTAs, each paced by its own TMA row rate); 32-column tiles put twice as many on the machine
      L.bn = ((images + 127) / 128) * (co_pad / 64) < c->sm_count ? 32 : 64;
      L.swz = 128;
      std::vector<float> w(static_cast<size_t>(co_pad) * L.K, 0.f), b(co_pad, 0.f);
      for (int o = 0; o < s.c_out; ++o) {
        b[o] = s.bias[o];
        for (int p = 0; p < pix; ++p)
          for (int ci = 0; ci < in->c; ++ci)
            w[static_cast<size_t>(o) * L.K + static_cast<size_t>(p) * in->c_pad + ci] = s.weight[static_cast<size_t>(o) * s.c_in + static_cast<size_t>(p) * in->c + ci];
      }
      AVLD_TRY(upload_f32(&L.bias, b));
      AVLD_TRY(upload_split(L, w));
      AVLD_TRY(encode_tmap_2d(&L.tm_w_hi, L.w_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.K, co_pad, L.K * 2, 64, L.bn, 128));
      AVLD_TRY(encode_tmap_2d(&L.tm_w_lo, L.w_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.K, co_pad, L.K * 2, 64, L.bn, 128));
    } else if (s.kind == AVLD_OP_ADD) {
      AVLD_CHECK(tensor_at(s.in1) != nullptr && s.in0 != 0 && s.in1 != 0, AVLD_ERR_INVALID, "op %d: add reads an undefined tensor", i);
      const TensorDev in2_copy = *tensor_at(s.in1);
      const TensorDev* in2 = &in2_copy;
      AVLD_CHECK(in2->c == in->c && in2->h == in->h && in2->w == in->w && in2->c_pad == in->c_pad, AVLD_ERR_INVALID, "op %d: add of different shapes", i);
      L.kind = OP_ADD;
      AVLD_TRY(define(s.out, in->c, in->h, in->w, false));
      tens[s.out].c_pad = in->c_pad;
    } else if (s.kind == AVLD_OP_AFFINE) {
      AVLD_CHECK(s.weight && s.bias && s.in0 != 0 && s.c_in == in->c, AVLD_ERR_INVALID, "op %d: bad affine", i);
      L.kind = OP_AFFINE;
      AVLD_TRY(define(s.out, in->c, in->h, in->w, false));
      tens[s.out].c_pad = in->c_pad;
      std::vector<float> sc(in->c_pad, 0.f), sh(in->c_pad, 0.f);      // pad channels: 0 * 0 + 0
      for (int ch = 0; ch < in->c; ++ch) { sc[ch] = s.weight[ch]; sh[ch] = s.bias[ch]; }
      AVLD_TRY(upload_f32(&L.w_f32, sc));
      AVLD_TRY(upload_f32(&L.bias, sh));
    } else if (s.kind == AVLD_OP_POOL) {
      AVLD_CHECK(s.in0 != 0 && (s.pool == 1 || s.pool == 2), AVLD_ERR_INVALID, "op %d: bad pool", i);
      L.pool_avg = s.pool == 2;
      L.in_h = in->h; L.in_w = in->w;
      if (s.ksize == 0) {                       // global average
        AVLD_CHECK(s.pool == 2, AVLD_ERR_UNSUPPORTED, "op %d: global max pooling is not implemented", i);
        L.kind = OP_GAP;
        L.out_h = L.out_w = 1;
      } else {
        AVLD_CHECK(s.ksize >= 1 && s.stride >= 1 && s.ksize <= in->h && s.ksize <= in->w, AVLD_ERR_INVALID, "op %d: bad pool window", i);
        L.kind = OP_POOL;
        L.out_h = (in->h - s.ksize) / s.stride + 1;
        L.out_w = (in->w - s.ksize) / s.stride + 1;
      }
      AVLD_TRY(define(s.out, in->c, L.out_h, L.out_w, false));
      tens[s.out].c_pad = in->c_pad;
    } else {
      AVLD_CHECK(false, AVLD_ERR_INVALID, "op %d: unknown kind %d", i, s.kind);
    }
    out.push_back(L);
  }
  TensorDev* lat = tensor_at(out_tensor);
  AVLD_CHECK(lat != nullptr && out_tensor != 0, AVLD_ERR_INVALID, "latent tensor %d is never written", out_tensor);
  AVLD_CHECK(out_is_map || (lat->h == 1 && lat->w == 1), AVLD_ERR_INVALID, "the latent tensor is a feature map: pass out_is_map");
  const int latent_dim = lat->c * lat->h * lat->w;

  // ---- activation slots: a tensor lives from its producer to its last consumer; slots are reused greedily
  std::vector<int> last_use(tens.size(), -1);
  for (size_t i = 0; i < out.size(); ++i) {
    last_use[out[i].src] = static_cast<int>(i);
    if (out[i].src2 >= 0) last_use[out[i].src2] = static_cast<int>(i);
  }
  last_use[out_tensor] = static_cast<int>(out.size());
  std::vector<int> slot_free_at;                     // per slot: index of the op after which it is free again
  size_t slot_elems = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    TensorDev& T = tens[out[i].dst];
    slot_elems = std::max(slot_elems, T.elems());
    int pick = -1;
    for (size_t s2 = 0; s2 < slot_free_at.size(); ++s2)
      if (slot_free_at[s2] < static_cast<int>(i)) { pick = static_cast<int>(s2); break; }
    if (pick < 0) {
      pick = static_cast<int>(slot_free_at.size());
      slot_free_at.push_back(0);
    }
    T.slot = pick;
    slot_free_at[pick] = std::max(last_use[out[i].dst], static_cast<int>(i));
  }
  AVLD_CHECK(slot_free_at.size() <= 16, AVLD_ERR_UNSUPPORTED, "the encoder keeps more than 16 activations alive");
  const size_t total = slot_elems * static_cast<size_t>(images) + 128 * 256;
  c->d_slot_hi.assign(slot_free_at.size(), nullptr);
  c->d_slot_lo.assign(slot_free_at.size(), nullptr);
  for (size_t s2 = 0; s2 < slot_free_at.size(); ++s2) {
    AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_slot_hi[s2]), total * 2));
    AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_slot_lo[s2]), total * 2));
    AVLD_CUDA(cudaMemset(c->d_slot_hi[s2], 0, total * 2));
    AVLD_CUDA(cudaMemset(c->d_slot_lo[s2], 0, total * 2));
  }
  // ---- input tensor maps
  for (OpDev& L : out) {
    if (L.kind != OP_CONV_HALO && L.kind != OP_CONV_GEMM && L.kind != OP_LINEAR) continue;
    const TensorDev& in = tens[L.src];
    const __nv_bfloat16* in_hi = c->d_slot_hi[in.slot];
    const __nv_bfloat16* in_lo = c->d_slot_lo[in.slot];
    if (L.kind == OP_CONV_HALO) {
      AVLD_TRY(convh_encode_input_map(&L.tm_in_hi, in_hi, static_cast<int>(images), L.in_h, L.in_w, L.c_in, L.cblk));
      AVLD_TRY(convh_encode_input_map(&L.tm_in_lo, in_lo, static_cast<int>(images), L.in_h, L.in_w, L.c_in, L.cblk));
    } else if (L.kind == OP_CONV_GEMM) {
      const uint64_t dims[4] = {static_cast<uint64_t>(L.c_in), static_cast<uint64_t>(L.in_w), static_cast<uint64_t>(L.in_h),
                                static_cast<uint64_t>(images)};
      const uint64_t strides[3] = {static_cast<uint64_t>(L.c_in) * 2, static_cast<uint64_t>(L.in_w) * L.c_in * 2,
                                   static_cast<uint64_t>(L.in_h) * L.in_w * L.c_in * 2};
      // a strided convolution reads every stride-th pixel: the box spans tw * stride (th * stride) pixels, traversed with
      // element stride `stride`, and still delivers tw x th pixels
      const uint32_t st = static_cast<uint32_t>(L.stride);
      const uint32_t box[4] = {static_cast<uint32_t>(L.cblk), static_cast<uint32_t>(L.tw) * st, static_cast<uint32_t>(L.th) * st, 1};
      const uint32_t estr[4] = {1, st, st, 1};
      AVLD_CHECK(box[1] <= 256 && box[2] <= 256, AVLD_ERR_UNSUPPORTED, "strided convolution tile exceeds the TMA box limit");
      AVLD_TRY(encode_tmap_4d(&L.tm_in_hi, in_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, strides, box, L.swz, estr));
      AVLD_TRY(encode_tmap_4d(&L.tm_in_lo, in_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, strides, box, L.swz, estr));
    } else {
      const uint64_t rows = static_cast<uint64_t>(images);   // rows past n: stale but finite values, never stored (M_total)
      AVLD_TRY(encode_tmap_2d(&L.tm_in_hi, in_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.K, rows, L.K * 2, 64, 128, 128));
      AVLD_TRY(encode_tmap_2d(&L.tm_in_lo, in_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.K, rows, L.K * 2, 64, 128, 128));
    }
  }
  AVLD_CUDA(cudaDeviceSynchronize());
  c->latent_dim = latent_dim;
  c->n_seg = n_seg;
  c->seg_frames = seg_frames;
  c->enc_out = out_tensor;
  c->enc_out_nchw = out_is_map ? 1 : 0;
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_mu), static_cast<size_t>(c->max_batch) * latent_dim * sizeof(float)));
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_lat), static_cast<size_t>(images) * latent_dim * sizeof(float)));
  c->tensors = std::move(tens);
  c->ops = std::move(out);
  guard.ok = true;
  return AVLD_OK;
}

// the chain form of the first ABI revision: layer i reads layer i - 1
extern "C" int avld_encoder_load(avld_ctx* c, const avld_layer* layers, int32_t n_layers) {
  AVLD_CHECK(c && layers && n_layers > 0, AVLD_ERR_INVALID, "NULL / empty layer list");
  std::vector<avld_op> ops(n_layers);
  for (int i = 0; i < n_layers; ++i) {
    const avld_layer& s = layers[i];
    avld_op& o = ops[i];
    o = avld_op{};
    o.kind = s.kind == 0 ? AVLD_OP_CONV : AVLD_OP_LINEAR;
    o.in0 = i; o.in1 = -1; o.out = i + 1;
    o.c_in = s.c_in; o.c_out = s.c_out; o.ksize = s.ksize; o.stride = s.stride; o.pad = s.pad; o.relu = s.relu;
    o.pool = s.pool > 1 ? 1 : 0;
    o.in_h = s.in_h; o.in_w = s.in_w;
    o.weight = s.weight; o.bias = s.bias;
    AVLD_CHECK(s.kind == 0 || s.kind == 1, AVLD_ERR_INVALID, "layer %d: unknown kind %d", i, s.kind);
  }
  return avld_encoder_load_program(c, ops.data(), n_layers, n_layers, 0, c->T, 1);
}

extern "C" int avld_encoder_forward(avld_ctx* c, const float* feat, float* mu, int64_t n, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(feat && mu, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0, AVLD_ERR_INVALID, "negative n");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t i = 0; i < n; i += c->max_batch) {
    const int m = static_cast<int>(n - i < c->max_batch ? n - i : c->max_batch);
    AVLD_TRY(launch_encoder(c, feat + i * c->T * c->M, mu + i * c->latent_dim, m, st));
  }
  return AVLD_OK;
}
