// rms.cu -- R1/R2: bit-exact batched rms_normalize (00_normalize_dataset_rms.py:29-38) and the per-chunk
// parameters of the STFT operand (power-of-two scale, normalised PCM_16 integers), fused in one kernel:
// one CTA per chunk.
//
// Bit-exactness: np.mean(y**2) on contiguous float32 walks numpy's pairwise-summation tree
// (8 interleaved accumulators per <=128-element leaf, halves rounded down to a multiple of 8).
// The host builds that tree once per chunk length (ctx.cu::pairwise_plan); the kernel evaluates the
// leaves with the same accumulator order and combines them level by level, every operation an
// explicitly rounded __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn (no FMA contraction), following
// numpy-2 float32 scalar semantics for `rms + eps` and `target / (...)`.
//
// Memory: phase 1 streams the chunk once (4*L bytes), phase 2 re-reads it and writes either y (4*L bytes,
// avld_rms_normalize) or the normalised PCM_16 integers (2*L bytes, feature passes).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "sample.cuh"

namespace avld {

struct PrepParams {
  const float* x;
  const int16_t* x16;     // alternative input: PCM_16 samples, decoded as s / 32768 (librosa.load of a 16-bit WAV)
  float* y;               // nullable
  float4* chunk_par;      // nullable (feature passes): per chunk (scale, pow2, scaled flag, -) for fold3_kernel
  float* inv2;
  uint16_t* q16;          // nullable: normalised, PCM_16-rounded samples + 32768 for fold3_kernel (quantize passes)
  uint8_t* ok;            // nullable
  float* rms;             // nullable
  const int32_t* leaf_off;
  const int32_t* leaf_len;
  const PairNode* nodes;
  const int32_t* level_start;
  int n_leaves, n_nodes, n_levels;
  int L, n_fft, hop;
  float target_rms, rms_min, eps;
  int scalar_f64;         // numpy 1.x scalar rules: gate, `rms + eps` and `target / (...)` in float64 (the three below)
  double target_d, rms_min_d, eps_d;
  int normalize, quantize;
  int dft_scale_log2;
  int headroom_log2;      // max |scaled sample| < 2^(headroom_log2 + 1): 13 (sums of up to eight samples stay in fp16 range)
  int n;
};

// sample loaders: float32 chunk or PCM_16 chunk (exact: |s| < 2^15, scale 2^-15)
struct LoadF32 {
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return p[i]; }
  __device__ __forceinline__ float4 load4(int i) const { return *reinterpret_cast<const float4*>(p + i); }   // i % 4 == 0, aligned base
};
struct LoadPcm16 {
  const int16_t* p;
  __device__ __forceinline__ float operator()(int i) const { return static_cast<float>(p[i]) * (1.0f / 32768.0f); }
  __device__ __forceinline__ float4 load4(int i) const {
    const uint2 u = *reinterpret_cast<const uint2*>(p + i);
    const float k = 1.0f / 32768.0f;
    return make_float4(static_cast<float>(static_cast<int16_t>(u.x & 0xffffu)) * k, static_cast<float>(static_cast<int16_t>(u.x >> 16)) * k,
                       static_cast<float>(static_cast<int16_t>(u.y & 0xffffu)) * k, static_cast<float>(static_cast<int16_t>(u.y >> 16)) * k);
  }
};

template <bool kWide, typename Load>
__device__ __forceinline__ void prep_body(const PrepParams& P, const Load xc, float* s_val, int c);

// One CTA per chunk (a persistent one-CTA-per-SM variant that keeps the second read in L2 measured slower: 89 vs 60 ms
// per 100k chunks -- too little memory parallelism per SM).
__global__ void __launch_bounds__(512) prep_kernel(const PrepParams P) {
  extern __shared__ float s_val[];            // [n_leaves + n_nodes]
  const int c = blockIdx.x;
  const size_t base = static_cast<size_t>(c) * P.L;
  if (P.x16 != nullptr) prep_body<false>(P, LoadPcm16{P.x16 + base}, s_val, c);
  else prep_body<false>(P, LoadF32{P.x + base}, s_val, c);
}

// The same with one CTA of 1024 threads per chunk and therefore one chunk per SM: 148 chunks (85 MB of float32 audio) are in
// flight instead of 592, so the second sweep finds its chunk in the 126 MB L2 and DRAM sees every sample once.  What the
// single resident CTA loses in overlap it gets back from wider loads: two threads per leaf, each owning four of numpy's
// eight interleaved accumulators and loading eight 16-byte rows before the ordered adds (128 KB in flight per SM).
// Regular leaf plans only (every leaf a multiple of 8 samples at a multiple-of-8 offset: the 3 s and 5 s chunk lengths).
__global__ void __launch_bounds__(1024, 1) prep_wide_kernel(const PrepParams P) {
  extern __shared__ float s_val[];
  const int c = blockIdx.x;
  const size_t base = static_cast<size_t>(c) * P.L;
  if (P.x16 != nullptr) prep_body<true>(P, LoadPcm16{P.x16 + base}, s_val, c);
  else prep_body<true>(P, LoadF32{P.x + base}, s_val, c);
}

template <bool kWide, typename Load>
__device__ __forceinline__ void prep_body(const PrepParams& P, const Load xc, float* s_val, int c) {
  __shared__ float s_red[32];
  __shared__ float s_scale, s_pow2;
  __shared__ int s_scaled;
  const int tid = threadIdx.x;

  // ---------------------------------------------------------------- phase 1: leaves
  float mx = 0.f;
  if constexpr (kWide) {
    const int h4 = (tid & 1) * 4;                 // this thread's accumulators: r[h4 .. h4 + 3] of the leaf's eight
    for (int leaf = tid >> 1; leaf < P.n_leaves; leaf += blockDim.x >> 1) {
      const int off = P.leaf_off[leaf] + h4, rows = P.leaf_len[leaf] >> 3;     // 1 .. 16 rows of eight samples
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (u < rows) ? xc.load4(off + 8 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
      float r0 = __fmul_rn(v[0].x, v[0].x), r1 = __fmul_rn(v[0].y, v[0].y), r2 = __fmul_rn(v[0].z, v[0].z), r3 = __fmul_rn(v[0].w, v[0].w);
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[0].x), fabsf(v[0].y)), fmaxf(fabsf(v[0].z), fabsf(v[0].w))));
#pragma unroll
      for (int u = 1; u < 8; ++u)
        if (u < rows) {
          r0 = __fadd_rn(r0, __fmul_rn(v[u].x, v[u].x)); r1 = __fadd_rn(r1, __fmul_rn(v[u].y, v[u].y));
          r2 = __fadd_rn(r2, __fmul_rn(v[u].z, v[u].z)); r3 = __fadd_rn(r3, __fmul_rn(v[u].w, v[u].w));
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
        }
      for (int u0 = 8; u0 < rows; u0 += 4) {       // longer leaves (5 s chunks: 14 / 15 rows): four more rows at a time
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (u0 + u < rows) ? xc.load4(off + 8 * (u0 + u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u0 + u < rows) {
            r0 = __fadd_rn(r0, __fmul_rn(v[u].x, v[u].x)); r1 = __fadd_rn(r1, __fmul_rn(v[u].y, v[u].y));
            r2 = __fadd_rn(r2, __fmul_rn(v[u].z, v[u].z)); r3 = __fadd_rn(r3, __fmul_rn(v[u].w, v[u].w));
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
          }
      }
      // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)): each half locally, the halves across the thread pair
      float r = __fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3));
      r = __fadd_rn(r, __shfl_xor_sync(3u << (tid & 30), r, 1));      // pair mask: other lanes may have left the loop
      if (h4 == 0) s_val[leaf] = r;
    }
  } else {
  const int j = tid & 7;
  const unsigned gmask = 0xFFu << (8 * ((tid & 31) >> 3));
  for (int leaf = tid >> 3; leaf < P.n_leaves; leaf += blockDim.x >> 3) {
    const int off = P.leaf_off[leaf], len = P.leaf_len[leaf];
    float r;
    if (len < 8) {
      r = 0.f;
      if (j == 0) {
        for (int i = 0; i < len; ++i) {
          const float v = xc(off + i);
          mx = fmaxf(mx, fabsf(v));
          r = __fadd_rn(r, __fmul_rn(v, v));
        }
      }
    } else {
      const int n8 = len & ~7;
      // numpy's leaves hold at most 128 elements (64..72 for a 3 s chunk): the first nine 8-element rows are loaded at once
      // (nine independent loads in flight per thread instead of the four an unrolled loop gave), then added in order
      float v9[9];
#pragma unroll
      for (int u = 0; u < 9; ++u) v9[u] = (8 * u < n8) ? xc(off + 8 * u + j) : 0.f;
      float v = v9[0];
      mx = fmaxf(mx, fabsf(v));
      r = __fmul_rn(v, v);
#pragma unroll
      for (int u = 1; u < 9; ++u)
        if (8 * u < n8) {
          mx = fmaxf(mx, fabsf(v9[u]));
          r = __fadd_rn(r, __fmul_rn(v9[u], v9[u]));
        }
#pragma unroll 4
      for (int i = 72; i < n8; i += 8) {
        v = xc(off + i + j);
        mx = fmaxf(mx, fabsf(v));
        r = __fadd_rn(r, __fmul_rn(v, v));
      }
      // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7))
      r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 1, 8));
      r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 2, 8));
      r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 4, 8));
      if (j == 0) {
        for (int i = n8; i < len; ++i) {
          v = xc(off + i);
          mx = fmaxf(mx, fabsf(v));
          r = __fadd_rn(r, __fmul_rn(v, v));
        }
      }
    }
    if (j == 0) s_val[leaf] = r;
  }
  }
  // block max of |x| (only used to pick the power-of-two operand scale; any upper bound is valid)
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();

  // ---------------------------------------------------------------- tree combine, level by level
  for (int lv = 0; lv < P.n_levels; ++lv) {
    const int b = P.level_start[lv], e = P.level_start[lv + 1];
    for (int i = b + tid; i < e; i += blockDim.x) {
      const PairNode nd = P.nodes[i];
      s_val[P.n_leaves + i] = __fadd_rn(s_val[nd.a], s_val[nd.b]);
    }
    __syncthreads();
  }
  if (tid == 0) {
    float m = s_red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
    const float root = s_val[P.n_leaves + P.n_nodes - 1];   // last node = root (a single leaf when n_nodes == 0)
    const float mean = __fdiv_rn(root, static_cast<float>(P.L));
    const float rms = __fsqrt_rn(mean);
    int scaled = 0;
    float scale = 1.0f;
    if (P.scalar_f64) {
      // numpy 1.26.4 (the reference's pin): np.float32 (op) Python float is evaluated in float64; `y * scale` then casts the
      // float64 scalar to the array's float32 -- one rounding (00_normalize_dataset_rms.py:30-36)
      const double rd = static_cast<double>(rms);
      if (P.normalize && !(rd < P.rms_min_d)) {
        scaled = 1;
        scale = __double2float_rn(__ddiv_rn(P.target_d, __dadd_rn(rd, P.eps_d)));
      }
    } else if (P.normalize && !(rms < P.rms_min)) {    // silence gate: `if rms < rms_min: return y, False`
      scaled = 1;
      scale = __fdiv_rn(P.target_rms, __fadd_rn(rms, P.eps));
    }
    float bound = scaled ? fminf(__fmul_rn(m, scale), 1.0f) : m;
    int s = 0;
    if (bound > 0.f && bound < 3.0e38f) {
      s = P.headroom_log2 - ilogbf(bound);
      s = s > 60 ? 60 : (s < -60 ? -60 : s);
    }
    s_scale = scale;
    s_scaled = scaled;
    s_pow2 = ldexpf(1.0f, s);
    // A chunk holding NaN / Inf has a non-finite rms.  The PCM_16 round trip would launder it into digital silence
    // (cvt.rni(NaN) = 0) and a perfectly plausible latent; poisoning the power un-scale factor instead makes the
    // chunk's features, latent and distances NaN, i.e. NO_DETECT with best distance inf -- what the reference's float
    // path (librosa.load -> melspectrogram -> encoder, 09:416-436) yields for such a file.
    if (P.inv2) P.inv2[c] = (rms - rms == 0.0f) ? ldexpf(1.0f, -2 * (s + P.dft_scale_log2)) : NAN;
    if (P.chunk_par) P.chunk_par[c] = make_float4(scale, ldexpf(1.0f, s), scaled ? 1.0f : 0.0f, 0.0f);
    if (P.ok) P.ok[c] = P.normalize ? static_cast<uint8_t>(scaled) : static_cast<uint8_t>(1);
    if (P.rms) P.rms[c] = rms;
  }
  __syncthreads();
  const float scale = s_scale;
  const int scaled = s_scaled;

  // ---------------------------------------------------------------- phase 2a: y (float32)
  if (P.y != nullptr) {
    float* __restrict__ yc = P.y + static_cast<size_t>(c) * P.L;
    if ((P.L & 3) == 0 && P.x16 == nullptr) {
      const float4* x4 = reinterpret_cast<const float4*>(P.x + static_cast<size_t>(c) * P.L);
      float4* y4 = reinterpret_cast<float4*>(yc);
      for (int i = tid; i < (P.L >> 2); i += blockDim.x) {
        float4 v = x4[i];
        v.x = finish_sample(v.x, scale, scaled, P.quantize);
        v.y = finish_sample(v.y, scale, scaled, P.quantize);
        v.z = finish_sample(v.z, scale, scaled, P.quantize);
        v.w = finish_sample(v.w, scale, scaled, P.quantize);
        y4[i] = v;
      }
    } else {
      for (int i = tid; i < P.L; i += blockDim.x) yc[i] = finish_sample(xc(i), scale, scaled, P.quantize);
    }
  }

  // ---------------------------------------------------------------- phase 2c: normalised PCM_16 samples for fold3_kernel
  // (each sample is used by ~6 operand threads; rounding it once here instead of there halves that kernel's instructions)
  if (P.q16 != nullptr) {
    uint16_t* __restrict__ qc = P.q16 + static_cast<size_t>(c) * P.L;
    auto biased = [&](float v) -> uint32_t { return static_cast<uint32_t>(pcm16_of(v, scale, scaled) + 32768); };
    const int n8 = P.L >> 3;
    if (P.x16 == nullptr && (reinterpret_cast<uintptr_t>(P.x) & 15) == 0 && (P.L & 3) == 0) {
      const float4* x4 = reinterpret_cast<const float4*>(P.x + static_cast<size_t>(c) * P.L);
#pragma unroll 4
      for (int i = tid; i < n8; i += blockDim.x) {
        const float4 a = x4[2 * i], b = x4[2 * i + 1];
        uint4 o;
        o.x = biased(a.x) | (biased(a.y) << 16);
        o.y = biased(a.z) | (biased(a.w) << 16);
        o.z = biased(b.x) | (biased(b.y) << 16);
        o.w = biased(b.z) | (biased(b.w) << 16);
        reinterpret_cast<uint4*>(qc)[i] = o;
      }
    } else if (P.x16 != nullptr && (reinterpret_cast<uintptr_t>(P.x16) & 15) == 0) {
      const uint4* x8 = reinterpret_cast<const uint4*>(P.x16 + static_cast<size_t>(c) * P.L);
      auto pair = [&](uint32_t w) -> uint32_t {
        const float lo = static_cast<float>(static_cast<int16_t>(w & 0xffffu)) * (1.0f / 32768.0f);
        const float hi = static_cast<float>(static_cast<int16_t>(w >> 16)) * (1.0f / 32768.0f);
        return biased(lo) | (biased(hi) << 16);
      };
      for (int i = tid; i < n8; i += blockDim.x) {
        const uint4 u = x8[i];
        reinterpret_cast<uint4*>(qc)[i] = make_uint4(pair(u.x), pair(u.y), pair(u.z), pair(u.w));
      }
    } else {
      for (int i = tid; i < n8; i += blockDim.x) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = biased(xc(8 * i + 2 * j)) | (biased(xc(8 * i + 2 * j + 1)) << 16);
        reinterpret_cast<uint4*>(qc)[i] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    for (int i = (n8 << 3) + tid; i < P.L; i += blockDim.x) qc[i] = static_cast<uint16_t>(biased(xc(i)));
  }
}

// ------------------------------------------------------------------------------------------------
// prep_cluster_kernel: one chunk per cluster of eight 512-thread CTAs.  Each CTA owns an eighth of the leaves -- a whole
// subtree of numpy's (perfect) pairwise tree -- reduces it locally with the tree's own pairing (adjacent leaves, then
// adjacent nodes: warp butterflies), and the eight subtree roots (and block maxima) are exchanged through distributed shared
// memory with ONE cluster barrier; every CTA then forms the same root and scale and writes the second sweep of its own
// samples.  Four CTAs of four different chunks share an SM, so sweeps, tree and barrier waits of different chunks overlap
// like in prep_kernel, but only 74 chunks (43 MB) are in flight instead of 592: the second sweep is an L2 hit.
// Needs a regular, perfect plan (the 3 s and 5 s chunk lengths); everything else goes to prep_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kPrepCluster = 8;

__device__ __forceinline__ void st_cluster_f32(float* local_addr, uint32_t cta, float v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.f32 [ra], %2;\n\t}\n" ::"r"(smem_u32(local_addr)),
      "r"(cta), "f"(v)
      : "memory");
}

template <bool PCM>
__global__ void __cluster_dims__(kPrepCluster, 1, 1) __launch_bounds__(512, 4) prep_cluster_kernel(const PrepParams P) {
  __shared__ float s_leaf[512];                 // this CTA's leaf sums
  __shared__ float s_wroot[16], s_wmax[16];
  __shared__ float s_xroot[kPrepCluster], s_xmax[kPrepCluster];   // written by every CTA of the cluster
  __shared__ float s_scale;
  __shared__ int s_scaled;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int c = blockIdx.x / kPrepCluster;
  const int lpc = P.n_leaves / kPrepCluster;    // leaves per CTA: 32 .. 512, a power of two
  const int leaf0 = static_cast<int>(rank) * lpc;
  const size_t base = static_cast<size_t>(c) * P.L;
  const float* __restrict__ xf = P.x + (PCM ? 0 : base);
  const int16_t* __restrict__ xi = P.x16 + (PCM ? base : 0);
  using In = typename std::conditional<PCM, int16_t, float>::type;
  const In* __restrict__ xin = PCM ? reinterpret_cast<const In*>(xi) : reinterpret_cast<const In*>(xf);
  auto cvt = [](In s) -> float { return PCM ? static_cast<float>(s) * (1.0f / 32768.0f) : static_cast<float>(s); };

  // ---- sweep 1: leaf sums.  A thread owns two adjacent accumulators (2 jj, 2 jj + 1) of numpy's eight, four lanes own a
  // leaf: 8-byte loads, twice the bytes in flight per thread of the one-accumulator form, and the first level of the final
  // combination ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) is thread-local.
  float mx = 0.f;
  const int jj = tid & 3;
  const unsigned gmask = 0xFu << (4 * (lane >> 2));
  const int32_t* __restrict__ lo_p = P.leaf_off + leaf0;
  const int32_t* __restrict__ ll_p = P.leaf_len + leaf0;
  using In2 = typename std::conditional<PCM, uint32_t, float2>::type;
  auto cvt2 = [](In2 s) -> float2 {
    if constexpr (PCM) {
      return make_float2(static_cast<float>(static_cast<int16_t>(s & 0xffffu)) * (1.0f / 32768.0f),
                         static_cast<float>(static_cast<int16_t>(s >> 16)) * (1.0f / 32768.0f));
    } else {
      return s;
    }
  };
  for (int l = tid >> 2; l < lpc; l += 128) {
    const int rows = ll_p[l] >> 3;              // >= 8 (launch_prep checks the plan): the first eight rows load unconditionally
    const In2* __restrict__ q = reinterpret_cast<const In2*>(xin + lo_p[l] + 2 * jj);      // row u at q[4 u]
    float2 v8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v8[u] = cvt2(q[4 * u]);
    const float2 v9 = rows > 8 ? cvt2(q[32]) : make_float2(0.f, 0.f);   // the ninth row of a 72-sample leaf rides along with the first eight
    mx = fmaxf(mx, fmaxf(fabsf(v8[0].x), fabsf(v8[0].y)));
    float r0 = __fmul_rn(v8[0].x, v8[0].x), r1 = __fmul_rn(v8[0].y, v8[0].y);
#pragma unroll
    for (int u = 1; u < 8; ++u) {
      mx = fmaxf(mx, fmaxf(fabsf(v8[u].x), fabsf(v8[u].y)));
      r0 = __fadd_rn(r0, __fmul_rn(v8[u].x, v8[u].x));
      r1 = __fadd_rn(r1, __fmul_rn(v8[u].y, v8[u].y));
    }
    if (rows > 8) {
      mx = fmaxf(mx, fmaxf(fabsf(v9.x), fabsf(v9.y)));
      r0 = __fadd_rn(r0, __fmul_rn(v9.x, v9.x));
      r1 = __fadd_rn(r1, __fmul_rn(v9.y, v9.y));
#pragma unroll 1
      for (int u = 9; u < rows; ++u) {
        const float2 v = cvt2(q[4 * u]);
        mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
        r0 = __fadd_rn(r0, __fmul_rn(v.x, v.x));
        r1 = __fadd_rn(r1, __fmul_rn(v.y, v.y));
      }
    }
    float r = __fadd_rn(r0, r1);                             // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7))
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 1, 4));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 2, 4));
    if (jj == 0) s_leaf[l] = r;
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_wmax[warp] = mx;
  __syncthreads();

  // ---- this CTA's subtree: adjacent pairing at every level = xor butterflies (fl(a + b) = fl(b + a))
  const int nw = lpc >> 5;                      // warps holding 32 leaf sums each (1 .. 16)
  if (warp < nw) {
    float v = s_leaf[tid];
    for (int o = 1; o < 32; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s_wroot[warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
    float v = lane < nw ? s_wroot[lane] : 0.f;
    for (int o = 1; o < nw; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    v = __shfl_sync(0xffffffffu, v, 0);         // lanes >= nw took no part in the butterfly
    float m = lane < 16 ? s_wmax[lane] : 0.f;
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    m = __shfl_sync(0xffffffffu, m, 0);
    if (lane < kPrepCluster) {                  // lane r delivers to CTA r
      st_cluster_f32(&s_xroot[rank], lane, v);
      st_cluster_f32(&s_xmax[rank], lane, m);
    }
  }
  cluster_sync_all();                           // release / acquire: the eight roots are visible in every CTA

  if (tid == 0) {
    const float root = __fadd_rn(__fadd_rn(__fadd_rn(s_xroot[0], s_xroot[1]), __fadd_rn(s_xroot[2], s_xroot[3])),
                                 __fadd_rn(__fadd_rn(s_xroot[4], s_xroot[5]), __fadd_rn(s_xroot[6], s_xroot[7])));
    float m = s_xmax[0];
#pragma unroll
    for (int r = 1; r < kPrepCluster; ++r) m = fmaxf(m, s_xmax[r]);
    const float rms = __fsqrt_rn(__fdiv_rn(root, static_cast<float>(P.L)));
    int scaled = 0;
    float scale = 1.0f;
    if (P.scalar_f64) {                         // numpy 1.26.4 scalar rules, see prep_body
      const double rd = static_cast<double>(rms);
      if (P.normalize && !(rd < P.rms_min_d)) {
        scaled = 1;
        scale = __double2float_rn(__ddiv_rn(P.target_d, __dadd_rn(rd, P.eps_d)));
      }
    } else if (P.normalize && !(rms < P.rms_min)) {
      scaled = 1;
      scale = __fdiv_rn(P.target_rms, __fadd_rn(rms, P.eps));
    }
    s_scale = scale;
    s_scaled = scaled;
    if (rank == 0) {
      const float bound = scaled ? fminf(__fmul_rn(m, scale), 1.0f) : m;
      int s = 0;
      if (bound > 0.f && bound < 3.0e38f) {
        s = P.headroom_log2 - ilogbf(bound);
        s = s > 60 ? 60 : (s < -60 ? -60 : s);
      }
      if (P.inv2) P.inv2[c] = (rms - rms == 0.0f) ? ldexpf(1.0f, -2 * (s + P.dft_scale_log2)) : NAN;   // NaN / Inf chunk: see prep_body
      if (P.chunk_par) P.chunk_par[c] = make_float4(scale, ldexpf(1.0f, s), scaled ? 1.0f : 0.0f, 0.0f);
      if (P.ok) P.ok[c] = P.normalize ? static_cast<uint8_t>(scaled) : static_cast<uint8_t>(1);
      if (P.rms) P.rms[c] = rms;
    }
  }
  __syncthreads();
  const float scale = s_scale;
  const int scaled = s_scaled;

  // ---- sweep 2 over this CTA's own samples [o0, o1): both multiples of 8
  const int o0 = P.leaf_off[leaf0], o1 = (leaf0 + lpc < P.n_leaves) ? P.leaf_off[leaf0 + lpc] : P.L;
  const int n8 = (o1 - o0) >> 3;
  auto load8 = [&](int i, float (&v)[8]) {
    if (PCM) {
      const uint4 u = *reinterpret_cast<const uint4*>(xi + o0 + 8 * i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[2 * k] = static_cast<float>(static_cast<int16_t>(w[k] & 0xffffu)) * (1.0f / 32768.0f);
        v[2 * k + 1] = static_cast<float>(static_cast<int16_t>(w[k] >> 16)) * (1.0f / 32768.0f);
      }
    } else {
      const float4 a = *reinterpret_cast<const float4*>(xf + o0 + 8 * i), b = *reinterpret_cast<const float4*>(xf + o0 + 8 * i + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
  };
  if (P.q16 != nullptr) {
    uint16_t* __restrict__ qc = P.q16 + base + o0;
#pragma unroll 4
    for (int i = tid; i < n8; i += 512) {
      float v[8];
      load8(i, v);
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w[k] = static_cast<uint32_t>(pcm16_of(v[2 * k], scale, scaled) + 32768) |
               (static_cast<uint32_t>(pcm16_of(v[2 * k + 1], scale, scaled) + 32768) << 16);
      __stcs(reinterpret_cast<uint4*>(qc) + i, make_uint4(w[0], w[1], w[2], w[3]));
    }
  }
  if (P.y != nullptr) {
    float* __restrict__ yc = P.y + base + o0;
#pragma unroll 2
    for (int i = tid; i < n8; i += 512) {
      float v[8];
      load8(i, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = finish_sample(v[k], scale, scaled, P.quantize);
      __stcs(reinterpret_cast<float4*>(yc) + 2 * i, make_float4(v[0], v[1], v[2], v[3]));
      __stcs(reinterpret_cast<float4*>(yc) + 2 * i + 1, make_float4(v[4], v[5], v[6], v[7]));
    }
  }
}

int launch_prep(avld_ctx* c, const float* x, const int16_t* x16, float* y_out, bool write_operand, bool normalize, uint8_t* ok, float* rms,
                int n, float target_rms, float rms_min, float eps, int quantize, cudaStream_t st) {
  if (n <= 0) return AVLD_OK;
  PrepParams P{};
  P.x = x;
  P.x16 = x16;
  P.y = y_out;
  P.chunk_par = write_operand ? c->d_chunk_par : nullptr;
  if (write_operand) {      // operand source of this pass, consumed by launch_fold3
    c->cur_x = x;
    c->cur_x16 = x16;
    c->cur_quantize = quantize ? 1 : 0;
    // the uint4 stores of phase 2c need 16-byte aligned chunk rows
    c->cur_q16 = (quantize && c->d_q16 && (c->L & 7) == 0) ? c->d_q16 : nullptr;
    P.q16 = c->d_q16 && c->cur_q16 ? c->d_q16 : nullptr;
  }
  P.inv2 = write_operand ? c->d_inv2 : nullptr;
  P.ok = ok;
  P.rms = rms;
  P.leaf_off = c->d_leaf_off;
  P.leaf_len = c->d_leaf_len;
  P.nodes = c->d_nodes;
  P.level_start = c->d_level_start;
  P.n_leaves = c->n_leaves;
  P.n_nodes = c->n_nodes;
  P.n_levels = c->n_levels;
  P.L = c->L;
  P.n_fft = c->p.n_fft;
  P.hop = c->p.hop;
  P.target_rms = target_rms;
  P.rms_min = rms_min;
  P.eps = eps;
  P.scalar_f64 = c->scalar_f64;
  P.target_d = c->norm_target;
  P.rms_min_d = c->norm_rms_min;
  P.eps_d = c->norm_eps;
  if (c->scalar_f64 && normalize)      // the per-call floats must be the float32 images of the context's doubles
    AVLD_CHECK(target_rms == static_cast<float>(c->norm_target) && rms_min == static_cast<float>(c->norm_rms_min) &&
                   eps == static_cast<float>(c->norm_eps),
               AVLD_ERR_INVALID, "numpy-1 scalar mode: target_rms / rms_min / eps differ from avld_ctx_set_normalization");
  P.normalize = normalize ? 1 : 0;
  P.quantize = quantize ? 1 : 0;
  P.dft_scale_log2 = c->dft_scale_log2;
  P.headroom_log2 = 13;
  const size_t smem = static_cast<size_t>(c->n_leaves + c->n_nodes + 1) * sizeof(float);
  P.n = n;
  // wide form: regular leaf plan, 16-byte aligned rows (chunk rows are L samples apart, L % 8 == 0)
  bool wide = c->leaves_regular && (x16 ? reinterpret_cast<uintptr_t>(x16) % 8 == 0 : reinterpret_cast<uintptr_t>(x) % 16 == 0);
#ifdef AVLD_BRINGUP
  if (std::getenv("AVLD_PREP_NARROW")) wide = false;
#endif
  // cluster form: perfect plan, 16-byte aligned rows on both sides
  bool clustered = wide && c->tree_perfect && c->min_leaf_rows >= 8 && c->n_leaves / kPrepCluster <= 512 &&
                   (x16 ? reinterpret_cast<uintptr_t>(x16) % 16 == 0 : true) &&
                   (y_out == nullptr || reinterpret_cast<uintptr_t>(y_out) % 16 == 0);
#ifdef AVLD_BRINGUP
  if (std::getenv("AVLD_PREP_NOCLUSTER")) clustered = false;
#endif
  if (clustered) {
    LaunchScope ls(c, ST_PREP, st);
    if (x16) prep_cluster_kernel<true><<<n * kPrepCluster, 512, 0, st>>>(P);
    else prep_cluster_kernel<false><<<n * kPrepCluster, 512, 0, st>>>(P);
  } else if (wide) {
    AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(prep_wide_kernel), 164 * 1024));
    LaunchScope ls(c, ST_PREP, st);
    prep_wide_kernel<<<n, 1024, smem, st>>>(P);
  } else {
    AVLD_TRY(ensure_dyn_smem(c, reinterpret_cast<const void*>(prep_kernel), 164 * 1024));
    LaunchScope ls(c, ST_PREP, st);
    prep_kernel<<<n, 512, smem, st>>>(P);
  }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld

using namespace avld;

extern "C" int avld_rms_normalize(avld_ctx* c, const float* x, float* y, uint8_t* ok, float* rms, int64_t n,
                                  float target_rms, float rms_min, float eps, int quantize_pcm16, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(x && y, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0, AVLD_ERR_INVALID, "negative n");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t step = 1 << 20;   // grid size limit is far above this; chunk the launch only for int safety
  for (int64_t i = 0; i < n; i += step) {
    const int m = static_cast<int>(n - i < step ? n - i : step);
    AVLD_TRY(launch_prep(c, x + i * c->L, nullptr, y + i * c->L, false, true, ok ? ok + i : nullptr, rms ? rms + i : nullptr, m,
                         target_rms, rms_min, eps, quantize_pcm16, st));
  }
  return AVLD_OK;
}
