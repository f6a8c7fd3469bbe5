// fold3.cu -- operand preparation of the three-times folded STFT: raw chunk -> windowed, folded fp16 hi/lo tiles.
// With u[k] = w[k] xs[f*hop + k] (w = periodic Hann, xs = normalised, clipped, PCM_16-quantised, power-of-two-scaled,
// reflect-padded audio; replaces the frame / window / rfft front of librosa.stft, map_detector_core.py:219-228), N = n_fft,
// H = N/2, Q = N/4, E = N/8, the time-reversal symmetry of a real DFT is applied up to three times: the odd
// bins keep Q taps per cos / sin part, the bins = 0 and = 2 mod 4 keep E taps each.
//
// Operand layout ("A3", tile-major): [frame tile of 128 rows][64-tap K block of the N columns][hi | lo][128 rows][64 taps]
// fp16, columns = [ odd: cos (Q) | sin (Q) | b = 0 mod 4: cos (E) | sin (E) | b = 2 mod 4: cos (E) | sin (E) ].  One
// (tile, K block) is 32 KB of contiguous memory = exactly one TMA box of the GEMM (dftf3.cu): a row-major [frame][N]
// matrix made every box 128 row segments of 128 bytes with a 4 KB pitch, which the GEMM could only stream at about half
// the DRAM rate (profiles/r02a_dft_probes.txt).
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "sample.cuh"

namespace avld {

struct Fold3Params {
  const float* x;       // [n][L] or NULL
  const int16_t* x16;   // [n][L] PCM_16 or NULL
  const uint16_t* q16;  // [n][L] normalised PCM_16 + 32768 written by prep_kernel, or NULL (fold3_kernel<2> reads it instead of x)
  const float4* chunk_par;
  const float* win;     // [H + 1] periodic Hann, win[k] = 0.5 - 0.5 cos(2 pi k / N)
  __half* a3;           // tile-major folded operand, see the header
  float4* edge;         // [n*F] one self-paired tap per bin class (indexed by class), see the kernels
  int F, hop, n_fft, L, quantize;
  int vec_ok;
  long long total;      // n * F * (Q / 8) threads
};

namespace {

template <bool PCM>
__device__ __forceinline__ float raw_sample(const float* xf, const int16_t* xi, int src) {
  return PCM ? static_cast<float>(xi[src]) * (1.0f / 32768.0f) : xf[src];
}

// padded index p -> source index of np.pad(y, n_fft // 2, mode="reflect")
__device__ __forceinline__ int reflect_src(int p, int half, int L) {
  int src = p - half;
  if (src < 0) src = -src;
  if (src >= L) src = 2 * (L - 1) - src;
  return src;
}

// 8 consecutive samples starting at the 16-byte aligned source index `src`
template <bool PCM>
__device__ __forceinline__ void load8(const float* xf, const int16_t* xi, int src, float (&r)[8]) {
  if (PCM) {
    const uint4 u = *reinterpret_cast<const uint4*>(xi + src);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[2 * i] = static_cast<float>(static_cast<int16_t>(w[i] & 0xffffu)) * (1.0f / 32768.0f);
      r[2 * i + 1] = static_cast<float>(static_cast<int16_t>(w[i] >> 16)) * (1.0f / 32768.0f);
    }
  } else {
    const float4 v0 = *reinterpret_cast<const float4*>(xf + src), v1 = *reinterpret_cast<const float4*>(xf + src + 4);
    r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
  }
}

// 8 consecutive taps of one frame row into the tile-major operand: fp16 hi / lo split, hi tile at dst, lo tile 128 * 64 halves on
__device__ __forceinline__ void split_store_at(__half* dst, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 hh = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
    const float2 r = sub2(make_float2(v[2 * q], v[2 * q + 1]), __half22float2(hh));
    const __half2 ll = __floats2half2_rn(r.x, r.y);
    h[q] = *reinterpret_cast<const uint32_t*>(&hh);
    l[q] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  __stcs(reinterpret_cast<uint4*>(dst), make_uint4(h[0], h[1], h[2], h[3]));
  __stcs(reinterpret_cast<uint4*>(dst + 128 * 64), make_uint4(l[0], l[1], l[2], l[3]));
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// fold3_kernel.  With Q = N/4, E = N/8 and, for
// k = 1 .. E-1,  u1 = u[k], u2 = u[N-k], u3 = u[2Q-k], u4 = u[2Q+k], u5 = u[Q-k], u6 = u[3Q+k], u7 = u[Q+k], u8 = u[3Q-k]:
//   P = u1+u2+u3+u4, P' = u5+u6+u7+u8, R = (u1-u2)-(u3-u4), R' = (u5-u6)-(u7-u8)
//   b = 0 mod 4:  Re X = sum_k (P + P') cos(2 pi k b / N) + edge cos(pi b / 4),    -Im X = sum_k (R - R') sin(.)
//   b = 2 mod 4:  Re X = sum_k (P - P') cos(.),                                   -Im X = sum_k (R + R') sin(.) + edge sin(pi b / 4)
//   odd b (k < Q): Re X = sum_k ((u1+u2)-(u3+u4)) cos(.),  -Im X = sum_k ((u1-u2)+(u3-u4)) sin(.) + edge sin(pi b / 2)
// One thread = 8 consecutive k < E of one frame: eight runs of the chunk (the four of the second half also give the odd
// class at k' = Q - k, which lands on an aligned block when shifted by one tap), 256 B out.
// ------------------------------------------------------------------------------------------------
namespace {

// Sample source of fold3_kernel.  SRC 0: float32 chunk, 1: raw PCM_16 chunk (both normalised on the fly by fin()),
// 2: the normalised PCM_16 integers q (+32768) left by prep_kernel -- the power-of-two factor pow2 / 32768 then goes into the
// window values instead (exact, so all three give the same bits).  SRC 2 hands the samples out *biased*: the float
// kBias + q = 2^23 + (q + 32768), one PRMT away from the stored integer.  The fold only ever needs a sample of an ascending
// run together with one of a descending run, as a - b and a + b: a - b is the difference of the biased values, a + b is
// (biased a - 2 kBias) + biased b -- both exact (integers below 2^24), one bias removal per *pair* of samples instead of two.
template <int SRC>
struct Samples {
  const float* xf;
  const int16_t* xi;
  const uint16_t* xq;
  static constexpr float kBias = SRC == 2 ? 8421376.0f : 0.0f;                        // 2^23 + 2^15
  __device__ __forceinline__ float biased(int i) const {
    if (SRC == 2) return __uint_as_float(0x4B000000u | xq[i]);
    return raw_sample<SRC == 1>(xf, xi, i);
  }
  __device__ __forceinline__ float raw(int i) const { return SRC == 2 ? biased(i) - kBias : biased(i); }
  // 8 consecutive (biased) samples from the 16-byte aligned index i
  __device__ __forceinline__ void load8(int i, float (&r)[8]) const {
    if (SRC == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(xq + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[2 * j] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7610));
        r[2 * j + 1] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7632));
      }
    } else {
      avld::load8<SRC == 1>(xf, xi, i, r);
    }
  }
  // 9 samples x[i0 + j], j = 0..8 (ascending): aligned vector of 8 + one scalar
  __device__ __forceinline__ void load9_up(int i0, float (&r)[9]) const {
    float v[8];
    load8(i0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = v[j];
    r[8] = biased(i0 + 8);
  }
  // 9 samples x[i0 - j], j = 0..8 (descending): scalar at i0 + aligned vector of 8 below it
  __device__ __forceinline__ void load9_down(int i0, float (&r)[9]) const {
    float v[8];
    load8(i0 - 8, v);
    r[0] = biased(i0);
#pragma unroll
    for (int j = 1; j < 9; ++j) r[j] = v[8 - j];
  }
};

}  // namespace

// NFFT = 0: n_fft at run time; otherwise a compile-time n_fft (every row offset becomes an immediate).  Thread indices are
// 32-bit (launch_fold3 checks total < 2^31): the 64-bit divisions of the generic form were ~10 % of the instructions.
template <int SRC, int MINB, int NFFT>
__global__ void __launch_bounds__(128, MINB) fold3_kernel(const Fold3Params P) {
  const int N = NFFT ? NFFT : P.n_fft, H = N >> 1, Q = N >> 2, E = N >> 3, per_frame = E >> 3;
  const unsigned total = static_cast<unsigned>(P.total), F = static_cast<unsigned>(P.F);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned row = t / static_cast<unsigned>(per_frame);
    const int k0 = static_cast<int>(t - row * per_frame) << 3;
    const unsigned chunk = row / F;
    const int f = static_cast<int>(row - chunk * F);
    const float4 par = P.chunk_par[chunk];
    const float scale = par.x, pow2 = par.y;
    const int scaled = par.z != 0.f;
    const size_t x0 = static_cast<size_t>(chunk) * P.L;
    const Samples<SRC> X{SRC == 0 ? P.x + x0 : nullptr, SRC == 1 ? P.x16 + x0 : nullptr, SRC == 2 ? P.q16 + x0 : nullptr};
    const float wscale = SRC == 2 ? pow2 * (1.0f / 32768.0f) : 1.0f;      // a power of two
    auto fin = [&](float v) -> float { return SRC == 2 ? v : finish_sample(v, scale, scaled, P.quantize) * pow2; };
    const int pf = f * P.hop;
    const int s0 = pf - H;                               // source index of tap 0 when nothing is reflected
    // x1[q] = xs[k], x4[q] = xs[2Q+k], x6[q] = xs[3Q+k], x7[q] = xs[Q+k]   (ascending in q, k = k0 + q)
    // x2[q] = xs[N-k], x3[q] = xs[2Q-k], x5[q] = xs[Q-k], x8[q] = xs[3Q-k] (descending)
    float x1[9], x2[9], x3[9], x4[9], x5[9], x6[9], x7[9], x8[9];
    if (P.vec_ok && s0 >= 0 && s0 + N + 8 <= P.L) {
      X.load9_up(s0 + k0, x1);
      X.load9_up(s0 + H + k0, x4);
      X.load9_up(s0 + H + Q + k0, x6);
      X.load9_up(s0 + Q + k0, x7);
      X.load9_down(s0 + N - k0, x2);
      X.load9_down(s0 + H - k0, x3);
      X.load9_down(s0 + Q - k0, x5);
      X.load9_down(s0 + H + Q - k0, x8);
    } else {
      const int plen = P.L + N;
      auto get = [&](int p) -> float { return (p >= 0 && p < plen) ? X.biased(reflect_src(p, H, P.L)) : (SRC == 2 ? Samples<SRC>::kBias : 0.f); };
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const int k = k0 + q;
        x1[q] = get(pf + k);         x2[q] = get(pf + N - k);
        x3[q] = get(pf + H - k);     x4[q] = get(pf + H + k);
        x5[q] = get(pf + Q - k);     x6[q] = get(pf + H + Q + k);
        x7[q] = get(pf + Q + k);     x8[q] = get(pf + H + Q - k);
      }
    }
    float oc[8], os[8], c0[8], s0v[8], c2[8], s2v[8], oc2[8], os2[8];
    // window values of this 8-tap block: nine coalesced float4 loads (table layout: ctx.cu)
    const float4* wt = reinterpret_cast<const float4*>(P.win + ((H + 1 + 3) & ~3)) + (k0 >> 3);
    float wk8[8], wh8[8], w58[9], w78[9];
    {
      const float4 t0 = __ldg(wt), t1 = __ldg(wt + per_frame), t2 = __ldg(wt + 2 * per_frame), t3 = __ldg(wt + 3 * per_frame);
      const float4 t4 = __ldg(wt + 4 * per_frame), t5 = __ldg(wt + 5 * per_frame), t6 = __ldg(wt + 6 * per_frame);
      const float4 t7 = __ldg(wt + 7 * per_frame), t8 = __ldg(wt + 8 * per_frame);
      wk8[0] = t0.x; wk8[1] = t0.y; wk8[2] = t0.z; wk8[3] = t0.w; wk8[4] = t1.x; wk8[5] = t1.y; wk8[6] = t1.z; wk8[7] = t1.w;
      wh8[0] = t2.x; wh8[1] = t2.y; wh8[2] = t2.z; wh8[3] = t2.w; wh8[4] = t3.x; wh8[5] = t3.y; wh8[6] = t3.z; wh8[7] = t3.w;
      w58[0] = t4.x; w58[1] = t4.y; w58[2] = t4.z; w58[3] = t4.w; w58[4] = t5.x; w58[5] = t5.y; w58[6] = t5.z; w58[7] = t5.w;
      w78[0] = t6.x; w78[1] = t6.y; w78[2] = t6.z; w78[3] = t6.w; w78[4] = t7.x; w78[5] = t7.y; w78[6] = t7.z; w78[7] = t7.w;
      w58[8] = t8.x; w78[8] = t8.y;
    }
    if (SRC == 2) {                                      // exact (a power of two), two taps per FMUL2
      const float2 ws2 = make_float2(wscale, wscale);
#pragma unroll
      for (int q = 0; q < 8; q += 2) {
        const float2 a = mul2(make_float2(w58[q], w58[q + 1]), ws2), b = mul2(make_float2(w78[q], w78[q + 1]), ws2);
        const float2 c = mul2(make_float2(wk8[q], wk8[q + 1]), ws2), d = mul2(make_float2(wh8[q], wh8[q + 1]), ws2);
        w58[q] = a.x; w58[q + 1] = a.y; w78[q] = b.x; w78[q + 1] = b.y;
        wk8[q] = c.x; wk8[q + 1] = c.y; wh8[q] = d.x; wh8[q + 1] = d.y;
      }
      w58[8] *= wscale;
      w78[8] *= wscale;
    }
    // SRC 0 / 1: finish every sample (scale, clip, PCM_16 round trip, power of two); SRC 2: nothing to do, the bias stays on
    constexpr float kB = Samples<SRC>::kBias;
    if (SRC != 2) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        x5[q] = fin(x5[q]); x6[q] = fin(x6[q]); x7[q] = fin(x7[q]); x8[q] = fin(x8[q]);
        if (q < 8) { x1[q] = fin(x1[q]); x2[q] = fin(x2[q]); x3[q] = fin(x3[q]); x4[q] = fin(x4[q]); }
      }
    }
    if (k0 == 0) {                                       // k = 0: u[0] and u[H] pair with nothing, u[Q] and u[3Q] are one pair, not two
      x2[0] = kB; x4[0] = kB; x7[0] = kB; x8[0] = kB;
    }
    // Two taps per instruction (FADD2 / FFMA2): the integer sums and differences are exact, only the window products round
    // (each separately, mul2_rn), so the bits are those of the scalar form.  Ascending runs: x1, x4, x6, x7.
    const float2 b2 = make_float2(2.0f * kB, 2.0f * kB);
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      const float2 W5 = make_float2(w58[q], w58[q + 1]), W7 = make_float2(w78[q], w78[q + 1]);
      const float2 WK = make_float2(wk8[q], wk8[q + 1]), WH = make_float2(wh8[q], wh8[q + 1]);
      const float2 A1 = make_float2(x1[q], x1[q + 1]), A2 = make_float2(x2[q], x2[q + 1]), A3 = make_float2(x3[q], x3[q + 1]);
      const float2 A4 = make_float2(x4[q], x4[q + 1]), A5 = make_float2(x5[q], x5[q + 1]), A6 = make_float2(x6[q], x6[q + 1]);
      const float2 A7 = make_float2(x7[q], x7[q + 1]), A8 = make_float2(x8[q], x8[q + 1]);
      const float2 s56 = SRC == 2 ? add2(sub2(A6, b2), A5) : add2(A5, A6), s78 = SRC == 2 ? add2(sub2(A7, b2), A8) : add2(A7, A8);
      const float2 fp = mul2_rn(W5, s56), fm = mul2_rn(W7, s78);
      const float2 gp = mul2_rn(W5, sub2(A5, A6)), gm = mul2_rn(W7, sub2(A7, A8));
      const float2 o2c = sub2(fp, fm), o2s = add2(gp, gm);      // odd bins at k' = Q - k: block element 8 - q of [Q-k0-8, Q-k0)
      if (q >= 1) {
        oc2[8 - q] = o2c.x;
        os2[8 - q] = o2s.x;
      }
      oc2[7 - q] = o2c.y;
      os2[7 - q] = o2s.y;
      const float2 s12 = SRC == 2 ? add2(sub2(A1, b2), A2) : add2(A1, A2), s34 = SRC == 2 ? add2(sub2(A4, b2), A3) : add2(A3, A4);
      const float2 ep = mul2_rn(WK, s12), em = mul2_rn(WH, s34);
      const float2 op = mul2_rn(WK, sub2(A1, A2)), om = mul2_rn(WH, sub2(A3, A4));
      const float2 Pk = add2(ep, em), Pm = add2(fp, fm), Rk = sub2(op, om), Rm = sub2(gp, gm);
      const float2 voc = sub2(ep, em), vos = add2(op, om), vc0 = add2(Pk, Pm), vc2 = sub2(Pk, Pm);
      const float2 vs0 = sub2(Rk, Rm), vs2 = add2(Rk, Rm);
      oc[q] = voc.x;  oc[q + 1] = voc.y;
      os[q] = vos.x;  os[q + 1] = vos.y;
      c0[q] = vc0.x;  c0[q + 1] = vc0.y;
      c2[q] = vc2.x;  c2[q + 1] = vc2.y;
      s0v[q] = vs0.x; s0v[q + 1] = vs0.y;
      s2v[q] = vs2.x; s2v[q + 1] = vs2.y;
    }
    {                                                    // the ninth tap only feeds the odd class (k' = Q - k0 - 8)
      const float s56 = SRC == 2 ? (x6[8] - 2.0f * kB) + x5[8] : x5[8] + x6[8], s78 = SRC == 2 ? (x7[8] - 2.0f * kB) + x8[8] : x7[8] + x8[8];
      const float fp = __fmul_rn(w58[8], s56), fm = __fmul_rn(w78[8], s78);
      const float gp = __fmul_rn(w58[8], x5[8] - x6[8]), gm = __fmul_rn(w78[8], x7[8] - x8[8]);
      oc2[0] = fp - fm;
      os2[0] = gp + gm;
    }
    if (k0 == 0) {                                       // the sine parts have no tap 0
      os[0] = 0.f;
      s0v[0] = 0.f;
      s2v[0] = 0.f;
    }
    // Destination of an 8-tap block at column col of this frame row: (tile, K block col / 64, hi, row, col % 64).  The six
    // ascending streams sit at k0 plus a multiple of 64 columns, the two odd-class streams at Q - 8 - k0 (+ Q): two row
    // pointers, every stream a compile-time offset from one of them (the per-stream address arithmetic was ~50 instructions).
    const size_t base = (static_cast<size_t>(row >> 7) * (N >> 6) * 256 + (row & 127)) * 64;
    constexpr size_t kBlk = 2 * 128 * 64;                  // halves per K block (hi + lo tiles)
    const int kr = Q - 8 - k0;
    __half* const fwd = P.a3 + base + static_cast<size_t>(k0 >> 6) * kBlk + (k0 & 63);
    __half* const rev = P.a3 + base + static_cast<size_t>(kr >> 6) * kBlk + (kr & 63);
    split_store_at(fwd, oc);
    split_store_at(fwd + (Q >> 6) * kBlk, os);
    split_store_at(rev, oc2);
    split_store_at(rev + (Q >> 6) * kBlk, os2);
    split_store_at(fwd + (H >> 6) * kBlk, c0);
    split_store_at(fwd + ((H + E) >> 6) * kBlk, s0v);
    split_store_at(fwd + ((H + Q) >> 6) * kBlk, c2);
    split_store_at(fwd + ((H + Q + E) >> 6) * kBlk, s2v);
    // ---- edge terms of this frame row: O[Q], P[E], R[E] from the six samples at taps E, N-E, 3E, 5E, Q, 3Q
    auto xs = [&](int tap) -> float {
      const float r = X.raw(reflect_src(pf + tap, H, P.L));
      return SRC == 2 ? r * wscale : fin(r);
    };
    auto edge_of = [&](float xe, float x7e, float x3e, float x5e, float xq, float x3q) -> float4 {
      const float wq = P.win[Q], we = P.win[E], w3 = P.win[Q + E];     // w[3E] = w[N - 5E] ...: w[Q+E] = w[H+Q-E+...]
      const float e_odd = __fmul_rn(wq, xq - x3q);                                               // O[Q]
      const float e_m0 = __fadd_rn(__fmul_rn(we, xe + x7e), __fmul_rn(w3, x3e + x5e));           // P[E] = u[E] + u[N-E] + u[3E] + u[5E]
      const float e_m2 = __fsub_rn(__fmul_rn(we, xe - x7e), __fmul_rn(w3, x3e - x5e));           // R[E]
      return make_float4(e_odd, e_m0, e_m2, 0.f);                      // class 0 = odd, 1 = 0 mod 4, 2 = 2 mod 4
    };
    if (per_frame == 32) {
      // One warp = one frame row (the loop bounds are multiples of 32).  Left to the k0 = 0 thread alone, i.e. lane 0 of
      // every warp, these ~100 instructions ran with one active lane and took a tenth of the kernel's issue slots (ncu source
      // page, r02f); here lanes 0..5 fetch one sample each and lane 0 combines them.
      const int lane = threadIdx.x & 31;
      const int tap = E * ((0x625371 >> (4 * lane)) & 0xF);            // E * {1, 7, 3, 5, 2, 6}: E, N-E, 3E, 5E, Q, 3Q
      const float sv = lane < 6 ? xs(tap) : 0.f;
      const float xe = __shfl_sync(0xffffffffu, sv, 0), x7e = __shfl_sync(0xffffffffu, sv, 1), x3e = __shfl_sync(0xffffffffu, sv, 2);
      const float x5e = __shfl_sync(0xffffffffu, sv, 3), xq = __shfl_sync(0xffffffffu, sv, 4), x3q = __shfl_sync(0xffffffffu, sv, 5);
      if (lane == 0) P.edge[row] = edge_of(xe, x7e, x3e, x5e, xq, x3q);
    } else if (k0 == 0) {
      P.edge[row] = edge_of(xs(E), xs(N - E), xs(Q + E), xs(H + E), xs(Q), xs(H + Q));
    }
  }
}

int launch_fold3(avld_ctx* c, int n, cudaStream_t st) {
  if (n <= 0) return AVLD_OK;
  Fold3Params P{};
  P.x = c->cur_x;
  P.x16 = c->cur_x16;
  P.q16 = c->cur_q16;
  P.chunk_par = c->d_chunk_par;
  P.win = c->d_win;
  P.a3 = c->d_A3;
  P.edge = c->d_edge;
  P.F = c->F;
  P.hop = c->p.hop;
  P.n_fft = c->p.n_fft;
  P.L = c->L;
  P.quantize = c->cur_quantize;
  P.vec_ok = (c->L % 8 == 0) && (c->p.hop % 8 == 0) && (reinterpret_cast<uintptr_t>(P.x) % 16 == 0) &&
             (reinterpret_cast<uintptr_t>(P.x16) % 16 == 0);
  P.total = static_cast<long long>(n) * c->F * (c->p.n_fft / 64);
  AVLD_CHECK(P.total < (1ll << 31), AVLD_ERR_UNSUPPORTED, "fold3: more than 2^31 operand threads in one pass");
  // ~145 registers per thread: 128-thread blocks keep three to four blocks per SM resident
  const long long b3 = (P.total + 127) / 128, cap = static_cast<long long>(c->sm_count) * 64;
  const int g3 = static_cast<int>(b3 < cap ? b3 : cap);
  const bool n2k = c->p.n_fft == 2048;
  {
    LaunchScope ls(c, ST_FOLD, st);
    if (P.q16 != nullptr) {
      if (n2k) fold3_kernel<2, 4, 2048><<<g3, 128, 0, st>>>(P);
      else fold3_kernel<2, 4, 0><<<g3, 128, 0, st>>>(P);
    } else if (P.x16 != nullptr) {
      if (n2k) fold3_kernel<1, 3, 2048><<<g3, 128, 0, st>>>(P);
      else fold3_kernel<1, 3, 0><<<g3, 128, 0, st>>>(P);
    } else {
      if (n2k) fold3_kernel<0, 3, 2048><<<g3, 128, 0, st>>>(P);
      else fold3_kernel<0, 3, 0><<<g3, 128, 0, st>>>(P);
    }
  }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld
