// fold2.cu -- operand preparation of the twice-folded STFT ("fold2", see common.cuh): raw chunk -> windowed, two-stage
// folded fp16 hi/lo rows.  With u[k] = w[k] xs[f*hop + k] (w = periodic Hann, xs = normalised, clipped, PCM_16-quantised,
// power-of-two-scaled, reflect-padded audio), N = n_fft, H = N/2, Q = N/4 and, for k = 1 .. Q-1,
//   a = u[k], b = u[N-k], c = u[H-k], d = u[H+k]:
//   even bins:  Re X[b] = sum_k (a+b+c+d) cos(2 pi k b / N) + edge_e cos(pi b / 2),  -Im X[b] = sum_k ((a-b)-(c-d)) sin(.)
//   odd  bins:  Re X[b] = sum_k (a+b-c-d) cos(.),  -Im X[b] = sum_k ((a-b)+(c-d)) sin(.) + edge_o sin(pi b / 2)
//   k = 0 column: cos parts u[0] +- u[H], sin parts 0;   edge_e = u[Q] + u[N-Q],  edge_o = u[Q] - u[N-Q].
// (time-reversal symmetry of a real DFT, applied twice; w[N-k] = w[k], w[H-k] = w[H+k].)  Row layout of A3 (N columns):
//   [ even: cos part (Q) | sin part (Q) | odd: cos part (Q) | sin part (Q) ].
// One thread = 8 consecutive k of one frame: four runs of 8 samples (two ascending, two descending), 64 B + 64 B out.
// The window is applied here in fp32 (the products are no longer exact integers; the hi/lo split keeps 22 bits).
#include <cstdlib>
#include "common.cuh"
#include "sample.cuh"

namespace avld {

struct Fold2Params {
  const float* x;       // [n][L] or NULL
  const int16_t* x16;   // [n][L] PCM_16 or NULL
  const uint16_t* q16;  // [n][L] normalised PCM_16 + 32768 written by prep_kernel, or NULL (fold3_kernel<2> reads it instead of x)
  const float4* chunk_par;
  const float* win;     // [H + 1] periodic Hann, win[k] = 0.5 - 0.5 cos(2 pi k / N)
  __half* a_hi;         // [n*F][N]
  __half* a_lo;
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

__device__ __forceinline__ void split_store(__half* hi, __half* lo, size_t at, const float (&v)[8]) {
  __align__(16) __half h[8], l[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    h[q] = __float2half_rn(v[q]);
    l[q] = __float2half_rn(v[q] - __half2float(h[q]));
  }
  *reinterpret_cast<uint4*>(hi + at) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo + at) = *reinterpret_cast<const uint4*>(l);
}

}  // namespace

template <bool PCM>
__global__ void __launch_bounds__(256) fold2_kernel(const Fold2Params P) {
  const int N = P.n_fft, H = N >> 1, Q = N >> 2, per_frame = Q >> 3;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < P.total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = t / per_frame;                 // global frame index = chunk * F + f
    const int k0 = static_cast<int>(t - row * per_frame) << 3;
    const long long chunk = row / P.F;
    const int f = static_cast<int>(row - chunk * P.F);
    const float4 par = P.chunk_par[chunk];
    const float scale = par.x, pow2 = par.y;
    const int scaled = par.z != 0.f;
    const float* xf = PCM ? nullptr : P.x + chunk * P.L;
    const int16_t* xi = PCM ? P.x16 + chunk * P.L : nullptr;
    const int pf = f * P.hop;                            // padded index of the frame's tap 0
    // xa[q] = xs[pf + k0 + q], xd[q] = xs[pf + H + k0 + q], xb[q] = xs[pf + N - k0 - q], xc[q] = xs[pf + H - k0 - q]
    float xa[8], xb[8], xc[8], xd[8];
    const int s0 = pf - H;                               // source index of tap 0 when nothing is reflected
    if (P.vec_ok && s0 - 8 >= 0 && s0 + N + 8 <= P.L) {
      float rb[8], rc[8];
      load8<PCM>(xf, xi, s0 + k0, xa);
      load8<PCM>(xf, xi, s0 + H + k0, xd);
      load8<PCM>(xf, xi, s0 + N - k0 - 8, rb);           // taps N-k0-8 .. N-k0-1
      load8<PCM>(xf, xi, s0 + H - k0 - 8, rc);           // taps H-k0-8 .. H-k0-1
      xb[0] = raw_sample<PCM>(xf, xi, s0 + N - k0);
      xc[0] = raw_sample<PCM>(xf, xi, s0 + H - k0);
#pragma unroll
      for (int q = 1; q < 8; ++q) {
        xb[q] = rb[8 - q];
        xc[q] = rc[8 - q];
      }
    } else {
      // the first / last frames touch the reflect padding (and the last tap N of the last frame does not exist)
      const int plen = P.L + N;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int ia = pf + k0 + q, id = pf + H + k0 + q, ib = pf + N - k0 - q, ic = pf + H - k0 - q;
        xa[q] = raw_sample<PCM>(xf, xi, reflect_src(ia, H, P.L));
        xd[q] = raw_sample<PCM>(xf, xi, reflect_src(id, H, P.L));
        xb[q] = ib < plen ? raw_sample<PCM>(xf, xi, reflect_src(ib, H, P.L)) : 0.f;
        xc[q] = raw_sample<PCM>(xf, xi, reflect_src(ic, H, P.L));
      }
    }
    float c0[8], s0v[8], c1[8], s1v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = k0 + q;
      const float wk = P.win[k], wh = P.win[H - k];
      float a = finish_sample(xa[q], scale, scaled, P.quantize) * pow2;
      float b = finish_sample(xb[q], scale, scaled, P.quantize) * pow2;
      float c = finish_sample(xc[q], scale, scaled, P.quantize) * pow2;
      float d = finish_sample(xd[q], scale, scaled, P.quantize) * pow2;
      if (k == 0) {                                      // u[0] and u[H] pair with nothing
        b = 0.f;
        d = 0.f;
      }
      const float ep = wk * (a + b), em = wh * (c + d);  // a + b and c + d are exact (integers times a power of two)
      const float op = wk * (a - b), om = wh * (c - d);
      c0[q] = ep + em;
      c1[q] = ep - em;
      s0v[q] = k == 0 ? 0.f : op - om;
      s1v[q] = k == 0 ? 0.f : op + om;
    }
    const size_t base = static_cast<size_t>(row) * N + k0;
    split_store(P.a_hi, P.a_lo, base, c0);
    split_store(P.a_hi, P.a_lo, base + Q, s0v);
    split_store(P.a_hi, P.a_lo, base + H, c1);
    split_store(P.a_hi, P.a_lo, base + H + Q, s1v);
    if (k0 == 0) {
      const float wq = P.win[Q];
      const float p = finish_sample(raw_sample<PCM>(xf, xi, reflect_src(pf + Q, H, P.L)), scale, scaled, P.quantize) * pow2;
      const float m = finish_sample(raw_sample<PCM>(xf, xi, reflect_src(pf + N - Q, H, P.L)), scale, scaled, P.quantize) * pow2;
      P.edge[row] = make_float4(wq * (p + m), wq * (p - m), 0.f, 0.f);     // class 0 = even bins, class 1 = odd bins
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fold3_kernel: as fold2_kernel, with one more fold for the even bins (their cos / sin kernels are again symmetric
// about k = N/8 once restricted to b = 0 or 2 mod 4; the odd bins' symmetry is spent).  With Q = N/4, E = N/8 and, for
// k = 1 .. E-1,  u1 = u[k], u2 = u[N-k], u3 = u[2Q-k], u4 = u[2Q+k], u5 = u[Q-k], u6 = u[3Q+k], u7 = u[Q+k], u8 = u[3Q-k]:
//   P = u1+u2+u3+u4, P' = u5+u6+u7+u8, R = (u1-u2)-(u3-u4), R' = (u5-u6)-(u7-u8)
//   b = 0 mod 4:  Re X = sum_k (P + P') cos(2 pi k b / N) + edge cos(pi b / 4),    -Im X = sum_k (R - R') sin(.)
//   b = 2 mod 4:  Re X = sum_k (P - P') cos(.),                                   -Im X = sum_k (R + R') sin(.) + edge sin(pi b / 4)
//   odd b (k < Q): Re X = sum_k ((u1+u2)-(u3+u4)) cos(.),  -Im X = sum_k ((u1-u2)+(u3-u4)) sin(.) + edge sin(pi b / 2)
// Row layout of A3: [ odd: cos (Q) | sin (Q) | b = 0 mod 4: cos (E) | sin (E) | b = 2 mod 4: cos (E) | sin (E) ].
// One thread = 8 consecutive k < E of one frame: eight runs of the chunk (the four of the second half also give the odd
// class at k' = Q - k, which lands on an aligned block when shifted by one tap), 256 B out.
// ------------------------------------------------------------------------------------------------
namespace {

// Sample source of fold3_kernel.  SRC 0: float32 chunk, 1: raw PCM_16 chunk (both normalised on the fly by fin()),
// 2: the normalised PCM_16 integers q (+32768) left by prep_kernel -- raw() is then q itself and the power-of-two factor
// pow2 / 32768 goes into the window values instead (exact, so all three give the same bits).
template <int SRC>
struct Samples {
  const float* xf;
  const int16_t* xi;
  const uint16_t* xq;
  __device__ __forceinline__ float raw(int i) const {
    if (SRC == 2) return __uint_as_float(0x4B000000u | xq[i]) - 8421376.0f;       // 2^23 + u - (2^23 + 2^15)
    return raw_sample<SRC == 1>(xf, xi, i);
  }
  // 8 consecutive samples from the 16-byte aligned index i
  __device__ __forceinline__ void load8(int i, float (&r)[8]) const {
    if (SRC == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(xq + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[2 * j] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7610)) - 8421376.0f;
        r[2 * j + 1] = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7632)) - 8421376.0f;
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
    r[8] = raw(i0 + 8);
  }
  // 9 samples x[i0 - j], j = 0..8 (descending): scalar at i0 + aligned vector of 8 below it
  __device__ __forceinline__ void load9_down(int i0, float (&r)[9]) const {
    float v[8];
    load8(i0 - 8, v);
    r[0] = raw(i0);
#pragma unroll
    for (int j = 1; j < 9; ++j) r[j] = v[8 - j];
  }
};

}  // namespace

// NFFT = 0: n_fft at run time; otherwise a compile-time n_fft (every row offset becomes an immediate).  Thread indices are
// 32-bit (launch_fold2 checks total < 2^31): the 64-bit divisions of the generic form were ~10 % of the instructions.
template <int SRC, int MINB, int NFFT>
__global__ void __launch_bounds__(128, MINB) fold3_kernel(const Fold2Params P) {
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
      auto get = [&](int p) -> float { return (p >= 0 && p < plen) ? X.raw(reflect_src(p, H, P.L)) : 0.f; };
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
    if (SRC == 2) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        w58[q] *= wscale;
        w78[q] *= wscale;
        if (q < 8) {
          wk8[q] *= wscale;
          wh8[q] *= wscale;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const int k = k0 + q;
      const float w5 = w58[q], w7 = w78[q];
      const float a5 = fin(x5[q]);
      const float a6 = fin(x6[q]);
      float a7 = fin(x7[q]);
      float a8 = fin(x8[q]);
      if (k == 0) {                                      // u[Q] and u[3Q] are one pair, not two
        a7 = 0.f;
        a8 = 0.f;
      }
      const float fp = __fmul_rn(w5, a5 + a6), fm = __fmul_rn(w7, a7 + a8);   // integer sums are exact; only the window products round
      const float gp = __fmul_rn(w5, a5 - a6), gm = __fmul_rn(w7, a7 - a8);
      if (q >= 1) {                                      // odd bins at k' = Q - k: block element 8 - q of [Q-k0-8, Q-k0)
        oc2[8 - q] = fp - fm;
        os2[8 - q] = gp + gm;
      }
      if (q < 8) {
        const float wk = wk8[q], wh = wh8[q];
        const float a1 = fin(x1[q]);
        float a2 = fin(x2[q]);
        const float a3 = fin(x3[q]);
        float a4 = fin(x4[q]);
        if (k == 0) {                                    // u[0] and u[H] pair with nothing
          a2 = 0.f;
          a4 = 0.f;
        }
        const float ep = __fmul_rn(wk, a1 + a2), em = __fmul_rn(wh, a3 + a4);
        const float op = __fmul_rn(wk, a1 - a2), om = __fmul_rn(wh, a3 - a4);
        const float Pk = ep + em, Pm = fp + fm, Rk = op - om, Rm = gp - gm;
        const bool z = k == 0;
        oc[q] = ep - em;
        os[q] = z ? 0.f : op + om;
        c0[q] = Pk + Pm;
        c2[q] = Pk - Pm;
        s0v[q] = z ? 0.f : Rk - Rm;
        s2v[q] = z ? 0.f : Rk + Rm;
      }
    }
    const size_t base = static_cast<size_t>(row) * N;
    split_store(P.a_hi, P.a_lo, base + k0, oc);
    split_store(P.a_hi, P.a_lo, base + Q + k0, os);
    split_store(P.a_hi, P.a_lo, base + (Q - k0 - 8), oc2);
    split_store(P.a_hi, P.a_lo, base + Q + (Q - k0 - 8), os2);
    split_store(P.a_hi, P.a_lo, base + H + k0, c0);
    split_store(P.a_hi, P.a_lo, base + H + E + k0, s0v);
    split_store(P.a_hi, P.a_lo, base + H + Q + k0, c2);
    split_store(P.a_hi, P.a_lo, base + H + Q + E + k0, s2v);
    if (k0 == 0) {
      auto xs = [&](int tap) -> float {
        const float r = X.raw(reflect_src(pf + tap, H, P.L));
        return SRC == 2 ? r * wscale : fin(r);
      };
      const float wq = P.win[Q], we = P.win[E], w3 = P.win[Q + E];     // w[3E] = w[N - 5E] ...: w[Q+E] = w[H+Q-E+...]
      const float xe = xs(E), x7e = xs(N - E), x3e = xs(Q + E), x5e = xs(H + E);
      const float e_odd = __fmul_rn(wq, xs(Q) - xs(H + Q));                                      // O[Q]
      const float e_m0 = __fadd_rn(__fmul_rn(we, xe + x7e), __fmul_rn(w3, x3e + x5e));           // P[E] = u[E] + u[N-E] + u[3E] + u[5E]
      const float e_m2 = __fsub_rn(__fmul_rn(we, xe - x7e), __fmul_rn(w3, x3e - x5e));           // R[E]
      P.edge[row] = make_float4(e_odd, e_m0, e_m2, 0.f);                    // class 0 = odd, 1 = 0 mod 4, 2 = 2 mod 4
    }
  }
}

int launch_fold2(avld_ctx* c, int n, cudaStream_t st) {
  if (n <= 0) return AVLD_OK;
  Fold2Params P{};
  P.x = c->cur_x;
  P.x16 = c->cur_x16;
  P.q16 = c->cur_q16;
  P.chunk_par = c->d_chunk_par;
  P.win = c->d_win;
  P.a_hi = c->d_A2hi;
  P.a_lo = c->d_A2lo;
  P.edge = c->d_edge;
  P.F = c->F;
  P.hop = c->p.hop;
  P.n_fft = c->p.n_fft;
  P.L = c->L;
  P.quantize = c->cur_quantize;
  P.vec_ok = (c->L % 8 == 0) && (c->p.hop % 8 == 0) && (reinterpret_cast<uintptr_t>(P.x) % 16 == 0) &&
             (reinterpret_cast<uintptr_t>(P.x16) % 16 == 0);
  const bool three = c->f2_levels == 3;
  P.total = static_cast<long long>(n) * c->F * (c->p.n_fft / (three ? 64 : 32));
  const long long blocks = (P.total + 255) / 256;
  const long long cap = static_cast<long long>(c->sm_count) * 32;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  {
    LaunchScope ls(c, ST_FOLD, st);
    if (three) {      // ~145 registers per thread: 128-thread blocks keep three blocks per SM resident
      const long long b3 = (P.total + 127) / 128;
      const int g3 = static_cast<int>(b3 < 2 * cap ? b3 : 2 * cap);
      AVLD_CHECK(P.total < (1ll << 31), AVLD_ERR_UNSUPPORTED, "fold3: more than 2^31 operand threads in one pass");
      const bool n2k = c->p.n_fft == 2048;
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
    } else {
      if (P.x16 != nullptr) fold2_kernel<true><<<grid, 256, 0, st>>>(P);
      else fold2_kernel<false><<<grid, 256, 0, st>>>(P);
    }
  }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld
