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
#include "common.cuh"
#include "sample.cuh"

namespace avld {

struct Fold2Params {
  const float* x;       // [n][L] or NULL
  const int16_t* x16;   // [n][L] PCM_16 or NULL
  const float4* chunk_par;
  const float* win;     // [H + 1] periodic Hann, win[k] = 0.5 - 0.5 cos(2 pi k / N)
  __half* a_hi;         // [n*F][N]
  __half* a_lo;
  float2* edge;         // [n*F] (edge_e, edge_o)
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
      P.edge[row] = make_float2(wq * (p + m), wq * (p - m));
    }
  }
}

int launch_fold2(avld_ctx* c, int n, cudaStream_t st) {
  if (n <= 0) return AVLD_OK;
  Fold2Params P{};
  P.x = c->cur_x;
  P.x16 = c->cur_x16;
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
  P.total = static_cast<long long>(n) * c->F * (c->p.n_fft / 32);
  const long long blocks = (P.total + 255) / 256;
  const long long cap = static_cast<long long>(c->sm_count) * 32;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  {
    LaunchScope ls(c, ST_FOLD, st);
    if (P.x16 != nullptr) fold2_kernel<true><<<grid, 256, 0, st>>>(P);
    else fold2_kernel<false><<<grid, 256, 0, st>>>(P);
  }
  AVLD_CUDA(cudaGetLastError());
  return AVLD_OK;
}

}  // namespace avld
