// fma_peak.cu -- what the fp32 FMA pipe of this GPU sustains, by instruction form (the yardstick for conv1_kernel):
//   ffma      scalar FFMA, three register operands
//   ffma2     packed fma.rn.f32x2, three 64-bit register operands
//   ffma2_bc  packed FFMA2 whose first operand is one scalar broadcast to both halves (the form conv1_kernel issues)
// 16 independent accumulator chains per thread, 256-thread blocks, 8 blocks per SM.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_peak.bin tools/fma_peak.cu && tools/fma_peak.bin
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int kChains = 16;

template <int MODE>
__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float seed) {
  float2 acc[kChains], w[4];
  float s[4];
#pragma unroll
  for (int i = 0; i < kChains; ++i) acc[i] = make_float2(seed * (i + 1), seed * (i + 2) + threadIdx.x);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    w[i] = make_float2(1.0f + seed * (i + 1), 1.0f - seed * (i + 3));
    s[i] = 0.5f + seed * i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        if (MODE == 0) {                        // two scalar FFMAs = the flops of one FFMA2
          asm("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i].x) : "f"(w[r].x), "f"(s[r]));          // inline PTX: nvcc pairs plain fmaf calls into FFMA2
          asm("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i].y) : "f"(w[r].y), "f"(s[(r + 1) & 3]));
        } else if (MODE == 1) {
          acc[i] = fma2(acc[i], w[r], w[(r + 1) & 3]);
        } else {
          acc[i] = fma2(make_float2(s[r], s[r]), w[(r + i) & 3], acc[i]);
        }
      }
    }
  }
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) t += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
double run(const char* name, int sms, float* d_out) {
  const int blocks = sms * 8, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) fma_kernel<MODE><<<blocks, 256>>>(d_out, iters, 1e-7f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) fma_kernel<MODE><<<blocks, 256>>>(d_out, iters, 1e-7f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 2.0 * kChains * 4.0 * iters * 256.0 * blocks * reps;      // 2 flops x 2 lanes per FFMA2 (or FFMA pair)
  const double tf = flops / (ms * 1e-3) / 1e12;
  printf("%s\"%s_tflops\": %.2f", MODE ? ", " : "", name, tf);
  return tf;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* d_out;
  cudaMalloc(&d_out, sizeof(float) * p.multiProcessorCount * 8 * 256);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %d, \"nominal_tflops\": %.2f, ", p.name, p.multiProcessorCount, clk / 1000,
         p.multiProcessorCount * 128.0 * 2.0 * clk * 1e3 / 1e12);
  run<0>("ffma", p.multiProcessorCount, d_out);
  run<1>("ffma2", p.multiProcessorCount, d_out);
  run<2>("ffma2_bc", p.multiProcessorCount, d_out);
  printf(", \"method\": \"16 independent chains per thread, 256-thread blocks, 8 per SM, 4096 x 64 FFMA2 per thread, 20 timed launches (CUDA events), short burst: boost clock\"}\n");
  return cudaGetLastError() != cudaSuccess;
}
