// sample.cuh -- the per-sample tail of rms_normalize + the PCM_16 round trip, shared by the kernels that re-apply the
// normalisation on the fly.  Every operation is explicitly rounded (__fmul_rn), so the result does not depend on the
// translation unit's -fmad setting and stays bit-identical to numpy (00_normalize_dataset_rms.py:36-37, :57).
#pragma once
#include <cuda_runtime.h>

namespace avld {

__device__ __forceinline__ float scale_clip(float v, float scale, int scaled) {
  if (scaled) {
    v = __fmul_rn(v, scale);
    v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);   // np.clip keeps NaN
  }
  return v;
}

// sf.write PCM_16: lrintf(x * 0x7FFF), saturated
__device__ __forceinline__ int pcm16_round(float v) {
  int q = __float2int_rn(__fmul_rn(v, 32767.0f));
  return q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
}

// The same integer as pcm16_round(scale_clip(v, scale, scaled)) in four branch-free instructions.  `scale` is 1.0 for a
// chunk that is not scaled (v * 1.0 = v); the float -> s16 conversion saturates by itself and maps NaN to 0, so the clip to
// [-1, 1] only has to be repaired on the negative side (clip gives -32767 where saturation alone would give -32768).
__device__ __forceinline__ int pcm16_of(float v, float scale, int scaled) {
  short q;
  asm("cvt.rni.s16.f32 %0, %1;" : "=h"(q) : "f"(__fmul_rn(__fmul_rn(v, scale), 32767.0f)));
  const int lo = scaled ? -32767 : -32768;
  return q < lo ? lo : q;
}

__device__ __forceinline__ float finish_sample(float v, float scale, int scaled, int quantize) {
  v = scale_clip(v, scale, scaled);
  if (quantize)    // sf.write PCM_16 + librosa.load (s / 0x8000); through int so that -0.0 -> +0.0
    v = __fmul_rn(static_cast<float>(pcm16_round(v)), 1.0f / 32768.0f);
  return v;
}

}  // namespace avld
