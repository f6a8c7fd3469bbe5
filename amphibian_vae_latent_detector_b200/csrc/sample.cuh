// sample.cuh -- the per-sample tail of rms_normalize + the PCM_16 round trip, shared by the kernels that re-apply the
// normalisation on the fly.  Every operation is explicitly rounded (__fmul_rn), so the result does not depend on the
// translation unit's -fmad setting and stays bit-identical to numpy (00_normalize_dataset_rms.py:36-37, :57).
#pragma once
#include <cuda_runtime.h>

namespace avld {

__device__ __forceinline__ float finish_sample(float v, float scale, int scaled, int quantize) {
  if (scaled) {
    v = __fmul_rn(v, scale);
    v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);   // np.clip keeps NaN
  }
  if (quantize) {   // sf.write PCM_16 (lrintf(x * 0x7FFF)) + librosa.load (s / 0x8000); through int so that -0.0 -> +0.0
    int q = __float2int_rn(__fmul_rn(v, 32767.0f));
    q = q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
    v = __fmul_rn(static_cast<float>(q), 1.0f / 32768.0f);
  }
  return v;
}

}  // namespace avld
