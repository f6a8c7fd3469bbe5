// gemm3.cu -- the one translation unit that instantiates the tcgen05 GEMM core.
#define AVLD_GEMM3_IMPL
#include "gemm3.cuh"

namespace avld {

template <int EPI>
static int by_shape(avld_ctx* c, int bn, int swz, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                    const CUtensorMap& b_lo, const Gemm3Params& P, cudaStream_t st) {
  if (swz == 128) {
    if (bn == 32 && EPI == EPI_PLAIN) return launch_gemm3<32, 128, EPI_PLAIN>(c, a_hi, a_lo, b_hi, b_lo, P, st);
    if (bn == 64) return launch_gemm3<64, 128, EPI>(c, a_hi, a_lo, b_hi, b_lo, P, st);
    if (bn == 128) return launch_gemm3<128, 128, EPI>(c, a_hi, a_lo, b_hi, b_lo, P, st);
    if (bn == 256) return launch_gemm3<256, 128, EPI>(c, a_hi, a_lo, b_hi, b_lo, P, st);
  } else if (swz == 64 && EPI == EPI_CONV) {
    if (bn == 64) return launch_gemm3<64, 64, EPI_CONV>(c, a_hi, a_lo, b_hi, b_lo, P, st);
    if (bn == 128) return launch_gemm3<128, 64, EPI_CONV>(c, a_hi, a_lo, b_hi, b_lo, P, st);
    if (bn == 256) return launch_gemm3<256, 64, EPI_CONV>(c, a_hi, a_lo, b_hi, b_lo, P, st);
  }
  set_error("run_gemm3: no kernel for BN=%d swizzle=%d epilogue=%d", bn, swz, EPI);
  return AVLD_ERR_UNSUPPORTED;
}

int run_gemm3(avld_ctx* c, int bn, int swz, int epi, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
              const CUtensorMap& b_lo, const Gemm3Params& P, cudaStream_t st) {
  if (epi == EPI_PLAIN) return by_shape<EPI_PLAIN>(c, bn, swz, a_hi, a_lo, b_hi, b_lo, P, st);
  if (epi == EPI_CONV) return by_shape<EPI_CONV>(c, bn, swz, a_hi, a_lo, b_hi, b_lo, P, st);
  set_error("run_gemm3: no kernel for BN=%d swizzle=%d epilogue=%d", bn, swz, epi);
  return AVLD_ERR_UNSUPPORTED;
}

}  // namespace avld
