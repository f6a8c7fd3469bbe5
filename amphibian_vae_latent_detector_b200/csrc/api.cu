// api.cu -- whole-path entry points: device-resident encode, host-buffer encode+detect with
// double-buffered copies, and the GEMM bring-up entry used by the tests.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "gemm3.cuh"

using namespace avld;

// `input_consumed` (optional) is recorded once the audio buffer has been read for the last time: the operand kernel is its
// last reader (it reads the chunk itself in passes without the PCM_16 round trip, prep_kernel's integers otherwise)
static int encode_pass(avld_ctx* c, const float* x, const int16_t* x16, float* mu, uint8_t* ok, int m, float target_rms,
                       float rms_min, float eps, int quantize, cudaStream_t st, cudaEvent_t input_consumed = nullptr) {
  AVLD_CHECK(c->features_ok, AVLD_ERR_UNSUPPORTED, "chunk_len %d is outside the feature kernels' range", c->L);
  AVLD_TRY(launch_prep(c, x, x16, nullptr, true, true, ok, nullptr, m, target_rms, rms_min, eps, quantize, st));
  AVLD_TRY(launch_fold3(c, m, st));
  if (input_consumed) AVLD_CUDA(cudaEventRecord(input_consumed, st));
  AVLD_TRY(launch_dftf3(c, m, st));
  AVLD_TRY(launch_logmel_post(c, c->d_feat, m, st));
  AVLD_TRY(launch_encoder(c, c->d_feat, mu, m, st));
  return AVLD_OK;
}

extern "C" int avld_encode(avld_ctx* c, const float* x, float* mu, uint8_t* ok, int64_t n, float target_rms,
                           float rms_min, float eps, int quantize_pcm16, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(x && mu, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0, AVLD_ERR_INVALID, "negative n");
  AVLD_CHECK(!c->ops.empty(), AVLD_ERR_STATE, "avld_encoder_load has not been called");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t i = 0; i < n; i += c->max_batch) {
    const int m = static_cast<int>(std::min<int64_t>(n - i, c->max_batch));
    AVLD_TRY(encode_pass(c, x + i * c->L, nullptr, mu + i * c->latent_dim, ok ? ok + i : nullptr, m, target_rms, rms_min,
                         eps, quantize_pcm16, st));
  }
  return AVLD_OK;
}

extern "C" int avld_encode_pcm16(avld_ctx* c, const int16_t* pcm, float* mu, uint8_t* ok, int64_t n, float target_rms,
                                 float rms_min, float eps, int quantize_pcm16, void* stream) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(pcm && mu, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0, AVLD_ERR_INVALID, "negative n");
  AVLD_CHECK(!c->ops.empty(), AVLD_ERR_STATE, "avld_encoder_load has not been called");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t i = 0; i < n; i += c->max_batch) {
    const int m = static_cast<int>(std::min<int64_t>(n - i, c->max_batch));
    AVLD_TRY(encode_pass(c, nullptr, pcm + i * c->L, mu + i * c->latent_dim, ok ? ok + i : nullptr, m, target_rms, rms_min,
                         eps, quantize_pcm16, st));
  }
  return AVLD_OK;
}

static int encode_detect_host_impl(avld_ctx* c, const void* x_host, int sample_bytes, int64_t n, int quantize_pcm16,
                                   const float* centroid, const double* thr, const int32_t* priority_rank, int32_t K,
                                   int32_t* pred_host, float* best_host, float* mu_host, uint8_t* ok_host) {
  AVLD_ENTER(c);
  if (n == 0) return AVLD_OK;                 // an empty batch is valid (and its pointers may be NULL)
  AVLD_CHECK(x_host && centroid && thr && priority_rank && pred_host && best_host, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(n >= 0 && K >= 1 && K <= 64, AVLD_ERR_INVALID, "bad n / K");
  AVLD_CHECK(!c->ops.empty(), AVLD_ERR_STATE, "avld_encoder_load has not been called");
  const int D = c->latent_dim;
  const size_t xbytes = static_cast<size_t>(c->max_batch) * c->L * sizeof(float);
  for (int b = 0; b < 2; ++b)
    if (!c->d_xbuf[b]) AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_xbuf[b]), xbytes));
  if (!c->d_cent) {
    AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_cent), 64 * 4096 * sizeof(float)));
    AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_thr), 64 * sizeof(double)));
    AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_prio), 64 * sizeof(int32_t)));
  }
  AVLD_CHECK(static_cast<size_t>(K) * D <= 64 * 4096, AVLD_ERR_UNSUPPORTED, "K*D too large");
  // pinned staging: pred | best | mu | ok
  const size_t n_sz = static_cast<size_t>(n);
  const size_t off_best = n_sz * 4, off_mu = off_best + n_sz * 4, off_ok = off_mu + (mu_host ? n_sz * D * 4 : 0);
  const size_t need = off_ok + n_sz + 64;
  if (need > c->h_stage_bytes) {
    if (c->h_stage) AVLD_CUDA(cudaFreeHost(c->h_stage));
    c->h_stage = nullptr;
    c->h_stage_bytes = 0;
    AVLD_CUDA(cudaHostAlloc(&c->h_stage, need, cudaHostAllocDefault));
    c->h_stage_bytes = need;
  }
  char* hs = static_cast<char*>(c->h_stage);
  int32_t* s_pred = reinterpret_cast<int32_t*>(hs);
  float* s_best = reinterpret_cast<float*>(hs + off_best);
  float* s_mu = reinterpret_cast<float*>(hs + off_mu);
  uint8_t* s_ok = reinterpret_cast<uint8_t*>(hs + off_ok);
  cudaStream_t sc = c->s_compute, sx = c->s_copy;
  AVLD_CUDA(cudaMemcpyAsync(c->d_cent, centroid, static_cast<size_t>(K) * D * sizeof(float), cudaMemcpyHostToDevice, sc));
  AVLD_CUDA(cudaMemcpyAsync(c->d_thr, thr, K * sizeof(double), cudaMemcpyHostToDevice, sc));
  AVLD_CUDA(cudaMemcpyAsync(c->d_prio, priority_rank, K * sizeof(int32_t), cudaMemcpyHostToDevice, sc));

  // AVLD_HOST_TRACE=<file> (read once, at context creation): per-slab device timeline (copy start/end, compute start/end,
  // ms since the first copy)
  const char* trace_path = c->host_trace_path.empty() ? nullptr : c->host_trace_path.c_str();
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t s) {
    if (!trace_path) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    tev.push_back(e);
  };

  // slab i: H2D on the copy stream into buffer i&1 while the compute stream works on slab i-1;
  // results of slab i are read back on the compute stream right after its kernels.
  // Slab sizes.  The copies run back to back (a slab buffer is free again as soon as the operand kernel has read it), so
  // the call takes the copy time plus whatever is left to compute when the last byte has landed: the last slab's kernels,
  // and any backlog in front of them.  Full passes first, then a ramp down to a quarter pass in steps the kernels keep up
  // with (a pass computes ~1.6x faster than it copies, so each slab is done before the next, <= 1.67x smaller one has
  // arrived); below a quarter pass the kernels lose efficiency faster than the tail shrinks (measured: tools/host_trace_probe.py).
  std::vector<int> sizes;
  {
    const int64_t mb = c->max_batch;
    std::vector<int> ramp;
    int64_t left = n;
    if (n >= 3 * mb && mb >= 8)
      for (int64_t sz : {mb * 5 / 8, mb * 3 / 8, mb / 4}) {
        ramp.push_back(static_cast<int>(sz));
        left -= sz;
      }
    while (left > 0) {
      const int64_t m = std::min(left, mb);
      sizes.push_back(static_cast<int>(m));
      left -= m;
    }
    if (!ramp.empty() && sizes.size() >= 2 && sizes.back() < mb) std::swap(sizes.back(), sizes.front());   // the odd remainder goes first
    sizes.insert(sizes.end(), ramp.begin(), ramp.end());
  }
  // one slab; any failure leaves copies / kernels in flight on both streams, so the caller drains them before returning
  auto run_slab = [&](int64_t slab, int64_t i) -> int {
    const int m = sizes[slab];
    const int b = static_cast<int>(slab & 1);
    if (slab >= 2) AVLD_CUDA(cudaStreamWaitEvent(sx, c->ev_done[b], 0));   // buffer b free again
    mark(sx);
    AVLD_CUDA(cudaMemcpyAsync(c->d_xbuf[b], static_cast<const char*>(x_host) + static_cast<size_t>(i) * c->L * sample_bytes,
                              static_cast<size_t>(m) * c->L * sample_bytes, cudaMemcpyHostToDevice, sx));
    AVLD_CUDA(cudaEventRecord(c->ev_h2d[b], sx));
    mark(sx);
    AVLD_CUDA(cudaStreamWaitEvent(sc, c->ev_h2d[b], 0));
    mark(sc);
    AVLD_TRY(encode_pass(c, sample_bytes == 4 ? c->d_xbuf[b] : nullptr,
                         sample_bytes == 2 ? reinterpret_cast<const int16_t*>(c->d_xbuf[b]) : nullptr, c->d_mu, c->d_ok, m,
                         static_cast<float>(c->norm_target), static_cast<float>(c->norm_rms_min),
                         static_cast<float>(c->norm_eps), quantize_pcm16, sc, c->ev_done[b]));   // buffer b is free after the operand kernel
    AVLD_TRY(avld_radii(c, c->d_mu, c->d_cent, c->d_radii, m, K, D, sc));
    AVLD_TRY(avld_decide(c, c->d_radii, c->d_thr, c->d_prio, c->d_pred, c->d_best, m, K, sc));
    AVLD_CUDA(cudaMemcpyAsync(s_pred + i, c->d_pred, static_cast<size_t>(m) * sizeof(int32_t), cudaMemcpyDeviceToHost, sc));
    AVLD_CUDA(cudaMemcpyAsync(s_best + i, c->d_best, static_cast<size_t>(m) * sizeof(float), cudaMemcpyDeviceToHost, sc));
    if (mu_host)
      AVLD_CUDA(cudaMemcpyAsync(s_mu + i * D, c->d_mu, static_cast<size_t>(m) * D * sizeof(float), cudaMemcpyDeviceToHost, sc));
    if (ok_host)
      AVLD_CUDA(cudaMemcpyAsync(s_ok + i, c->d_ok, static_cast<size_t>(m), cudaMemcpyDeviceToHost, sc));
    mark(sc);
    return AVLD_OK;
  };
  int rc = AVLD_OK;
  {
    int64_t i = 0;
    for (int64_t slab = 0; slab < static_cast<int64_t>(sizes.size()) && rc == AVLD_OK; i += sizes[slab], ++slab) rc = run_slab(slab, i);
  }
  const cudaError_t e_sc = cudaStreamSynchronize(sc), e_sx = cudaStreamSynchronize(sx);
  if (rc != AVLD_OK || e_sc != cudaSuccess || e_sx != cudaSuccess) {
    // nothing of this call is in flight any more: the next call may reuse the slab buffers, d_cent and the staging area
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    if (rc == AVLD_OK) {
      set_error("avld_encode_detect_host: %s", cudaGetErrorString(e_sc != cudaSuccess ? e_sc : e_sx));
      rc = AVLD_ERR_CUDA;
    }
    return rc;
  }
  if (trace_path && !tev.empty()) {
    if (FILE* f = fopen(trace_path, "w")) {
      fprintf(f, "slab copy_start copy_end compute_start compute_end (ms)\n");
      for (size_t k = 0; k + 3 < tev.size(); k += 4) {
        float t[4];
        for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[k + j]);
        fprintf(f, "%zu %.3f %.3f %.3f %.3f\n", k / 4, t[0], t[1], t[2], t[3]);
      }
      fclose(f);
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  memcpy(pred_host, s_pred, n_sz * sizeof(int32_t));
  memcpy(best_host, s_best, n_sz * sizeof(float));
  if (mu_host) memcpy(mu_host, s_mu, n_sz * D * sizeof(float));
  if (ok_host) memcpy(ok_host, s_ok, n_sz);
  return AVLD_OK;
}

extern "C" int avld_encode_detect_host(avld_ctx* c, const float* x_host, int64_t n, int quantize_pcm16,
                                       const float* centroid, const double* thr, const int32_t* priority_rank,
                                       int32_t K, int32_t* pred_host, float* best_host, float* mu_host,
                                       uint8_t* ok_host) {
  return encode_detect_host_impl(c, x_host, 4, n, quantize_pcm16, centroid, thr, priority_rank, K, pred_host, best_host,
                                 mu_host, ok_host);
}

extern "C" int avld_encode_detect_host_pcm16(avld_ctx* c, const int16_t* pcm_host, int64_t n, int quantize_pcm16,
                                             const float* centroid, const double* thr, const int32_t* priority_rank,
                                             int32_t K, int32_t* pred_host, float* best_host, float* mu_host,
                                             uint8_t* ok_host) {
  return encode_detect_host_impl(c, pcm_host, 2, n, quantize_pcm16, centroid, thr, priority_rank, K, pred_host, best_host,
                                 mu_host, ok_host);
}

// ------------------------------------------------------------------------------------------------
// bring-up / test entry: C = A * B^T through the tcgen05 core with plain row-major operands
// ------------------------------------------------------------------------------------------------
extern "C" int avld_dbg_gemm(avld_ctx* c, const float* A, const float* B, float* C, int64_t M, int32_t N, int32_t K,
                             int32_t mode, void* stream) {
  AVLD_ENTER(c);
  AVLD_CHECK(A && B && C, AVLD_ERR_INVALID, "NULL argument");
  AVLD_CHECK(M >= 1 && N >= 16 && N % 16 == 0 && K >= 64 && K % 64 == 0, AVLD_ERR_INVALID, "need N %% 16 == 0 and K %% 64 == 0");
  AVLD_CHECK(mode == 0 || mode == 1, AVLD_ERR_INVALID, "mode must be 0 or 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  const size_t na = static_cast<size_t>(M) * K, nb = static_cast<size_t>(N) * K;
  AVLD_CUDA(cudaMalloc(&a_hi, na * 2));
  AVLD_CUDA(cudaMalloc(&a_lo, na * 2));
  AVLD_CUDA(cudaMalloc(&b_hi, nb * 2));
  AVLD_CUDA(cudaMalloc(&b_lo, nb * 2));
  int rc = AVLD_OK;
  do {
    CUtensorMapDataType hi_t = mode == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (mode == 0) {
      if ((rc = launch_split_f16(A, static_cast<__half*>(a_hi), static_cast<__half*>(a_lo), na, st))) break;
      if ((rc = launch_split_f16(B, static_cast<__half*>(b_hi), static_cast<__half*>(b_lo), nb, st))) break;
    } else {
      if ((rc = launch_split_bf16(A, static_cast<__nv_bfloat16*>(a_hi), static_cast<__nv_bfloat16*>(a_lo), na, st))) break;
      if ((rc = launch_split_bf16(B, static_cast<__nv_bfloat16*>(b_hi), static_cast<__nv_bfloat16*>(b_lo), nb, st))) break;
    }
    const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if ((rc = encode_tmap_2d(&ta_hi, a_hi, hi_t, K, M, static_cast<uint64_t>(K) * 2, 64, 128, 128))) break;
    if ((rc = encode_tmap_2d(&ta_lo, a_lo, hi_t, K, M, static_cast<uint64_t>(K) * 2, 64, 128, 128))) break;
    if ((rc = encode_tmap_2d(&tb_hi, b_hi, hi_t, K, N, static_cast<uint64_t>(K) * 2, 64, bn, 128))) break;
    if ((rc = encode_tmap_2d(&tb_lo, b_lo, hi_t, K, N, static_cast<uint64_t>(K) * 2, 64, bn, 128))) break;
    Gemm3Params P{};
    P.num_m_tiles = static_cast<int>((M + 127) / 128);
    P.num_n_tiles = (N + bn - 1) / bn;
    P.num_k_blocks = K / 64;
    const int hb = mode == 0 ? 0 : 1;
    P.idesc_hh = P.idesc_lh = P.idesc_hl = avld_make_idesc(hb, hb, 128, bn);
    P.a_mode = 0;
#ifdef AVLD_BRINGUP
    if (const char* e = getenv("AVLD_DBG_SHIFT")) P.dbg_shift = atoi(e);      // tools/probe_shift.py (libavld_bringup.so)
    if (const char* e = getenv("AVLD_DBG_BASEOFF")) P.dbg_baseoff = atoi(e);
#endif
    P.M_total = M;
    P.N_total = N;
    P.out_f32 = C;
    P.ldc = N;
    LaunchScope ls(c, ST_DENSE_GEMM, st);
    rc = run_gemm3(c, bn, 128, EPI_PLAIN, ta_hi, ta_lo, tb_hi, tb_lo, P, st);
    if (rc) break;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("dbg_gemm kernel failed: %s", cudaGetErrorString(e));
      rc = AVLD_ERR_CUDA;
    }
  } while (0);
  cudaFree(a_hi);
  cudaFree(a_lo);
  cudaFree(b_hi);
  cudaFree(b_lo);
  return rc;
}
