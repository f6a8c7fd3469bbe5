// resample.cu -- M1: the resampling step of `librosa.load(path, sr=sr)` (map_detector_core.py:210,
// 00_normalize_dataset_rms.py:51) for files whose rate differs from `sr`: librosa 0.9.2 resample(res_type="kaiser_best")
// = resampy's band-limited sinc interpolation with a Kaiser-windowed filter table (64 zero crossings, 512 table samples per
// crossing, rolloff 0.9475937167399596, beta 14.769656459379492), linear interpolation between table entries, a left and
// a right filter wing per output sample, the output accumulated in float32 one tap at a time.  One thread per output
// sample walks its taps in the reference's order with the reference's roundings (float64 weight and product, float32 sum).
#include <cmath>
#include <vector>

#include "common.cuh"

namespace avld {

namespace {
constexpr int kNumZeros = 64, kPrecision = 9;
constexpr double kRolloff = 0.9475937167399596, kBeta = 14.769656459379492;

double bessel_i0(double x) {          // power series; x <= ~15 here
  long double term = 1.0L, sum = 1.0L;
  const long double q = static_cast<long double>(x) * x / 4.0L;
  for (int k = 1; k < 200; ++k) {
    term *= q / (static_cast<long double>(k) * k);
    sum += term;
    if (term < sum * 1e-22L) break;
  }
  return static_cast<double>(sum);
}
}  // namespace

struct ResampleParams {
  const float* x;
  float* y;
  const double* time_reg;   // [n_res] the reference's running `time_register` (repeated addition, done on the host)
  const double* win;        // [nwin] filter half-window (times the ratio when down-sampling)
  const double* delta;      // [nwin] forward differences
  long long n_in, n_res, n_out;
  int nwin, num_table, index_step;
  double scale;
};

__global__ void __launch_bounds__(256) resample_kernel(const ResampleParams P) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < P.n_out;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (t >= P.n_res) {                 // librosa's fix_length pads with zeros
      P.y[t] = 0.f;
      continue;
    }
    const double tr = P.time_reg[t];
    const long long n = static_cast<long long>(tr);
    double frac = P.scale * (tr - static_cast<double>(n));
    float acc = 0.f;
    {   // left wing: x[n], x[n-1], ...
      const double index_frac = frac * P.num_table;
      const int offset = static_cast<int>(index_frac);
      const double eta = index_frac - offset;
      const long long lim = (P.nwin - offset) / P.index_step;
      const long long i_max = n + 1 < lim ? n + 1 : lim;
      for (long long i = 0; i < i_max; ++i) {
        const int at = offset + static_cast<int>(i) * P.index_step;
        const double w = P.win[at] + eta * P.delta[at];
        acc = static_cast<float>(static_cast<double>(acc) + w * static_cast<double>(P.x[n - i]));
      }
    }
    {   // right wing: x[n+1], x[n+2], ...
      frac = P.scale - frac;
      const double index_frac = frac * P.num_table;
      const int offset = static_cast<int>(index_frac);
      const double eta = index_frac - offset;
      const long long lim = (P.nwin - offset) / P.index_step;
      const long long k_max = P.n_in - n - 1 < lim ? P.n_in - n - 1 : lim;
      for (long long k = 0; k < k_max; ++k) {
        const int at = offset + static_cast<int>(k) * P.index_step;
        const double w = P.win[at] + eta * P.delta[at];
        acc = static_cast<float>(static_cast<double>(acc) + w * static_cast<double>(P.x[n + k + 1]));
      }
    }
    P.y[t] = acc;
  }
}

}  // namespace avld

using namespace avld;

extern "C" int64_t avld_resample_len(int64_t n_in, int32_t sr_in, int32_t sr_out) {
  if (n_in < 0 || sr_in <= 0 || sr_out <= 0) return -1;
  const double ratio = static_cast<double>(sr_out) / static_cast<double>(sr_in);     // librosa: float(target_sr) / orig_sr
  return static_cast<int64_t>(std::ceil(static_cast<double>(n_in) * ratio));
}

extern "C" int avld_resample(avld_ctx* c, const float* x, int64_t n_in, int32_t sr_in, int32_t sr_out, float* y, int64_t n_out,
                             void* stream) {
  AVLD_ENTER(c);
  AVLD_CHECK(sr_in > 0 && sr_out > 0 && n_in >= 0, AVLD_ERR_INVALID, "bad rates / length");
  AVLD_CHECK(n_out == avld_resample_len(n_in, sr_in, sr_out), AVLD_ERR_INVALID, "n_out must be avld_resample_len(n_in, sr_in, sr_out)");
  if (n_out == 0) return AVLD_OK;
  AVLD_CHECK(x && y, AVLD_ERR_INVALID, "NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double ratio = static_cast<double>(sr_out) / static_cast<double>(sr_in);
  const int num_table = 1 << kPrecision, n_tab = num_table * kNumZeros, nwin = n_tab + 1;
  if (c->rs_sr_in != sr_in || c->rs_sr_out != sr_out) {        // filter table of this rate pair (tiny: 2 x 256 KB)
    std::vector<double> win(nwin), delta(nwin, 0.0);
    const double i0b = bessel_i0(kBeta);
    for (int j = 0; j < nwin; ++j) {
      const double tt = static_cast<double>(j) * (static_cast<double>(kNumZeros) / n_tab);   // np.linspace(0, 64, n + 1)
      const double v = kRolloff * tt;
      const double sinc = v == 0.0 ? 1.0 : std::sin(M_PI * v) / (M_PI * v);                  // np.sinc
      const double r = static_cast<double>(j) / n_tab;                                       // kaiser(2 n + 1, beta)[n + j]
      const double taper = bessel_i0(kBeta * std::sqrt(std::fmax(0.0, 1.0 - r * r))) / i0b;
      win[j] = taper * (kRolloff * sinc);
      if (ratio < 1.0) win[j] *= ratio;
    }
    for (int j = 0; j + 1 < nwin; ++j) delta[j] = win[j + 1] - win[j];
    if (!c->d_rs_win) {
      AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_rs_win), nwin * sizeof(double)));
      AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_rs_delta), nwin * sizeof(double)));
    }
    AVLD_CUDA(cudaMemcpyAsync(c->d_rs_win, win.data(), nwin * sizeof(double), cudaMemcpyHostToDevice, st));
    AVLD_CUDA(cudaMemcpyAsync(c->d_rs_delta, delta.data(), nwin * sizeof(double), cudaMemcpyHostToDevice, st));
    AVLD_CUDA(cudaStreamSynchronize(st));     // the host vectors go out of scope
    c->rs_sr_in = sr_in;
    c->rs_sr_out = sr_out;
  }
  ResampleParams P{};
  P.n_in = n_in;
  P.n_res = static_cast<long long>(static_cast<double>(n_in) * ratio);     // resampy: int(shape * sample_ratio)
  P.n_out = n_out;
  if (P.n_res > n_out) P.n_res = n_out;
  // the reference advances `time_register += time_increment` once per output sample: reproduce the sequence of sums
  const double inc = 1.0 / ratio;
  std::vector<double> tr(static_cast<size_t>(P.n_res));
  double acc = 0.0;
  for (long long i = 0; i < P.n_res; ++i) {
    tr[i] = acc;
    acc += inc;
  }
  double* d_tr = nullptr;
  AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_tr), std::max<size_t>(tr.size(), 1) * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(d_tr, tr.data(), tr.size() * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    P.x = x;
    P.y = y;
    P.time_reg = d_tr;
    P.win = c->d_rs_win;
    P.delta = c->d_rs_delta;
    P.nwin = nwin;
    P.num_table = num_table;
    P.scale = ratio < 1.0 ? ratio : 1.0;
    P.index_step = static_cast<int>(P.scale * num_table);
    const long long blocks = (n_out + 255) / 256;
    const int grid = static_cast<int>(blocks < static_cast<long long>(c->sm_count) * 16 ? blocks : static_cast<long long>(c->sm_count) * 16);
    { LaunchScope ls(c, ST_ELEMENTWISE, st); resample_kernel<<<grid, 256, 0, st>>>(P); }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // `tr` and d_tr are released below
  }
  cudaFree(d_tr);
  AVLD_CHECK(e == cudaSuccess, AVLD_ERR_CUDA, "avld_resample: %s", cudaGetErrorString(e));
  return AVLD_OK;
}
