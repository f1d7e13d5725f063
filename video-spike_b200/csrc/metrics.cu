// Evaluation metrics on the device (E2): the per-epoch loops of src/utils/utils.py:122-181 over
// src/utils/metric_utils.py:36-102 (bits per spike) and sklearn's r2_score, without copying the predictions to the host.
//   bps[n]     = (NLL_null[n] - NLL_model[n]) / sum spikes[n] / ln 2,  NLL = sum (r - s log r + lgamma(s + 1)),
//                null rate = mean spike count of neuron n; rates equal to 0 are replaced by 1e-9 (metric_utils.py:68-73);
//                bins whose spike count is NaN are masked out (metric_utils.py:57-60)
//   r2[k][t]   = sklearn r2_score over the N neurons of trial k, time bin t (utils.py:158: y_true = gt[:, :, k] is (N, T),
//                i.e. samples = neurons, outputs = time bins; the host wrapper averages over t like multioutput=
//                'uniform_average'); SS_tot == 0 gives 1.0 if SS_res == 0 else 0.0 like sklearn
// All sums in fp64, fixed order (lanes over neurons, warps over rows, warps combined sequentially).
#include "common.cuh"

namespace vs {
namespace metrics {

// block = 32 neurons (lanes) x 8 warps striding over the K*T rows
__global__ void __launch_bounds__(256) bps_kernel(const float* __restrict__ rates, const float* __restrict__ spikes, long long rows,
                                                  long long N, double* __restrict__ bps) {
  __shared__ double red[8][32][5];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long n = (long long)blockIdx.x * 32 + lane;
  double sr = 0.0, sslogr = 0.0, ss = 0.0, cnt = 0.0, slg = 0.0;
  if (n < N) {
    for (long long row = w; row < rows; row += 8) {
      const double s = (double)spikes[row * N + n];
      if (isnan(s)) continue;
      double r = (double)rates[row * N + n];
      if (r == 0.0) r = 1e-9;
      sr += r;
      sslogr += s * log(r);
      slg += lgamma(s + 1.0);
      ss += s;
      cnt += 1.0;
    }
  }
  red[w][lane][0] = sr; red[w][lane][1] = sslogr; red[w][lane][2] = ss; red[w][lane][3] = cnt; red[w][lane][4] = slg;
  __syncthreads();
  if (w == 0 && n < N) {
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int w2 = 0; w2 < 8; ++w2)
      for (int e = 0; e < 5; ++e) a[e] += red[w2][lane][e];
    const double nll_model = a[0] - a[1] + a[4];
    double null_rate = a[2] / a[3];
    if (null_rate == 0.0) null_rate = 1e-9;
    const double nll_null = a[3] * null_rate - a[2] * log(null_rate) + a[4];
    bps[n] = (nll_null - nll_model) / a[2] / 0.6931471805599453;
  }
}

// one warp per (k, t) row of N neurons
__global__ void __launch_bounds__(256) r2_rows_kernel(const float* __restrict__ gt, const float* __restrict__ pred, long long rows,
                                                      long long N, double* __restrict__ r2) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* g = gt + row * N;
  const float* p = pred + row * N;
  double sg = 0.0;
  for (long long n = lane; n < N; n += 32) sg += (double)g[n];
  sg = warp_sum(sg);
  const double mean = sg / (double)N;
  double res = 0.0, tot = 0.0;
  for (long long n = lane; n < N; n += 32) {
    const double y = (double)g[n], d = y - (double)p[n], c = y - mean;
    res = fma(d, d, res);
    tot = fma(c, c, tot);
  }
  res = warp_sum(res);
  tot = warp_sum(tot);
  if (lane == 0) r2[row] = tot != 0.0 ? 1.0 - res / tot : (res == 0.0 ? 1.0 : 0.0);
}

}  // namespace metrics
}  // namespace vs

using namespace vs;

extern "C" int vs_bits_per_spike(const float* rates, const float* spikes, int64_t K, int64_t T, int64_t N, double* bps_n, void* stream) {
  VS_REQUIRE(rates && spikes && bps_n && K > 0 && T > 0 && N > 0, VS_ERR_INVALID, "vs_bits_per_spike: bad arguments");
  VS_LAUNCH(metrics::bps_kernel, (unsigned)ceil_div(N, 32), 256, 0, stream, rates, spikes, (long long)(K * T), (long long)N, bps_n);
  return VS_OK;
}

extern "C" int vs_r2_rows(const float* gt, const float* pred, int64_t K, int64_t T, int64_t N, double* r2_kt, void* stream) {
  VS_REQUIRE(gt && pred && r2_kt && K > 0 && T > 0 && N > 0, VS_ERR_INVALID, "vs_r2_rows: bad arguments");
  VS_LAUNCH(metrics::r2_rows_kernel, (unsigned)ceil_div(K * T, 8), 256, 0, stream, gt, pred, (long long)(K * T), (long long)N, r2_kt);
  return VS_OK;
}
