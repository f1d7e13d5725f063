// L-BFGS vector kernels (R5): the full-batch optimiser of src/model/rrr.py:177,199
// (`torch.optim.LBFGS(...).step(closure)`, no line search) on flat fp64 parameter / gradient vectors.
//
// torch's two-loop recursion walks the history twice with a dot and an axpy per pair (4 kernel launches and
// ~5 vector passes per pair and iteration).  Here the recursion runs in COEFFICIENT space on the host: the
// direction is a linear combination of {g, s_i, y_i}, and everything the recursion needs is the Gram matrix of
// those vectors (host side: optim.py, FusedLBFGS).  The device therefore does exactly two streaming passes per
// iteration:
//   vs_lbfgs_dots       one pass over g, g_prev, s_new and the 2m history vectors: writes y_new = g - g_prev
//                       and every inner product the Gram update needs, plus |g|_inf and |g|_1
//   vs_lbfgs_direction  one pass: d = cg*g + sum cs_i s_i + cy_i y_i;  s_out = t*d;  x += t*d;  max|t*d|
// Both are HBM-bound (8 B/element/vector).  Reductions are two-stage with a fixed order: bit-reproducible.
#include "common.cuh"

namespace vs {
namespace lbfgs {

constexpr int kHG = 8;        // history vectors handled by one block row of the dots kernel
constexpr int kBase = 8;      // base scalars: g.g, |g|_1, |g|_inf, y.y, y.s, s.g, y.g, (unused)
constexpr int kMaxChunks = 1024;

struct Slots { int32_t s[2 * VS_LBFGS_MAX_HIST]; };              // element offsets / stride of the 2m vectors
struct Coef { double c[2 * VS_LBFGS_MAX_HIST + 1]; };            // cg, then one coefficient per history vector

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
  return s;  // valid in thread 0
}
__device__ __forceinline__ double block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = fmax(s, sh[w]);
  return s;
}

// grid (chunks, groups).  Block (c, q) covers elements [c*per, (c+1)*per) and history vectors [q*kHG, (q+1)*kHG).
// Row q == 0 also produces the base scalars and writes y_new.  part[(c * nout) + o].
__global__ void __launch_bounds__(256) dots_kernel(long long n, long long per, const double* __restrict__ g,
                                                   const double* __restrict__ gp, const double* __restrict__ s_new,
                                                   double* __restrict__ y_out, const double* __restrict__ hist, long long stride,
                                                   const Slots slots, int nh, int nout, double* __restrict__ part) {
  __shared__ double sh[8];
  const long long c = blockIdx.x;
  const int q = blockIdx.y;
  const long long i0 = c * per, i1 = min(n, i0 + per);
  const int h0 = q * kHG;
  const double* hp[kHG];
#pragma unroll
  for (int h = 0; h < kHG; ++h) hp[h] = (h0 + h < nh) ? hist + (long long)slots.s[h0 + h] * stride : nullptr;
  double acc[kHG][3];
#pragma unroll
  for (int h = 0; h < kHG; ++h) acc[h][0] = acc[h][1] = acc[h][2] = 0.0;
  double gg = 0.0, g1 = 0.0, gm = 0.0, yy = 0.0, ys = 0.0, sg = 0.0, yg = 0.0;
  const bool base = (q == 0);
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const double gv = g[i];
    const double pv = gp ? gp[i] : 0.0;
    const double sv = s_new ? s_new[i] : 0.0;
    const double yv = gv - pv;
    if (base) {
      gg = fma(gv, gv, gg); g1 += fabs(gv); gm = fmax(gm, fabs(gv));
      yy = fma(yv, yv, yy); ys = fma(yv, sv, ys); sg = fma(sv, gv, sg); yg = fma(yv, gv, yg);
      if (y_out) y_out[i] = yv;
    }
#pragma unroll
    for (int h = 0; h < kHG; ++h) {
      if (hp[h]) {
        const double hv = hp[h][i];
        acc[h][0] = fma(hv, gv, acc[h][0]);
        acc[h][1] = fma(hv, yv, acc[h][1]);
        acc[h][2] = fma(hv, sv, acc[h][2]);
      }
    }
  }
  double* out = part + c * nout;
  if (base) {
    double v;
    v = block_sum(gg, sh); if (threadIdx.x == 0) out[0] = v;
    v = block_sum(g1, sh); if (threadIdx.x == 0) out[1] = v;
    v = block_max(gm, sh); if (threadIdx.x == 0) out[2] = v;
    v = block_sum(yy, sh); if (threadIdx.x == 0) out[3] = v;
    v = block_sum(ys, sh); if (threadIdx.x == 0) out[4] = v;
    v = block_sum(sg, sh); if (threadIdx.x == 0) out[5] = v;
    v = block_sum(yg, sh); if (threadIdx.x == 0) out[6] = v;
    if (threadIdx.x == 0) out[7] = 0.0;
  }
#pragma unroll
  for (int h = 0; h < kHG; ++h) {
    if (h0 + h < nh) {   // uniform across the block
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const double v = block_sum(acc[h][e], sh);
        if (threadIdx.x == 0) out[kBase + 3 * (h0 + h) + e] = v;
      }
    }
  }
}

// out[o] = ordered sum (max for o == 2) over the chunk partials
__global__ void dots_reduce_kernel(const double* __restrict__ part, int chunks, int nout, double* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nout) return;
  double s = 0.0;
  if (o == 2) {
    for (int c = 0; c < chunks; ++c) s = fmax(s, part[(long long)c * nout + o]);
  } else {
    for (int c = 0; c < chunks; ++c) s += part[(long long)c * nout + o];
  }
  out[o] = s;
}

__global__ void __launch_bounds__(256) direction_kernel(long long n, const double* __restrict__ g, const double* __restrict__ hist,
                                                        long long stride, const Slots slots, int nh, const Coef coef, double t,
                                                        double* __restrict__ x, double* __restrict__ s_out,
                                                        unsigned long long* __restrict__ dmax_bits) {
  __shared__ double sh[8];
  double m = 0.0;
  const long long step = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += step) {
    double d = coef.c[0] * g[i];
#pragma unroll 8
    for (int h = 0; h < nh; ++h) d = fma(coef.c[1 + h], hist[(long long)slots.s[h] * stride + i], d);
    const double sd = t * d;
    s_out[i] = sd;
    if (x) x[i] += sd;
    m = fmax(m, fabs(sd));
  }
  m = block_max(m, sh);
  // non-negative doubles order like their bit patterns: an integer max is exact and order-independent
  if (threadIdx.x == 0) atomicMax(dmax_bits, (unsigned long long)__double_as_longlong(m));
}

}  // namespace lbfgs
}  // namespace vs

using namespace vs;
using namespace vs::lbfgs;

static int pick_chunks(long long n, long long* per) {
  long long chunks = ceil_div(n, 4096);
  if (chunks > 2 * kNumSMs) chunks = 2 * kNumSMs;
  if (chunks < 1) chunks = 1;
  *per = ceil_div(n, chunks);
  return (int)ceil_div(n, *per);
}

extern "C" size_t vs_lbfgs_workspace(int64_t n, int m) {
  if (n <= 0 || m < 0 || m > VS_LBFGS_MAX_HIST) return 0;
  long long per;
  const int chunks = pick_chunks(n, &per);
  return (size_t)chunks * (kBase + 6 * (size_t)m) * sizeof(double) + 64;
}

extern "C" int vs_lbfgs_dots(int64_t n, const double* g, const double* g_prev, const double* s_new, double* y_out,
                             const double* hist, int64_t hist_stride, const int32_t* s_slots_host, const int32_t* y_slots_host,
                             int m, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(n > 0 && g && out, VS_ERR_INVALID, "vs_lbfgs_dots: null pointer or empty vector");
  VS_REQUIRE(m >= 0 && m <= VS_LBFGS_MAX_HIST, VS_ERR_UNSUPPORTED, "vs_lbfgs_dots: history %d outside 0..%d", m, VS_LBFGS_MAX_HIST);
  VS_REQUIRE(m == 0 || (hist && s_slots_host && y_slots_host && hist_stride >= n), VS_ERR_INVALID, "vs_lbfgs_dots: history arguments");
  VS_REQUIRE(workspace && workspace_bytes >= vs_lbfgs_workspace(n, m), VS_ERR_WORKSPACE, "vs_lbfgs_dots: workspace too small");
  VS_REQUIRE(((uintptr_t)workspace & 7) == 0, VS_ERR_INVALID, "vs_lbfgs_dots: workspace must be 8-byte aligned");
  Slots sl;
  for (int i = 0; i < m; ++i) { sl.s[i] = s_slots_host[i]; sl.s[m + i] = y_slots_host[i]; }
  const int nh = 2 * m, nout = kBase + 3 * nh;
  long long per;
  const int chunks = pick_chunks(n, &per);
  const int groups = nh > 0 ? (int)ceil_div(nh, kHG) : 1;
  double* part = reinterpret_cast<double*>(workspace);
  VS_LAUNCH(dots_kernel, dim3(chunks, groups), 256, 0, stream, (long long)n, per, g, g_prev, s_new, y_out, hist, (long long)hist_stride,
            sl, nh, nout, part);
  VS_LAUNCH(dots_reduce_kernel, (unsigned)ceil_div(nout, 128), 128, 0, stream, part, chunks, nout, out);
  return VS_OK;
}

extern "C" int vs_lbfgs_direction(int64_t n, const double* g, const double* hist, int64_t hist_stride, const int32_t* s_slots_host,
                                  const int32_t* y_slots_host, int m, const double* coef_host, double t, double* x, double* s_out,
                                  double* dmax_out, void* stream) {
  VS_REQUIRE(n > 0 && g && s_out && dmax_out && coef_host, VS_ERR_INVALID, "vs_lbfgs_direction: null pointer or empty vector");
  VS_REQUIRE(m >= 0 && m <= VS_LBFGS_MAX_HIST, VS_ERR_UNSUPPORTED, "vs_lbfgs_direction: history %d outside 0..%d", m, VS_LBFGS_MAX_HIST);
  VS_REQUIRE(m == 0 || (hist && s_slots_host && y_slots_host && hist_stride >= n), VS_ERR_INVALID, "vs_lbfgs_direction: history arguments");
  Slots sl;
  Coef cf;
  for (int i = 0; i < m; ++i) { sl.s[i] = s_slots_host[i]; sl.s[m + i] = y_slots_host[i]; }
  for (int i = 0; i < 2 * m + 1; ++i) cf.c[i] = coef_host[i];
  VS_CHECK_CUDA(cudaMemsetAsync(dmax_out, 0, sizeof(double), (cudaStream_t)stream));
  long long blocks = ceil_div(n, 256 * 4);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  VS_LAUNCH(direction_kernel, (unsigned)blocks, 256, 0, stream, (long long)n, g, hist, (long long)hist_stride, sl, 2 * m, cf, t, x, s_out,
            reinterpret_cast<unsigned long long*>(dmax_out));
  return VS_OK;
}
