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

constexpr int kBase = 8;      // base scalars: g.g, |g|_1, |g|_inf, y.y, y.s, s.g, y.g, (unused)

struct Slots { int32_t s[2 * VS_LBFGS_MAX_HIST]; };              // element offsets / stride of the 2m vectors
struct Coef { double c[2 * VS_LBFGS_MAX_HIST + 1]; };            // cg, then one coefficient per history vector

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

// One block per chunk of kChunk elements.  Each thread keeps its 8 elements of g, y = g - g_prev and s_new in
// registers (also writes y), then streams the 2m history vectors ONCE each with 128-bit loads, two vectors in
// flight at a time; per-warp partial sums go to shared memory without intermediate barriers and are combined
// in a fixed order at the end.  part[c * nout + o].
constexpr int kChunk = 2048;           // elements per block: 256 threads x 4 double2
constexpr int kPerThread = 4;          // double2 per thread

__device__ __forceinline__ double2 ld_stream_d2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

template <bool kVec>
__device__ __forceinline__ double2 load2(const double* base, long long i, long long n) {
  if constexpr (kVec) {
    if (i + 1 < n) return ld_stream_d2(base + i);
  }
  double2 r;
  r.x = i < n ? base[i] : 0.0;
  r.y = i + 1 < n ? base[i + 1] : 0.0;
  return r;
}

template <bool kVec>
__global__ void __launch_bounds__(256) dots_kernel(long long n, const double* __restrict__ g, const double* __restrict__ gp,
                                                   const double* __restrict__ s_new, double* __restrict__ y_out,
                                                   const double* __restrict__ hist, long long stride, const Slots slots, int nh,
                                                   int nout, double* __restrict__ part) {
  extern __shared__ double red[];          // [nout][8 warps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i0 = (long long)blockIdx.x * kChunk + 2 * threadIdx.x;
  double2 gv[kPerThread], yv[kPerThread], sv[kPerThread];
  double gg = 0.0, g1 = 0.0, gm = 0.0, yy = 0.0, ys = 0.0, sg = 0.0, yg = 0.0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const long long i = i0 + 512 * j;
    gv[j] = load2<kVec>(g, i, n);
    const double2 pv = gp ? load2<kVec>(gp, i, n) : make_double2(0.0, 0.0);
    sv[j] = s_new ? load2<kVec>(s_new, i, n) : make_double2(0.0, 0.0);
    yv[j] = make_double2(gv[j].x - pv.x, gv[j].y - pv.y);
    if (y_out) {
      if (kVec && i + 1 < n) *reinterpret_cast<double2*>(y_out + i) = yv[j];
      else { if (i < n) y_out[i] = yv[j].x; if (i + 1 < n) y_out[i + 1] = yv[j].y; }
    }
    gg = fma(gv[j].x, gv[j].x, gg); gg = fma(gv[j].y, gv[j].y, gg);
    g1 += fabs(gv[j].x) + fabs(gv[j].y);
    gm = fmax(gm, fmax(fabs(gv[j].x), fabs(gv[j].y)));
    yy = fma(yv[j].x, yv[j].x, yy); yy = fma(yv[j].y, yv[j].y, yy);
    ys = fma(yv[j].x, sv[j].x, ys); ys = fma(yv[j].y, sv[j].y, ys);
    sg = fma(sv[j].x, gv[j].x, sg); sg = fma(sv[j].y, gv[j].y, sg);
    yg = fma(yv[j].x, gv[j].x, yg); yg = fma(yv[j].y, gv[j].y, yg);
  }
  {
    const double b0 = warp_sum(gg), b1 = warp_sum(g1), b3 = warp_sum(yy), b4 = warp_sum(ys), b5 = warp_sum(sg), b6 = warp_sum(yg);
    double b2 = gm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b2 = fmax(b2, __shfl_xor_sync(0xffffffffu, b2, o));
    if (lane == 0) {
      red[0 * 8 + warp] = b0; red[1 * 8 + warp] = b1; red[2 * 8 + warp] = b2; red[3 * 8 + warp] = b3;
      red[4 * 8 + warp] = b4; red[5 * 8 + warp] = b5; red[6 * 8 + warp] = b6; red[7 * 8 + warp] = 0.0;
    }
  }
  // history vectors, two per iteration so 8 independent 128-bit loads are in flight per thread
  for (int h = 0; h < nh; h += 2) {
    const double* h0 = hist + (long long)slots.s[h] * stride;
    const bool two = h + 1 < nh;
    const double* h1 = two ? hist + (long long)slots.s[h + 1] * stride : h0;
    double2 a[kPerThread], b[kPerThread];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) a[j] = load2<kVec>(h0, i0 + 512 * j, n);
    if (two) {
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) b[j] = load2<kVec>(h1, i0 + 512 * j, n);
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      a0 = fma(a[j].x, gv[j].x, a0); a0 = fma(a[j].y, gv[j].y, a0);
      a1 = fma(a[j].x, yv[j].x, a1); a1 = fma(a[j].y, yv[j].y, a1);
      a2 = fma(a[j].x, sv[j].x, a2); a2 = fma(a[j].y, sv[j].y, a2);
    }
    if (two) {
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        c0 = fma(b[j].x, gv[j].x, c0); c0 = fma(b[j].y, gv[j].y, c0);
        c1 = fma(b[j].x, yv[j].x, c1); c1 = fma(b[j].y, yv[j].y, c1);
        c2 = fma(b[j].x, sv[j].x, c2); c2 = fma(b[j].y, sv[j].y, c2);
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { red[(kBase + 3 * h + 0) * 8 + warp] = a0; red[(kBase + 3 * h + 1) * 8 + warp] = a1; red[(kBase + 3 * h + 2) * 8 + warp] = a2; }
    if (two) {
      c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
      if (lane == 0) { red[(kBase + 3 * h + 3) * 8 + warp] = c0; red[(kBase + 3 * h + 4) * 8 + warp] = c1; red[(kBase + 3 * h + 5) * 8 + warp] = c2; }
    }
  }
  __syncthreads();
  double* out = part + (long long)blockIdx.x * nout;
  for (int o = threadIdx.x; o < nout; o += 256) {
    double s = 0.0;
    if (o == 2) {
      for (int w = 0; w < 8; ++w) s = fmax(s, red[o * 8 + w]);
    } else {
      for (int w = 0; w < 8; ++w) s += red[o * 8 + w];
    }
    out[o] = s;
  }
}

// out[o] = sum (max for o == 2) over the chunk partials: one block per output, thread t takes chunks t, t+128, ...
// and the 128 lane sums are combined by a fixed tree, so the result does not depend on scheduling
__global__ void __launch_bounds__(128) dots_reduce_kernel(const double* __restrict__ part, int chunks, int nout, double* __restrict__ out) {
  __shared__ double sh[128];
  const int o = blockIdx.x;
  const bool is_max = (o == 2);
  double s = 0.0;
  for (int c = threadIdx.x; c < chunks; c += 128) {
    const double v = part[(long long)c * nout + o];
    s = is_max ? fmax(s, v) : s + v;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int h = 64; h > 0; h >>= 1) {
    if (threadIdx.x < h) sh[threadIdx.x] = is_max ? fmax(sh[threadIdx.x], sh[threadIdx.x + h]) : sh[threadIdx.x] + sh[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[o] = sh[0];
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

static int num_chunks(long long n) { return (int)ceil_div(n, kChunk); }

extern "C" size_t vs_lbfgs_workspace(int64_t n, int m) {
  if (n <= 0 || m < 0 || m > VS_LBFGS_MAX_HIST) return 0;
  return (size_t)num_chunks(n) * (kBase + 6 * (size_t)m) * sizeof(double) + 64;
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
  VS_REQUIRE(n < (1ll << 40), VS_ERR_UNSUPPORTED, "vs_lbfgs_dots: vector too long");
  const int chunks = num_chunks(n);
  double* part = reinterpret_cast<double*>(workspace);
  const size_t smem = (size_t)nout * 8 * sizeof(double);
  // 128-bit loads need every vector base 16-byte aligned (hist slots: even stride)
  const bool vec = ((((uintptr_t)g | (uintptr_t)g_prev | (uintptr_t)s_new | (uintptr_t)y_out | (uintptr_t)hist) & 15) == 0) && (hist_stride % 2 == 0 || m == 0);
  if (vec) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(dots_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(dots_kernel<true>, chunks, 256, smem, stream, (long long)n, g, g_prev, s_new, y_out, hist, (long long)hist_stride, sl, nh, nout, part);
  } else {
    VS_CHECK_CUDA(cudaFuncSetAttribute(dots_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(dots_kernel<false>, chunks, 256, smem, stream, (long long)n, g, g_prev, s_new, y_out, hist, (long long)hist_stride, sl, nh, nout, part);
  }
  VS_LAUNCH(dots_reduce_kernel, (unsigned)nout, 128, 0, stream, part, chunks, nout, out);
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
