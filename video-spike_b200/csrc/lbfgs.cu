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
// Both are HBM-bound (8 or 4 B/element/history vector: the history may be stored in float32).  Reductions are two-stage with a fixed order: bit-reproducible.
#include <string.h>
#include "common.cuh"

namespace vs {
namespace lbfgs {

constexpr int kBase = 8;      // base scalars: g.g, |g|_1, |g|_inf, y.y, y.s, s.g, y.g, (unused)

struct Slots { int32_t s[2 * VS_LBFGS_MAX_HIST]; };              // element offsets / stride of the 2m vectors

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

// One block per chunk of kChunk elements.  Each thread keeps its 8 elements (two groups of 4 consecutive) of g,
// y = g - g_prev and s_new in registers (also writes y), then streams the 2m history vectors ONCE each with 128-bit
// loads, two vectors in flight at a time; per-warp partial sums go to shared memory without intermediate barriers and
// are combined in a fixed order at the end.  part[c * nout + o].
// HT = storage type of the history vectors (s_i, y_i, and therefore s_new / y_out): double, or float to halve the
// traffic of both passes.  All arithmetic is fp64; with float storage every inner product is taken with the STORED
// (rounded) vectors so the Gram matrix stays consistent with what later iterations read back.
constexpr int kChunk = 2048;           // elements per block: 256 threads x 2 groups x 4
constexpr int kGroups = 2;

struct D4 { double v[4]; };

template <bool kVec>
__device__ __forceinline__ D4 load4(const double* base, long long i, long long n) {
  D4 r;
  if (kVec && i + 3 < n) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(base + i));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[2]), "=d"(r.v[3]) : "l"(base + i + 2));
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) r.v[q] = i + q < n ? base[i + q] : 0.0;
  }
  return r;
}
template <bool kVec>
__device__ __forceinline__ D4 load4(const float* base, long long i, long long n) {
  D4 r;
  if (kVec && i + 3 < n) {
    float4 f;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "l"(base + i));
    r.v[0] = f.x; r.v[1] = f.y; r.v[2] = f.z; r.v[3] = f.w;
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) r.v[q] = i + q < n ? (double)base[i + q] : 0.0;
  }
  return r;
}
template <bool kVec>
__device__ __forceinline__ void store4(double* base, long long i, long long n, const D4& v) {
  if (kVec && i + 3 < n) {
    *reinterpret_cast<double2*>(base + i) = make_double2(v.v[0], v.v[1]);
    *reinterpret_cast<double2*>(base + i + 2) = make_double2(v.v[2], v.v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i + q < n) base[i + q] = v.v[q];
  }
}
template <bool kVec>
__device__ __forceinline__ void store4(float* base, long long i, long long n, const D4& v) {
  if (kVec && i + 3 < n) {
    *reinterpret_cast<float4*>(base + i) = make_float4((float)v.v[0], (float)v.v[1], (float)v.v[2], (float)v.v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i + q < n) base[i + q] = (float)v.v[q];
  }
}
template <typename T> struct H4 { T v[4]; };
template <bool kVec>
__device__ __forceinline__ H4<double> load4h(const double* base, long long i, long long n) {
  const D4 d = load4<kVec>(base, i, n);
  H4<double> r;
#pragma unroll
  for (int q = 0; q < 4; ++q) r.v[q] = d.v[q];
  return r;
}
template <bool kVec>
__device__ __forceinline__ H4<float> load4h(const float* base, long long i, long long n) {
  H4<float> r;
  if (kVec && i + 3 < n) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(base + i));
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) r.v[q] = i + q < n ? base[i + q] : 0.f;
  }
  return r;
}
__device__ __forceinline__ double round_to(double v, const double*) { return v; }
__device__ __forceinline__ double round_to(double v, const float*) { return (double)(float)v; }

template <typename HT, bool kVec>
__global__ void __launch_bounds__(256) dots_kernel(long long n, const double* __restrict__ g, const double* __restrict__ gp,
                                                   const HT* __restrict__ s_new, HT* __restrict__ y_out,
                                                   const HT* __restrict__ hist, long long stride, const Slots slots, int nh,
                                                   int nout, double* __restrict__ part, const vs_lbfgs_dev* __restrict__ dev,
                                                   const vs_lbfgs_cdev* __restrict__ cdev) {
  extern __shared__ double red[];          // [nout][8 warps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int m_dev = 0;
  if (cdev) {  // compact history: the basis {g_0, y_0, y_1, ...} lives in slots 0 .. nb-1; this pass appends y_new (g_0 on the first one)
    if (cdev->done) return;
    nh = cdev->nb;
    if (!cdev->have_prev) gp = nullptr;
    s_new = nullptr;
    y_out = const_cast<HT*>(hist) + (long long)cdev->nb * stride;
  }
  if (dev) {   // device-driven optimiser: everything that depends on its decisions comes from the state
    if (dev->done) return;
    m_dev = dev->m;
    nh = 2 * m_dev;
    if (!dev->have_prev) gp = nullptr;
    s_new = dev->have_s ? hist + (long long)dev->s_cur * stride : nullptr;
    y_out = (dev->have_prev && dev->have_s) ? const_cast<HT*>(hist) + (long long)dev->y_next * stride : nullptr;
  }
  auto slot_of = [&](int h) -> int { return cdev ? h : (dev ? (h < m_dev ? dev->s_slots[h] : dev->y_slots[h - m_dev]) : slots.s[h]); };
  const int nvalid = kBase + 3 * nh;       // outputs this launch produces (nout is the stride of the partial rows)
  const long long i0 = (long long)blockIdx.x * kChunk + 4 * threadIdx.x;
  D4 gv[kGroups], yv[kGroups], sv[kGroups];
  double gg = 0.0, g1 = 0.0, gm = 0.0, yy = 0.0, ys = 0.0, sg = 0.0, yg = 0.0;
#pragma unroll
  for (int j = 0; j < kGroups; ++j) {
    const long long i = i0 + 1024 * j;
    gv[j] = load4<kVec>(g, i, n);
    D4 pv;
    if (gp) pv = load4<kVec>(gp, i, n);
    if (s_new) sv[j] = load4<kVec>(s_new, i, n);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!gp) pv.v[q] = 0.0;
      if (!s_new) sv[j].v[q] = 0.0;
      yv[j].v[q] = round_to(gv[j].v[q] - pv.v[q], hist);
    }
    if (y_out) store4<kVec>(y_out, i, n, yv[j]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double a = gv[j].v[q], y = yv[j].v[q], sn = sv[j].v[q];
      gg = fma(a, a, gg); g1 += fabs(a); gm = fmax(gm, fabs(a));
      yy = fma(y, y, yy); ys = fma(y, sn, ys); sg = fma(sn, a, sg); yg = fma(y, a, yg);
    }
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
  // history vectors, two per iteration so several independent 128-bit loads are in flight per thread.  The vector
  // fp64 pipe of the B200 is narrow (measured ~6 ops/clk/SM: a float64 history is bandwidth-bound at 3 DFMA per
  // element, a float32 one would be conversion+DFMA-bound), so with float storage the 8 products of a thread are
  // accumulated in fp32 -- the same order of rounding as the storage itself -- and only the per-thread partials are
  // widened; every cross-thread sum stays fp64.
  using AT = HT;                                  // per-thread accumulation type
  AT gf[kGroups][4], yf[kGroups][4], sf[kGroups][4];
#pragma unroll
  for (int j = 0; j < kGroups; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) { gf[j][q] = (AT)gv[j].v[q]; yf[j][q] = (AT)yv[j].v[q]; sf[j][q] = (AT)sv[j].v[q]; }
  for (int h = 0; h < nh; h += 2) {
    const HT* h0 = hist + (long long)slot_of(h) * stride;
    const bool two = h + 1 < nh;
    const HT* h1 = two ? hist + (long long)slot_of(h + 1) * stride : h0;
    H4<HT> a[kGroups], b[kGroups];
#pragma unroll
    for (int j = 0; j < kGroups; ++j) a[j] = load4h<kVec>(h0, i0 + 1024 * j, n);
    if (two) {
#pragma unroll
      for (int j = 0; j < kGroups; ++j) b[j] = load4h<kVec>(h1, i0 + 1024 * j, n);
    }
    AT a0 = 0, a1 = 0, a2 = 0, c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int j = 0; j < kGroups; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a0 = fma(a[j].v[q], gf[j][q], a0); a1 = fma(a[j].v[q], yf[j][q], a1); a2 = fma(a[j].v[q], sf[j][q], a2);
      }
    if (two) {
#pragma unroll
      for (int j = 0; j < kGroups; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          c0 = fma(b[j].v[q], gf[j][q], c0); c1 = fma(b[j].v[q], yf[j][q], c1); c2 = fma(b[j].v[q], sf[j][q], c2);
        }
    }
    const double d0 = warp_sum((double)a0), d1 = warp_sum((double)a1), d2 = warp_sum((double)a2);
    if (lane == 0) { red[(kBase + 3 * h + 0) * 8 + warp] = d0; red[(kBase + 3 * h + 1) * 8 + warp] = d1; red[(kBase + 3 * h + 2) * 8 + warp] = d2; }
    if (two) {
      const double e0 = warp_sum((double)c0), e1 = warp_sum((double)c1), e2 = warp_sum((double)c2);
      if (lane == 0) { red[(kBase + 3 * h + 3) * 8 + warp] = e0; red[(kBase + 3 * h + 4) * 8 + warp] = e1; red[(kBase + 3 * h + 5) * 8 + warp] = e2; }
    }
  }
  __syncthreads();
  double* out = part + (long long)blockIdx.x * nout;
  for (int o = threadIdx.x; o < nvalid; o += 256) {
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
__global__ void __launch_bounds__(128) dots_reduce_kernel(const double* __restrict__ part, int chunks, int nout, double* __restrict__ out,
                                                          vs_lbfgs_dev* __restrict__ dev, vs_lbfgs_cdev* __restrict__ cdev) {
  __shared__ double sh[128];
  const int o = blockIdx.x;
  if (dev) {
    if (dev->done || o >= kBase + 6 * dev->m) return;
    out = dev->out;
  }
  if (cdev) {
    if (cdev->done || o >= kBase + 3 * cdev->nb) return;
    out = cdev->out;
  }
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

// CT = arithmetic of the linear combination: double with a float64 history; float with a float32 one (the combination
// is then rounded like its operands; the parameter update x += t*d itself is always fp64)
template <typename HT>
struct CoefT { HT c[2 * VS_LBFGS_MAX_HIST + 1]; };

template <typename HT, bool kVec>
__global__ void __launch_bounds__(256) direction_kernel(long long n, const double* __restrict__ g, const HT* __restrict__ hist,
                                                        long long stride, const Slots slots, int nh, const CoefT<HT> coef, double cg,
                                                        double t, double* __restrict__ x, HT* __restrict__ s_out,
                                                        unsigned long long* __restrict__ dmax_bits, vs_lbfgs_dev* __restrict__ dev,
                                                        vs_lbfgs_cdev* __restrict__ cdev) {
  __shared__ double sh[8];
  __shared__ HT s_coef[2 * VS_LBFGS_MAX_HIST];
  __shared__ const HT* s_ptr[2 * VS_LBFGS_MAX_HIST];
  if (cdev) {  // compact history: d = cg*g + sum_l coef[1+l] * basis_l; the step s = t*d is kept as coefficients only (no store)
    if (cdev->done) return;
    nh = cdev->nb;
    for (int h = threadIdx.x; h < nh; h += 256) {
      s_coef[h] = (HT)cdev->coef[1 + h];
      s_ptr[h] = hist + (long long)h * stride;
    }
    cg = cdev->coef[0];
    t = cdev->t;
    s_out = nullptr;
    dmax_bits = reinterpret_cast<unsigned long long*>(&cdev->dmax);
  } else if (dev) {   // device-driven optimiser: coefficients, slots, step length and destination come from the state
    if (dev->done) return;
    const int md = dev->m;
    nh = 2 * md;
    for (int h = threadIdx.x; h < nh; h += 256) {
      s_coef[h] = (HT)dev->coef[1 + h];
      s_ptr[h] = hist + (long long)(h < md ? dev->s_slots[h] : dev->y_slots[h - md]) * stride;
    }
    cg = dev->coef[0];
    t = dev->t;
    s_out = const_cast<HT*>(hist) + (long long)dev->s_cur * stride;
    dmax_bits = reinterpret_cast<unsigned long long*>(&dev->dmax);
  } else {
    for (int h = threadIdx.x; h < nh; h += 256) { s_coef[h] = coef.c[1 + h]; s_ptr[h] = hist + (long long)slots.s[h] * stride; }
  }
  __syncthreads();
  double m = 0.0;
  // 4 consecutive elements per thread and iteration: one coefficient / pointer fetch per 128-bit history load
  const long long step = (long long)gridDim.x * 256 * 4;
  for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += step) {
    HT acc[4] = {0, 0, 0, 0};
#pragma unroll 4
    for (int h = 0; h < nh; ++h) {
      const H4<HT> v = load4h<kVec>(s_ptr[h], i, n);
      const HT c = s_coef[h];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(c, v.v[q], acc[q]);
    }
    const D4 gv = load4<kVec>(g, i, n);
    D4 sd;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sd.v[q] = t * fma(cg, gv.v[q], (double)acc[q]);
      if (i + q < n) m = fmax(m, fabs(sd.v[q]));
    }
    if (s_out) store4<kVec>(s_out, i, n, sd);
    if (x) {
      if (kVec && i + 3 < n) {
        double2* xp = reinterpret_cast<double2*>(x + i);
        double2 a = xp[0], b2 = xp[1];
        a.x += sd.v[0]; a.y += sd.v[1]; b2.x += sd.v[2]; b2.y += sd.v[3];
        xp[0] = a; xp[1] = b2;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i + q < n) x[i + q] += sd.v[q];
      }
    }
  }
  m = block_max(m, sh);
  // non-negative doubles order like their bit patterns: an integer max is exact and order-independent
  if (threadIdx.x == 0) atomicMax(dmax_bits, (unsigned long long)__double_as_longlong(m));
}

// ------------------------------------------------------------------ device-driven optimiser: the decisions
// Mirrors, statement for statement, the part of torch.optim.LBFGS.step between two closure evaluations (and
// optim.py::FusedLBFGS.step, the host-driven version of the same logic): post-evaluation termination tests, memory update
// with the ys > 1e-10 guard and the history_size window, two-loop recursion in coefficient space, step length,
// directional-derivative test, slot bookkeeping for the next passes.
__device__ __forceinline__ int pop_slot(vs_lbfgs_dev* d) { return d->free_slots[--d->n_free]; }
__device__ __forceinline__ void push_slot(vs_lbfgs_dev* d, int s) { d->free_slots[d->n_free++] = s; }

constexpr int kUpdSmemM = 32;   // memories up to this size run the recursion out of shared memory

// thread 0: everything up to and including the memory update; returns false when the iteration must not proceed
__device__ bool update_decide(vs_lbfgs_dev* __restrict__ d, const double* __restrict__ loss_ptr, double tol_g, double tol_c,
                              int max_eval, int hsize, int first_eval, double* sg, double* yg) {
  constexpr int H = VS_LBFGS_MAX_HIST;
  if (first_eval) { d->n_iter = 0; d->cur_evals = 0; d->done = 0; }
  if (d->done) return false;
  const double loss = *loss_ptr;
  const double* out = d->out;
  d->loss = loss;
  d->cur_evals += 1;
  d->func_evals += 1;
  if (first_eval) {
    if (out[2] <= tol_g) { d->done = 1; return false; }        // optimal condition at the first evaluation
  } else {
    if (d->cur_evals >= max_eval) d->done = 3;
    else if (out[2] <= tol_g) d->done = 4;
    else if (d->dmax <= tol_c) d->done = 5;
    else if (fabs(loss - d->prev_loss) < tol_c) d->done = 6;
    if (d->done) return false;
  }
  d->n_iter += 1;
  d->total_iter += 1;
  int m = d->m;
  for (int i = 0; i < m; ++i) { sg[i] = out[kBase + 3 * i]; yg[i] = out[kBase + 3 * (m + i)]; }
  double Hd = d->H_diag;
  if (d->total_iter == 1) {
    Hd = 1.0;
  } else if (d->have_prev && d->have_s) {
    const double yy = out[3], ys = out[4];
    if (ys > 1e-10) {
      int lo = 0;
      if (m == hsize) {                                          // limited memory: drop the oldest pair
        push_slot(d, d->s_slots[0]); push_slot(d, d->y_slots[0]);
        for (int i = 1; i < m; ++i) {
          d->s_slots[i - 1] = d->s_slots[i]; d->y_slots[i - 1] = d->y_slots[i];
          for (int j = 1; j < m; ++j) { d->SY[(i - 1) * H + (j - 1)] = d->SY[i * H + j]; d->YY[(i - 1) * H + (j - 1)] = d->YY[i * H + j]; }
          sg[i - 1] = sg[i]; yg[i - 1] = yg[i];
        }
        lo = 1; m -= 1;
      }
      const int mo = m + lo;                                     // pairs the dots pass saw
      for (int i = 0; i < m; ++i) {
        const double sy_col = out[kBase + 3 * (i + lo) + 1];           // s_i . y_new
        const double yy_col = out[kBase + 3 * (mo + i + lo) + 1];      // y_i . y_new
        const double ys_row = out[kBase + 3 * (mo + i + lo) + 2];      // y_i . s_new
        d->SY[i * H + m] = sy_col; d->SY[m * H + i] = ys_row;
        d->YY[i * H + m] = yy_col; d->YY[m * H + i] = yy_col;
      }
      d->SY[m * H + m] = ys; d->YY[m * H + m] = yy;
      d->s_slots[m] = d->s_cur; d->y_slots[m] = d->y_next;
      sg[m] = out[5]; yg[m] = out[6];
      m += 1;
      Hd = ys / yy;
    } else {
      push_slot(d, d->s_cur); push_slot(d, d->y_next);
    }
    d->have_s = 0;
  }
  d->m = m;
  d->H_diag = Hd;
  return true;
}

// two-loop recursion in coefficient space (same statement order as csrc/host_lbfgs.cpp); SY / YY with row pitch ld
__device__ void update_direction(vs_lbfgs_dev* __restrict__ d, const double* SY, const double* YY, int ld, const double* sg,
                                 const double* yg, double loss, double lr, double tol_c) {
  constexpr int H = VS_LBFGS_MAX_HIST;
  const int m = d->m;
  const double Hd = d->H_diag, gg = d->out[0], g1 = d->out[1];
  double al[H], ro[H];
  double* cs = d->coef + 1;
  double* cy = d->coef + 1 + m;
  for (int i = 0; i < m; ++i) ro[i] = 1.0 / SY[i * ld + i];
  for (int i = m - 1; i >= 0; --i) {
    double sq = -sg[i];
    for (int j = i + 1; j < m; ++j) sq -= al[j] * SY[i * ld + j];
    al[i] = sq * ro[i];
  }
  for (int i = 0; i < m; ++i) {
    double yr = -yg[i];
    for (int j = 0; j < m; ++j) yr -= al[j] * YY[i * ld + j];
    yr *= Hd;
    for (int j = 0; j < i; ++j) yr += cs[j] * SY[j * ld + i];
    cs[i] = al[i] - yr * ro[i];
  }
  d->coef[0] = -Hd;
  double dot_s = 0.0, dot_y = 0.0;
  for (int i = 0; i < m; ++i) cy[i] = -Hd * al[i];
  for (int i = 0; i < m; ++i) dot_s += cs[i] * sg[i];
  for (int i = 0; i < m; ++i) dot_y += cy[i] * yg[i];
  d->gtd = m ? -Hd * gg + dot_s + dot_y : -Hd * gg;
  d->prev_loss = loss;
  d->have_prev = 1;
  d->t = d->total_iter == 1 ? fmin(1.0, 1.0 / g1) * lr : lr;
  if (d->gtd > -tol_c) { d->done = 2; return; }                 // directional derivative below tolerance: no move
  if (d->have_s) push_slot(d, d->s_cur);                        // an unconsumed step (no previous gradient yet)
  d->s_cur = pop_slot(d);                                       // the direction pass writes s = t*d here
  d->have_s = 1;
  d->y_next = pop_slot(d);                                      // the next dots pass writes y here
  d->dmax = 0.0;
}

// One block.  Thread 0 takes the decisions; the other threads only stage the m x m Gram blocks in shared memory so that
// the O(m^2) recursion of thread 0 does not pay a global-memory round trip per element.
__global__ void __launch_bounds__(128) update_kernel(vs_lbfgs_dev* __restrict__ d, const double* __restrict__ loss_ptr, double lr,
                                                     double tol_g, double tol_c, int max_eval, int hsize, int first_eval) {
  constexpr int H = VS_LBFGS_MAX_HIST;
  __shared__ double s_SY[kUpdSmemM * kUpdSmemM], s_YY[kUpdSmemM * kUpdSmemM];
  __shared__ double s_sg[H], s_yg[H];
  __shared__ int s_go, s_m;
  if (threadIdx.x == 0) {
    s_go = update_decide(d, loss_ptr, tol_g, tol_c, max_eval, hsize, first_eval, s_sg, s_yg) ? 1 : 0;
    s_m = d->m;
    __threadfence_block();
  }
  __syncthreads();
  if (!s_go) return;
  const int m = s_m;
  const bool in_smem = m <= kUpdSmemM;
  if (in_smem)
    for (int e = threadIdx.x; e < m * m; e += 128) {
      const int i = e / m, j = e % m;
      s_SY[i * kUpdSmemM + j] = d->SY[i * H + j];
      s_YY[i * kUpdSmemM + j] = d->YY[i * H + j];
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (in_smem) update_direction(d, s_SY, s_YY, kUpdSmemM, s_sg, s_yg, d->loss, lr, tol_c);
    else update_direction(d, d->SY, d->YY, H, s_sg, s_yg, d->loss, lr, tol_c);
  }
}

// ------------------------------------------------------------------ compact history (device-driven, float64)
// Every vector torch.optim.LBFGS keeps -- the steps s_i = t_i d_i, the gradient differences y_i, the direction d -- lies
// in the span of the evaluated gradients.  Keep ONE vector per closure evaluation, the basis b_0 = g_0, b_l = g_l - g_(l-1)
// (the y vectors themselves, formed exactly as before), its Gram matrix P, and the steps as COEFFICIENT rows over the
// basis: each pass then streams nb vectors instead of the 2m of the (s_i, y_i) scheme, and no step vector is written.
//   dots pass      appends b_nb = g - g_prev, out[8+3l+{0,1}] = b_l.g, b_l.b_nb; out[3] = b_nb.b_nb, out[6] = b_nb.g
//   update         P grows by one row/column; s.y, s.g, y.y, y.g of the two-loop recursion are contractions of P and the
//                  coefficient rows; same decisions in the same order as update_kernel (= torch.optim.LBFGS.step)
//   direction pass d = cg*g + sum_l coef[1+l] b_l;  x += t*d;  the new step's coefficients are t*(cg + coef[1+l])
//                  because g = sum_l b_l
constexpr int CB = VS_LBFGS_CMAX;

__global__ void __launch_bounds__(128) update_compact_kernel(vs_lbfgs_cdev* __restrict__ d, const double* __restrict__ loss_ptr, double lr,
                                                             double tol_g, double tol_c, int max_eval, int first_eval) {
  __shared__ double sP[32 * 32], sA[32 * 32], sSY[32 * 32], sYY[32 * 32];
  __shared__ double pg[CB], sg[CB], yg[CB], al[CB], ro[CB], cs[CB], cy[CB];
  __shared__ int s_go, s_nb, s_m, s_small;
  if (threadIdx.x == 0) {
    s_go = 0;
    if (first_eval) { d->n_iter = 0; d->cur_evals = 0; d->done = 0; }
    if (!d->done) {
      const double* out = d->out;
      const int nbn = d->nb;                                    // index of the vector the dots pass just appended
      if (nbn >= CB) {
        d->done = 7;                                            // basis full: the caller must use the (s, y) scheme
      } else {
        for (int l = 0; l < nbn; ++l) {
          const double v = out[kBase + 3 * l + 1];
          d->P[l * CB + nbn] = v; d->P[nbn * CB + l] = v;
        }
        d->P[nbn * CB + nbn] = out[3];
        d->nb = nbn + 1;
        s_go = 1;
      }
    }
    s_nb = d->nb; s_m = d->m;
    s_small = (d->nb <= 32) ? 1 : 0;
    __threadfence_block();
  }
  __syncthreads();
  if (!s_go) return;
  const int nb = s_nb;
  const bool small = s_small != 0;
  if (small) {   // stage the Gram matrix and the coefficient rows: the O(m^2 nb) algebra below is a dependent chain on one thread
    for (int e = threadIdx.x; e < nb * nb; e += 128) sP[(e / nb) * 32 + (e % nb)] = d->P[(e / nb) * CB + (e % nb)];
    for (int e = threadIdx.x; e < s_m * nb; e += 128) sA[(e / nb) * 32 + (e % nb)] = d->A[(e / nb) * CB + (e % nb)];
    for (int e = threadIdx.x; e < s_m * s_m; e += 128) {
      sSY[(e / s_m) * 32 + (e % s_m)] = d->SY[(e / s_m) * CB + (e % s_m)];
      sYY[(e / s_m) * 32 + (e % s_m)] = d->YY[(e / s_m) * CB + (e % s_m)];
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double* P = small ? sP : d->P;
  double* A = small ? sA : d->A;
  double* SY = small ? sSY : d->SY;
  double* YY = small ? sYY : d->YY;
  const int ld = small ? 32 : CB;
  const double* out = d->out;
  const double loss = *loss_ptr;
  d->loss = loss;
  d->cur_evals += 1;
  d->func_evals += 1;
  if (first_eval) {
    if (out[2] <= tol_g) { d->done = 1; return; }
  } else {
    if (d->cur_evals >= max_eval) d->done = 3;
    else if (out[2] <= tol_g) d->done = 4;
    else if (d->dmax <= tol_c) d->done = 5;
    else if (fabs(loss - d->prev_loss) < tol_c) d->done = 6;
    if (d->done) return;
  }
  d->n_iter += 1;
  d->total_iter += 1;
  const int nn = nb - 1;                                        // the newest basis vector = y_new
  for (int l = 0; l < nn; ++l) pg[l] = out[kBase + 3 * l];
  pg[nn] = out[6];
  int m = d->m;
  double Hd = d->H_diag;
  if (d->total_iter == 1) {
    Hd = 1.0;
  } else if (d->have_prev && d->have_s) {
    // the step s = t*d taken before this evaluation, as coefficients over the basis it was built from (indices < nn)
    double ys = 0.0;
    for (int l = 0; l < nn; ++l) ys += d->Acur[l] * P[l * ld + nn];
    const double yy = P[nn * ld + nn];
    if (ys > 1e-10 && m < CB) {
      for (int i = 0; i < m; ++i) {
        double a = 0.0, b2 = 0.0;
        for (int l = 0; l < nn; ++l) { a += A[i * ld + l] * P[l * ld + nn]; b2 += d->Acur[l] * P[l * ld + d->iy[i]]; }
        const double c = P[d->iy[i] * ld + nn];
        d->SY[i * CB + m] = a;                                    // s_i . y_new
        d->SY[m * CB + i] = b2;                                   // s_new . y_i
        d->YY[i * CB + m] = c; d->YY[m * CB + i] = c;
        if (small) { SY[i * ld + m] = a; SY[m * ld + i] = b2; YY[i * ld + m] = c; YY[m * ld + i] = c; }
      }
      d->SY[m * CB + m] = ys; d->YY[m * CB + m] = yy;
      if (small) { SY[m * ld + m] = ys; YY[m * ld + m] = yy; }
      for (int l = 0; l < CB; ++l) { const double v = l < nn ? d->Acur[l] : 0.0; d->A[m * CB + l] = v; if (small && l < 32) A[m * ld + l] = v; }
      d->iy[m] = nn;
      m += 1;
      Hd = ys / yy;
    }
    d->have_s = 0;
  }
  d->m = m;
  d->H_diag = Hd;
  for (int i = 0; i < m; ++i) {
    double a = 0.0;
    for (int l = 0; l < nb; ++l) a += A[i * ld + l] * pg[l];
    sg[i] = a;
    yg[i] = pg[d->iy[i]];
  }
  // two-loop recursion in coefficient space (statement order of update_direction / csrc/host_lbfgs.cpp)
  const double gg = out[0], g1 = out[1];
  for (int i = 0; i < m; ++i) ro[i] = 1.0 / SY[i * ld + i];
  for (int i = m - 1; i >= 0; --i) {
    double sq = -sg[i];
    for (int j = i + 1; j < m; ++j) sq -= al[j] * SY[i * ld + j];
    al[i] = sq * ro[i];
  }
  for (int i = 0; i < m; ++i) {
    double yr = -yg[i];
    for (int j = 0; j < m; ++j) yr -= al[j] * YY[i * ld + j];
    yr *= Hd;
    for (int j = 0; j < i; ++j) yr += cs[j] * SY[j * ld + i];
    cs[i] = al[i] - yr * ro[i];
  }
  double dot_s = 0.0, dot_y = 0.0;
  for (int i = 0; i < m; ++i) cy[i] = -Hd * al[i];
  for (int i = 0; i < m; ++i) dot_s += cs[i] * sg[i];
  for (int i = 0; i < m; ++i) dot_y += cy[i] * yg[i];
  d->gtd = m ? -Hd * gg + dot_s + dot_y : -Hd * gg;
  d->prev_loss = loss;
  d->have_prev = 1;
  d->t = d->total_iter == 1 ? fmin(1.0, 1.0 / g1) * lr : lr;
  if (d->gtd > -tol_c) { d->done = 2; return; }
  // d over the basis (the cg*g term is taken from the gradient buffer by the direction pass)
  const double cg = -Hd;
  d->coef[0] = cg;
  for (int l = 0; l < nb; ++l) {
    double c = 0.0;
    for (int i = 0; i < m; ++i) c += cs[i] * A[i * ld + l];
    d->coef[1 + l] = c;
  }
  for (int i = 0; i < m; ++i) d->coef[1 + d->iy[i]] += cy[i];
  for (int l = 0; l < nb; ++l) d->Acur[l] = d->t * (cg + d->coef[1 + l]);        // g = sum_l b_l
  d->have_s = 1;
  d->dmax = 0.0;
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

template <typename HT>
static int launch_dots(long long n, const double* g, const double* g_prev, const void* s_new, void* y_out, const void* hist,
                       long long hist_stride, const Slots& sl, int nh, int nout, double* part, int chunks, void* stream,
                       const vs_lbfgs_dev* dev = nullptr, const vs_lbfgs_cdev* cdev = nullptr) {
  const size_t smem = (size_t)nout * 8 * sizeof(double);
  // 128-bit loads need every vector base 16-byte aligned (hist slots: pitch a multiple of 4 elements)
  const bool vec = ((((uintptr_t)g | (uintptr_t)g_prev | (uintptr_t)s_new | (uintptr_t)y_out | (uintptr_t)hist) & 15) == 0) &&
                   (hist_stride % 4 == 0 || (nh == 0 && !dev && !cdev));
  if (vec) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(dots_kernel<HT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH((dots_kernel<HT, true>), chunks, 256, smem, stream, n, g, g_prev, (const HT*)s_new, (HT*)y_out, (const HT*)hist, hist_stride, sl, nh, nout, part, dev, cdev);
  } else {
    VS_CHECK_CUDA(cudaFuncSetAttribute(dots_kernel<HT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH((dots_kernel<HT, false>), chunks, 256, smem, stream, n, g, g_prev, (const HT*)s_new, (HT*)y_out, (const HT*)hist, hist_stride, sl, nh, nout, part, dev, cdev);
  }
  return VS_OK;
}

extern "C" int vs_lbfgs_dots(int64_t n, const double* g, const double* g_prev, const void* s_new, void* y_out,
                             const void* hist, int64_t hist_stride, int hist_f32, const int32_t* s_slots_host,
                             const int32_t* y_slots_host, int m, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(n > 0 && g && out, VS_ERR_INVALID, "vs_lbfgs_dots: null pointer or empty vector");
  VS_REQUIRE(m >= 0 && m <= VS_LBFGS_MAX_HIST, VS_ERR_UNSUPPORTED, "vs_lbfgs_dots: history %d outside 0..%d", m, VS_LBFGS_MAX_HIST);
  VS_REQUIRE(m == 0 || (hist && s_slots_host && y_slots_host && hist_stride >= n), VS_ERR_INVALID, "vs_lbfgs_dots: history arguments");
  VS_REQUIRE(workspace && workspace_bytes >= vs_lbfgs_workspace(n, m), VS_ERR_WORKSPACE, "vs_lbfgs_dots: workspace too small");
  VS_REQUIRE(((uintptr_t)workspace & 7) == 0, VS_ERR_INVALID, "vs_lbfgs_dots: workspace must be 8-byte aligned");
  VS_REQUIRE(n < (1ll << 40), VS_ERR_UNSUPPORTED, "vs_lbfgs_dots: vector too long");
  Slots sl;
  for (int i = 0; i < m; ++i) { sl.s[i] = s_slots_host[i]; sl.s[m + i] = y_slots_host[i]; }
  const int nh = 2 * m, nout = kBase + 3 * nh;
  const int chunks = num_chunks(n);
  double* part = reinterpret_cast<double*>(workspace);
  int rc = hist_f32 ? launch_dots<float>(n, g, g_prev, s_new, y_out, hist, hist_stride, sl, nh, nout, part, chunks, stream)
                    : launch_dots<double>(n, g, g_prev, s_new, y_out, hist, hist_stride, sl, nh, nout, part, chunks, stream);
  if (rc) return rc;
  VS_LAUNCH(dots_reduce_kernel, (unsigned)nout, 128, 0, stream, part, chunks, nout, out, (vs_lbfgs_dev*)nullptr, (vs_lbfgs_cdev*)nullptr);
  return VS_OK;
}

extern "C" int vs_lbfgs_direction(int64_t n, const double* g, const void* hist, int64_t hist_stride, int hist_f32,
                                  const int32_t* s_slots_host, const int32_t* y_slots_host, int m, const double* coef_host,
                                  double t, double* x, void* s_out, double* dmax_out, void* stream) {
  VS_REQUIRE(n > 0 && g && s_out && dmax_out && coef_host, VS_ERR_INVALID, "vs_lbfgs_direction: null pointer or empty vector");
  VS_REQUIRE(m >= 0 && m <= VS_LBFGS_MAX_HIST, VS_ERR_UNSUPPORTED, "vs_lbfgs_direction: history %d outside 0..%d", m, VS_LBFGS_MAX_HIST);
  VS_REQUIRE(m == 0 || (hist && s_slots_host && y_slots_host && hist_stride >= n), VS_ERR_INVALID, "vs_lbfgs_direction: history arguments");
  Slots sl;
  for (int i = 0; i < m; ++i) { sl.s[i] = s_slots_host[i]; sl.s[m + i] = y_slots_host[i]; }
  VS_CHECK_CUDA(cudaMemsetAsync(dmax_out, 0, sizeof(double), (cudaStream_t)stream));
  long long blocks = ceil_div(n, 256 * 4);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  const bool vec = ((((uintptr_t)g | (uintptr_t)hist | (uintptr_t)x | (uintptr_t)s_out) & 15) == 0) && hist_stride % 4 == 0;
  unsigned long long* dm = reinterpret_cast<unsigned long long*>(dmax_out);
  if (hist_f32) {
    CoefT<float> cf;
    for (int i = 0; i < 2 * m + 1; ++i) cf.c[i] = (float)coef_host[i];
    if (vec) { VS_LAUNCH((direction_kernel<float, true>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const float*)hist, (long long)hist_stride, sl, 2 * m, cf, coef_host[0], t, x, (float*)s_out, dm, (vs_lbfgs_dev*)nullptr, (vs_lbfgs_cdev*)nullptr); }
    else { VS_LAUNCH((direction_kernel<float, false>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const float*)hist, (long long)hist_stride, sl, 2 * m, cf, coef_host[0], t, x, (float*)s_out, dm, (vs_lbfgs_dev*)nullptr, (vs_lbfgs_cdev*)nullptr); }
  } else {
    CoefT<double> cf;
    for (int i = 0; i < 2 * m + 1; ++i) cf.c[i] = coef_host[i];
    if (vec) { VS_LAUNCH((direction_kernel<double, true>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const double*)hist, (long long)hist_stride, sl, 2 * m, cf, coef_host[0], t, x, (double*)s_out, dm, (vs_lbfgs_dev*)nullptr, (vs_lbfgs_cdev*)nullptr); }
    else { VS_LAUNCH((direction_kernel<double, false>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const double*)hist, (long long)hist_stride, sl, 2 * m, cf, coef_host[0], t, x, (double*)s_out, dm, (vs_lbfgs_dev*)nullptr, (vs_lbfgs_cdev*)nullptr); }
  }
  return VS_OK;
}


// ------------------------------------------------------------------ device-driven entry points
extern "C" int vs_lbfgs_dev_init_host(vs_lbfgs_dev* st, int n_slots) {
  if (!st || n_slots < 4 || n_slots > 2 * VS_LBFGS_MAX_HIST + 8) return VS_ERR_INVALID;
  memset(st, 0, sizeof(*st));
  st->H_diag = 1.0;
  st->n_free = n_slots;
  for (int i = 0; i < n_slots; ++i) st->free_slots[i] = n_slots - 1 - i;     // slot 0 is handed out first
  return VS_OK;
}

extern "C" size_t vs_lbfgs_dev_state_bytes(void) { return sizeof(vs_lbfgs_dev); }

extern "C" size_t vs_lbfgs_dev_workspace(int64_t n) {
  if (n <= 0) return 0;
  return (size_t)num_chunks(n) * (kBase + 6 * (size_t)VS_LBFGS_MAX_HIST) * sizeof(double) + 64;
}

extern "C" int vs_lbfgs_dev_dots(vs_lbfgs_dev* state, int64_t n, const double* g, const double* g_prev, void* hist,
                                 int64_t hist_stride, int hist_f32, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(state && n > 0 && g && g_prev && hist && hist_stride >= n, VS_ERR_INVALID, "vs_lbfgs_dev_dots: bad arguments");
  VS_REQUIRE(workspace && workspace_bytes >= vs_lbfgs_dev_workspace(n) && ((uintptr_t)workspace & 7) == 0, VS_ERR_WORKSPACE,
             "vs_lbfgs_dev_dots: workspace too small or misaligned");
  VS_REQUIRE(n < (1ll << 40), VS_ERR_UNSUPPORTED, "vs_lbfgs_dev_dots: vector too long");
  const int nout = kBase + 6 * VS_LBFGS_MAX_HIST;
  const int chunks = num_chunks(n);
  double* part = reinterpret_cast<double*>(workspace);
  Slots sl;
  sl.s[0] = 0;
  int rc = hist_f32 ? launch_dots<float>(n, g, g_prev, nullptr, nullptr, hist, hist_stride, sl, 0, nout, part, chunks, stream, state)
                    : launch_dots<double>(n, g, g_prev, nullptr, nullptr, hist, hist_stride, sl, 0, nout, part, chunks, stream, state);
  if (rc) return rc;
  VS_LAUNCH(dots_reduce_kernel, (unsigned)nout, 128, 0, stream, part, chunks, nout, (double*)nullptr, state, (vs_lbfgs_cdev*)nullptr);
  return VS_OK;
}

extern "C" int vs_lbfgs_dev_update(vs_lbfgs_dev* state, const double* loss, double lr, double tolerance_grad, double tolerance_change,
                                   int max_eval, int history_size, int first_eval, void* stream) {
  VS_REQUIRE(state && loss, VS_ERR_INVALID, "vs_lbfgs_dev_update: null pointer");
  VS_REQUIRE(history_size >= 1 && history_size <= VS_LBFGS_MAX_HIST, VS_ERR_UNSUPPORTED, "vs_lbfgs_dev_update: history_size outside 1..%d",
             VS_LBFGS_MAX_HIST);
  VS_LAUNCH(update_kernel, 1, 128, 0, stream, state, loss, lr, tolerance_grad, tolerance_change, max_eval, history_size, first_eval);
  return VS_OK;
}

extern "C" int vs_lbfgs_dev_direction(vs_lbfgs_dev* state, int64_t n, const double* g, void* hist, int64_t hist_stride, int hist_f32,
                                      double* x, void* stream) {
  VS_REQUIRE(state && n > 0 && g && hist && x && hist_stride >= n, VS_ERR_INVALID, "vs_lbfgs_dev_direction: bad arguments");
  long long blocks = ceil_div(n, 256 * 4);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  Slots sl;
  sl.s[0] = 0;
  const bool vec = ((((uintptr_t)g | (uintptr_t)hist | (uintptr_t)x) & 15) == 0) && hist_stride % 4 == 0;
  if (hist_f32) {
    CoefT<float> cf;
    cf.c[0] = 0.f;
    if (vec) { VS_LAUNCH((direction_kernel<float, true>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const float*)hist, (long long)hist_stride, sl, 0, cf, 0.0, 0.0, x, (float*)nullptr, (unsigned long long*)nullptr, state, (vs_lbfgs_cdev*)nullptr); }
    else { VS_LAUNCH((direction_kernel<float, false>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const float*)hist, (long long)hist_stride, sl, 0, cf, 0.0, 0.0, x, (float*)nullptr, (unsigned long long*)nullptr, state, (vs_lbfgs_cdev*)nullptr); }
  } else {
    CoefT<double> cf;
    cf.c[0] = 0.0;
    if (vec) { VS_LAUNCH((direction_kernel<double, true>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const double*)hist, (long long)hist_stride, sl, 0, cf, 0.0, 0.0, x, (double*)nullptr, (unsigned long long*)nullptr, state, (vs_lbfgs_cdev*)nullptr); }
    else { VS_LAUNCH((direction_kernel<double, false>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const double*)hist, (long long)hist_stride, sl, 0, cf, 0.0, 0.0, x, (double*)nullptr, (unsigned long long*)nullptr, state, (vs_lbfgs_cdev*)nullptr); }
  }
  return VS_OK;
}


// ------------------------------------------------------------------ compact-history entry points (float64 history only)
extern "C" int vs_lbfgs_cdev_init_host(vs_lbfgs_cdev* st) {
  if (!st) return VS_ERR_INVALID;
  memset(st, 0, sizeof(*st));
  st->H_diag = 1.0;
  return VS_OK;
}

extern "C" size_t vs_lbfgs_cdev_state_bytes(void) { return sizeof(vs_lbfgs_cdev); }

extern "C" size_t vs_lbfgs_cdev_workspace(int64_t n) {
  if (n <= 0) return 0;
  return (size_t)num_chunks(n) * (kBase + 3 * (size_t)VS_LBFGS_CMAX) * sizeof(double) + 64;
}

extern "C" int vs_lbfgs_cdev_dots(vs_lbfgs_cdev* state, int64_t n, const double* g, const double* g_prev, double* hist,
                                  int64_t hist_stride, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(state && n > 0 && g && g_prev && hist && hist_stride >= n, VS_ERR_INVALID, "vs_lbfgs_cdev_dots: bad arguments");
  VS_REQUIRE(workspace && workspace_bytes >= vs_lbfgs_cdev_workspace(n) && ((uintptr_t)workspace & 7) == 0, VS_ERR_WORKSPACE,
             "vs_lbfgs_cdev_dots: workspace too small or misaligned");
  VS_REQUIRE(n < (1ll << 40), VS_ERR_UNSUPPORTED, "vs_lbfgs_cdev_dots: vector too long");
  const int nout = kBase + 3 * VS_LBFGS_CMAX;
  const int chunks = num_chunks(n);
  double* part = reinterpret_cast<double*>(workspace);
  Slots sl;
  sl.s[0] = 0;
  int rc = launch_dots<double>(n, g, g_prev, nullptr, nullptr, hist, hist_stride, sl, 0, nout, part, chunks, stream, nullptr, state);
  if (rc) return rc;
  VS_LAUNCH(dots_reduce_kernel, (unsigned)nout, 128, 0, stream, part, chunks, nout, (double*)nullptr, (vs_lbfgs_dev*)nullptr, state);
  return VS_OK;
}

extern "C" int vs_lbfgs_cdev_update(vs_lbfgs_cdev* state, const double* loss, double lr, double tolerance_grad, double tolerance_change,
                                    int max_eval, int first_eval, void* stream) {
  VS_REQUIRE(state && loss, VS_ERR_INVALID, "vs_lbfgs_cdev_update: null pointer");
  VS_LAUNCH(update_compact_kernel, 1, 128, 0, stream, state, loss, lr, tolerance_grad, tolerance_change, max_eval, first_eval);
  return VS_OK;
}

extern "C" int vs_lbfgs_cdev_direction(vs_lbfgs_cdev* state, int64_t n, const double* g, double* hist, int64_t hist_stride, double* x,
                                       void* stream) {
  VS_REQUIRE(state && n > 0 && g && hist && x && hist_stride >= n, VS_ERR_INVALID, "vs_lbfgs_cdev_direction: bad arguments");
  long long blocks = ceil_div(n, 256 * 4);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  Slots sl;
  sl.s[0] = 0;
  const bool vec = ((((uintptr_t)g | (uintptr_t)hist | (uintptr_t)x) & 15) == 0) && hist_stride % 4 == 0;
  CoefT<double> cf;
  cf.c[0] = 0.0;
  if (vec) { VS_LAUNCH((direction_kernel<double, true>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const double*)hist, (long long)hist_stride, sl, 0, cf, 0.0, 0.0, x, (double*)nullptr, (unsigned long long*)nullptr, (vs_lbfgs_dev*)nullptr, state); }
  else { VS_LAUNCH((direction_kernel<double, false>), (unsigned)blocks, 256, 0, stream, (long long)n, g, (const double*)hist, (long long)hist_stride, sl, 0, cf, 0.0, 0.0, x, (double*)nullptr, (unsigned long long*)nullptr, (vs_lbfgs_dev*)nullptr, state); }
  return VS_OK;
}
