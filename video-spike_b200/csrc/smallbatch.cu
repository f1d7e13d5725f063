// Small-batch (B <= 32) layer kernels of the `Linear` MLP (src/model/linear.py:29-32,47-53 and their
// autograd, src/trainer/base.py:150-151).  At the reference batch size (16) every one of these
// products is a weight-streaming problem, not a GEMM: the layer's weights are read once per pass
// with 128-bit coalesced loads while the handful of batch rows sits in shared memory / registers.
//   fwd_kernel      y = act(x W^T + b)            warp per R output rows, x staged in smem
//   dx_part_kernel  dx = g W (partials over row ranges) + dx_reduce_kernel (ordered sum, ReLU mask of
//                   the layer below folded in: no separate threshold-backward pass)
//   dw_kernel       dW = g^T x formed in registers and consumed by AdamW on the spot (MODE 0: W, m, v
//                   updated in place, bias included) or stored (MODE 1: autograd route)
// All reductions run in a fixed order: results are bit-reproducible run to run.
#include "common.cuh"
#include "adam.cuh"
#include "smallbatch.h"

namespace vs {
namespace sb {

constexpr int kMaxSmem = 200 * 1024;

// lane l ends up with the sum over the warp of v[l] (31 shuffles for 32 values)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------ forward
template <int BT, int R>
__global__ void __launch_bounds__(256) fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                  const float* __restrict__ bias, float* __restrict__ y, int batch, int in_dim,
                                                  int out_dim, int relu) {
  static_assert((R * BT) % 32 == 0, "R*BT must be a multiple of the warp size");
  extern __shared__ float xs[];  // [BT][in_dim], rows past `batch` zeroed
  const int n4 = BT * in_dim / 4;
  for (int e = threadIdx.x; e < n4; e += 256) {
    const int b = (e * 4) / in_dim, i = (e * 4) % in_dim;
    reinterpret_cast<float4*>(xs)[e] =
        b < batch ? __ldg(reinterpret_cast<const float4*>(x + (long long)b * in_dim + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = (blockIdx.x * 8 + warp) * R;
  if (o0 >= out_dim) return;
  float acc[R * BT];
#pragma unroll
  for (int i = 0; i < R * BT; ++i) acc[i] = 0.f;
  for (int i = lane * 4; i < in_dim; i += 128) {
    float4 w[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      w[r] = o0 + r < out_dim ? ld_stream_f4(reinterpret_cast<const float4*>(W + (long long)(o0 + r) * in_dim + i))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + b * in_dim + i);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a = acc[r * BT + b];
        a = fmaf(w[r].x, xv.x, a); a = fmaf(w[r].y, xv.y, a); a = fmaf(w[r].z, xv.z, a); a = fmaf(w[r].w, xv.w, a);
        acc[r * BT + b] = a;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < R * BT / 32; ++g) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = acc[g * 32 + i];
    float s = warp_transpose_sum32(v);
    const int idx = g * 32 + lane, r = idx / BT, b = idx % BT;
    if (b < batch && o0 + r < out_dim) {
      if (bias) s += bias[o0 + r];
      if (relu) s = fmaxf(s, 0.f);
      y[(long long)b * out_dim + o0 + r] = s;
    }
  }
}

// ------------------------------------------------------------------ backward data
// block (bx, by): column strip bx (4*TX columns), output rows [by*rpb, (by+1)*rpb).  thread (tx, ty) owns
// 4 columns and every TY-th row of the range; the TY row lanes are then summed in a fixed order.
template <int BT>
__global__ void __launch_bounds__(256) dx_part_kernel(const float* __restrict__ g, const float* __restrict__ W,
                                                      float* __restrict__ part, int batch, int in_dim, int out_dim, int log2_tx,
                                                      int rpb) {
  extern __shared__ float sm[];
  float* gs = sm;                          // [rpb][BT]
  float* red = sm + (size_t)rpb * BT;      // [TY][BT][4*TX]
  const int TX = 1 << log2_tx, TY = 256 >> log2_tx;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> log2_tx;
  const int o0 = blockIdx.y * rpb;
  const int nrows = min(rpb, out_dim - o0);
  for (int e = threadIdx.x; e < rpb * BT; e += 256) {
    const int b = e / rpb, rr = e % rpb;
    gs[rr * BT + b] = (b < batch && rr < nrows) ? g[(long long)b * out_dim + o0 + rr] : 0.f;
  }
  __syncthreads();
  const int col = (blockIdx.x * TX + tx) * 4;
  const bool active = col < in_dim;
  float acc[BT][4];
#pragma unroll
  for (int b = 0; b < BT; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
  if (active) {
#pragma unroll 2
    for (int rr = ty; rr < nrows; rr += TY) {
      const float4 w = ld_stream_f4(reinterpret_cast<const float4*>(W + (long long)(o0 + rr) * in_dim + col));
      const float4* g4 = reinterpret_cast<const float4*>(gs + rr * BT);
#pragma unroll
      for (int b4 = 0; b4 < BT / 4; ++b4) {
        const float4 d = g4[b4];
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int b = b4 * 4 + q;
          acc[b][0] = fmaf(dd[q], w.x, acc[b][0]); acc[b][1] = fmaf(dd[q], w.y, acc[b][1]);
          acc[b][2] = fmaf(dd[q], w.z, acc[b][2]); acc[b][3] = fmaf(dd[q], w.w, acc[b][3]);
        }
      }
    }
  }
  float* out = part + (long long)blockIdx.y * batch * in_dim;
  if (TY == 1) {
    if (active)
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b < batch)
          *reinterpret_cast<float4*>(out + (long long)b * in_dim + col) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
    return;
  }
  const int cw = 4 * TX;
#pragma unroll
  for (int b = 0; b < BT; ++b)
    *reinterpret_cast<float4*>(red + ((size_t)ty * BT + b) * cw + tx * 4) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
  __syncthreads();
  if (!active) return;
  for (int b = ty; b < batch; b += TY) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t2 = 0; t2 < TY; ++t2) {
      const float4 v = *reinterpret_cast<const float4*>(red + ((size_t)t2 * BT + b) * cw + tx * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (long long)b * in_dim + col) = s;
  }
}

// dx = sum over the row-range partials (fixed order); mask by the forward output of the layer below
__global__ void __launch_bounds__(256) dx_reduce_kernel(const float* __restrict__ part, int gy, const float* __restrict__ act_prev,
                                                        float* __restrict__ dx, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < gy; ++p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part) + (long long)p * n4 + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (act_prev) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(act_prev) + i);
    s.x = a.x > 0.f ? s.x : 0.f; s.y = a.y > 0.f ? s.y : 0.f; s.z = a.z > 0.f ? s.z : 0.f; s.w = a.w > 0.f ? s.w : 0.f;
  }
  reinterpret_cast<float4*>(dx)[i] = s;
}

// ------------------------------------------------------------------ backward weights (+ AdamW)
// MODE 0: AdamW on W, m, v (and bias, mb, vb) in place, the gradient lives only in registers.
// MODE 1: gradient stored to dW / dbias (autograd route with a stock optimizer).
template <int BT, int MODE, bool kU8>
__global__ void __launch_bounds__(256) dw_kernel(const float* __restrict__ dy, const float* __restrict__ xf,
                                                 const uint8_t* __restrict__ xu, float* __restrict__ W, float* __restrict__ M,
                                                 float* __restrict__ V, float* __restrict__ dW, float* __restrict__ bias,
                                                 float* __restrict__ mb, float* __restrict__ vb, float* __restrict__ dbias,
                                                 int batch, long long in_dim, int out_dim, int log2_tx, int rpb,
                                                 const AdamConsts c) {
  extern __shared__ float dys[];  // [rpb][BT], zero padded past batch / past the last row
  const int TX = 1 << log2_tx, TY = 256 >> log2_tx;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> log2_tx;
  const int o0 = blockIdx.y * rpb;
  const int nrows = min(rpb, out_dim - o0);
  for (int e = threadIdx.x; e < rpb * BT; e += 256) {
    const int b = e / rpb, rr = e % rpb;
    dys[rr * BT + b] = (b < batch && rr < nrows) ? dy[(long long)b * out_dim + o0 + rr] : 0.f;
  }
  const long long col = ((long long)blockIdx.x * TX + tx) * 4;
  const bool active = col < in_dim;
  float x[BT][4];
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    if (active && b < batch) {
      if constexpr (kU8) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(xu + (long long)b * in_dim + col));
        x[b][0] = (float)(w & 0xff); x[b][1] = (float)((w >> 8) & 0xff);
        x[b][2] = (float)((w >> 16) & 0xff); x[b][3] = (float)(w >> 24);
      } else {
        const float4 w = __ldg(reinterpret_cast<const float4*>(xf + (long long)b * in_dim + col));
        x[b][0] = w.x; x[b][1] = w.y; x[b][2] = w.z; x[b][3] = w.w;
      }
    } else {
      x[b][0] = x[b][1] = x[b][2] = x[b][3] = 0.f;
    }
  }
  __syncthreads();
  if (active) {
    for (int rr = ty; rr < nrows; rr += TY) {
      const long long off = (long long)(o0 + rr) * in_dim + col;
      float4 pv, mv, vv;
      if constexpr (MODE == 0) {
        pv = ld_f4(reinterpret_cast<float4*>(W + off));
        mv = ld_f4(reinterpret_cast<float4*>(M + off));
        vv = ld_f4(reinterpret_cast<float4*>(V + off));
      }
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
      const float4* d4 = reinterpret_cast<const float4*>(dys + rr * BT);
#pragma unroll
      for (int b4 = 0; b4 < BT / 4; ++b4) {
        const float4 d = d4[b4];
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int b = b4 * 4 + q;
          g0 = fmaf(dd[q], x[b][0], g0); g1 = fmaf(dd[q], x[b][1], g1);
          g2 = fmaf(dd[q], x[b][2], g2); g3 = fmaf(dd[q], x[b][3], g3);
        }
      }
      if constexpr (MODE == 0) {
        adamw_elem(pv.x, mv.x, vv.x, g0, c); adamw_elem(pv.y, mv.y, vv.y, g1, c);
        adamw_elem(pv.z, mv.z, vv.z, g2, c); adamw_elem(pv.w, mv.w, vv.w, g3, c);
        st_stream_f4(reinterpret_cast<float4*>(W + off), pv);
        st_stream_f4(reinterpret_cast<float4*>(M + off), mv);
        st_stream_f4(reinterpret_cast<float4*>(V + off), vv);
      } else {
        st_stream_f4(reinterpret_cast<float4*>(dW + off), make_float4(g0, g1, g2, g3));
      }
    }
  }
  // bias: dbias[o] = sum_b dy[b,o], handled by the first column strip
  if (blockIdx.x == 0 && (MODE == 0 ? bias != nullptr : dbias != nullptr)) {
    for (int rr = threadIdx.x; rr < nrows; rr += 256) {
      float gb = 0.f;
#pragma unroll
      for (int b = 0; b < BT; ++b) gb += dys[rr * BT + b];
      if constexpr (MODE == 0) {
        float p = bias[o0 + rr], m = mb[o0 + rr], v = vb[o0 + rr];
        adamw_elem(p, m, v, gb, c);
        bias[o0 + rr] = p; mb[o0 + rr] = m; vb[o0 + rr] = v;
      } else {
        dbias[o0 + rr] = gb;
      }
    }
  }
}

// ------------------------------------------------------------------ host side
static int log2_tx_for(long long in_dim) {
  const long long cols4 = ceil_div(in_dim, 4);
  int l = 0;
  while ((1 << l) < cols4 && l < 8) ++l;
  return l;
}

bool supported(long long batch, long long in_dim, long long out_dim) {
  return batch >= 1 && batch <= 32 && in_dim % 4 == 0 && in_dim >= 4 && out_dim >= 1 && out_dim < (1ll << 31) / 64;
}
bool fwd_supported(long long batch, long long in_dim, long long out_dim) {
  if (!supported(batch, in_dim, out_dim)) return false;
  const int BT = batch <= 8 ? 8 : (batch <= 16 ? 16 : 32);
  return (size_t)BT * in_dim * 4 <= (size_t)kMaxSmem && in_dim < (1 << 24);
}

template <int BT, int R>
static int launch_fwd(const float* x, const float* W, const float* bias, float* y, int batch, int in_dim, int out_dim, int relu,
                      cudaStream_t st) {
  const size_t smem = (size_t)BT * in_dim * 4;
  VS_CHECK_CUDA(cudaFuncSetAttribute(fwd_kernel<BT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH((fwd_kernel<BT, R>), (unsigned)ceil_div(out_dim, 8 * R), 256, smem, st, x, W, bias, y, batch, in_dim, out_dim, relu);
  return VS_OK;
}

int fwd(const float* x, const float* W, const float* bias, float* y, long long batch, long long in_dim, long long out_dim,
        int relu, cudaStream_t st) {
  VS_REQUIRE(fwd_supported(batch, in_dim, out_dim), VS_ERR_UNSUPPORTED, "small-batch fwd: shape not covered");
  VS_REQUIRE((((uintptr_t)x | (uintptr_t)W) & 15) == 0, VS_ERR_INVALID, "small-batch fwd: x and W must be 16-byte aligned");
  if (batch <= 8) return launch_fwd<8, 4>(x, W, bias, y, (int)batch, (int)in_dim, (int)out_dim, relu, st);
  if (batch <= 16) return launch_fwd<16, 4>(x, W, bias, y, (int)batch, (int)in_dim, (int)out_dim, relu, st);
  return launch_fwd<32, 2>(x, W, bias, y, (int)batch, (int)in_dim, (int)out_dim, relu, st);
}

// row-range decomposition shared by dx and dw: enough blocks to fill the machine, ranges a multiple of TY
static void row_split(long long in_dim, long long out_dim, int* log2_tx, int* rpb, int* gx, int* gy) {
  const int l = log2_tx_for(in_dim);
  const int TX = 1 << l, TY = 256 >> l;
  const long long strips = ceil_div(in_dim, 4ll * TX);
  long long want = ceil_div(2ll * kNumSMs, strips);           // blocks along the rows
  long long r = ceil_div(out_dim, want);
  if (r < 4ll * TY) r = 4ll * TY;                               // at least 4 rows per thread: amortise the staging
  r = round_up(r, TY);
  if (r > out_dim) r = round_up(out_dim, TY);
  if (r > 1024) r = 1024;                                       // bounds the dy staging buffer
  *log2_tx = l; *rpb = (int)r; *gx = (int)strips; *gy = (int)ceil_div(out_dim, r);
}

size_t dx_workspace(long long batch, long long in_dim, long long out_dim) {
  if (!supported(batch, in_dim, out_dim)) return 0;
  int l, rpb, gx, gy;
  row_split(in_dim, out_dim, &l, &rpb, &gx, &gy);
  return (size_t)gy * batch * in_dim * sizeof(float);
}

template <int BT>
static int launch_dx(const float* g, const float* W, float* part, int batch, int in_dim, int out_dim, int l, int rpb, int gx, int gy,
                     cudaStream_t st) {
  const int TY = 256 >> l;
  const size_t smem = (size_t)rpb * BT * 4 + (TY > 1 ? (size_t)256 * BT * 16 : 0);
  VS_REQUIRE(smem <= (size_t)kMaxSmem + 24 * 1024, VS_ERR_UNSUPPORTED, "small-batch dx: staging exceeds shared memory");
  VS_CHECK_CUDA(cudaFuncSetAttribute(dx_part_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH((dx_part_kernel<BT>), dim3(gx, gy), 256, smem, st, g, W, part, batch, in_dim, out_dim, l, rpb);
  return VS_OK;
}

int dx(const float* g, const float* W, const float* act_prev, float* dx_out, long long batch, long long in_dim, long long out_dim,
       void* workspace, size_t workspace_bytes, cudaStream_t st) {
  VS_REQUIRE(supported(batch, in_dim, out_dim), VS_ERR_UNSUPPORTED, "small-batch dx: shape not covered");
  VS_REQUIRE(workspace && workspace_bytes >= dx_workspace(batch, in_dim, out_dim), VS_ERR_WORKSPACE,
             "small-batch dx: workspace too small (%zu < %zu)", workspace_bytes, dx_workspace(batch, in_dim, out_dim));
  VS_REQUIRE((((uintptr_t)W | (uintptr_t)workspace | (uintptr_t)dx_out | (uintptr_t)act_prev) & 15) == 0, VS_ERR_INVALID,
             "small-batch dx: buffers must be 16-byte aligned");
  int l, rpb, gx, gy, rc;
  row_split(in_dim, out_dim, &l, &rpb, &gx, &gy);
  float* part = reinterpret_cast<float*>(workspace);
  if (batch <= 8) rc = launch_dx<8>(g, W, part, (int)batch, (int)in_dim, (int)out_dim, l, rpb, gx, gy, st);
  else if (batch <= 16) rc = launch_dx<16>(g, W, part, (int)batch, (int)in_dim, (int)out_dim, l, rpb, gx, gy, st);
  else rc = launch_dx<32>(g, W, part, (int)batch, (int)in_dim, (int)out_dim, l, rpb, gx, gy, st);
  if (rc) return rc;
  const long long n4 = batch * in_dim / 4;
  VS_LAUNCH(dx_reduce_kernel, (unsigned)ceil_div(n4, 256), 256, 0, st, part, gy, act_prev, dx_out, n4);
  return VS_OK;
}

template <int BT, int MODE>
static int launch_dw(const float* dy, const float* xf, const uint8_t* xu, float* W, float* M, float* V, float* dW, float* bias,
                     float* mb, float* vb, float* dbias, int batch, long long in_dim, int out_dim, const AdamConsts& c,
                     cudaStream_t st) {
  int l, rpb, gx, gy;
  row_split(in_dim, out_dim, &l, &rpb, &gx, &gy);
  VS_REQUIRE(gy <= 65535, VS_ERR_UNSUPPORTED, "small-batch dw: too many row ranges");
  const size_t smem = (size_t)rpb * BT * 4;
  if (xu) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(dw_kernel<BT, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH((dw_kernel<BT, MODE, true>), dim3(gx, gy), 256, smem, st, dy, xf, xu, W, M, V, dW, bias, mb, vb, dbias, batch, in_dim,
              out_dim, l, rpb, c);
  } else {
    VS_CHECK_CUDA(cudaFuncSetAttribute(dw_kernel<BT, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH((dw_kernel<BT, MODE, false>), dim3(gx, gy), 256, smem, st, dy, xf, xu, W, M, V, dW, bias, mb, vb, dbias, batch, in_dim,
              out_dim, l, rpb, c);
  }
  return VS_OK;
}

static int check_dw(const float* dy, const float* xf, const uint8_t* xu, long long batch, long long in_dim, long long out_dim) {
  VS_REQUIRE(supported(batch, in_dim, out_dim), VS_ERR_UNSUPPORTED, "small-batch dw: shape not covered");
  VS_REQUIRE(dy && (xf || xu), VS_ERR_INVALID, "small-batch dw: null pointer");
  VS_REQUIRE(((uintptr_t)xf & 15) == 0 && ((uintptr_t)xu & 3) == 0, VS_ERR_INVALID, "small-batch dw: misaligned input");
  return VS_OK;
}

int dw_adamw(const float* dy, const float* xf, const uint8_t* xu, float* W, float* M, float* V, float* bias, float* mb, float* vb,
             long long batch, long long in_dim, long long out_dim, const vs_adamw_hyper& h, cudaStream_t st) {
  int rc = check_dw(dy, xf, xu, batch, in_dim, out_dim);
  if (rc) return rc;
  VS_REQUIRE(W && M && V && (!bias || (mb && vb)), VS_ERR_INVALID, "small-batch dw: null parameter/state pointer");
  VS_REQUIRE((((uintptr_t)W | (uintptr_t)M | (uintptr_t)V) & 15) == 0, VS_ERR_INVALID, "small-batch dw: misaligned parameters");
  const AdamConsts c = make_consts(h);
  if (batch <= 8) return launch_dw<8, 0>(dy, xf, xu, W, M, V, nullptr, bias, mb, vb, nullptr, (int)batch, in_dim, (int)out_dim, c, st);
  if (batch <= 16) return launch_dw<16, 0>(dy, xf, xu, W, M, V, nullptr, bias, mb, vb, nullptr, (int)batch, in_dim, (int)out_dim, c, st);
  return launch_dw<32, 0>(dy, xf, xu, W, M, V, nullptr, bias, mb, vb, nullptr, (int)batch, in_dim, (int)out_dim, c, st);
}

int dw_store(const float* dy, const float* xf, const uint8_t* xu, float* dW, float* dbias, long long batch, long long in_dim,
             long long out_dim, cudaStream_t st) {
  int rc = check_dw(dy, xf, xu, batch, in_dim, out_dim);
  if (rc) return rc;
  VS_REQUIRE(dW && ((uintptr_t)dW & 15) == 0, VS_ERR_INVALID, "small-batch dw: dW null or misaligned");
  const AdamConsts c{};
  if (batch <= 8) return launch_dw<8, 1>(dy, xf, xu, nullptr, nullptr, nullptr, dW, nullptr, nullptr, nullptr, dbias, (int)batch, in_dim, (int)out_dim, c, st);
  if (batch <= 16) return launch_dw<16, 1>(dy, xf, xu, nullptr, nullptr, nullptr, dW, nullptr, nullptr, nullptr, dbias, (int)batch, in_dim, (int)out_dim, c, st);
  return launch_dw<32, 1>(dy, xf, xu, nullptr, nullptr, nullptr, dW, nullptr, nullptr, nullptr, dbias, (int)batch, in_dim, (int)out_dim, c, st);
}

}  // namespace sb
}  // namespace vs
