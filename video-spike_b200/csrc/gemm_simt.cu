// CUDA-core (fp32 FMA) GEMM with arbitrary operand strides.  It serves the five small layers
// of the `Linear` MLP (src/model/linear.py:29-32,47-53 -- 82 k weights, launch-bound, not worth a
// tensor-core pipeline), every backward-data / backward-weight product of those layers, and it is
// the on-device cross-check for the tcgen05 engine (tests compare the two bit patterns' sums).
#include <cuda_fp16.h>
#include "common.cuh"
#include "gemm.h"

namespace vs {
namespace simt {

constexpr int TM = 64, TN = 64, TK = 16;

struct DevOperand {
  const void* ptr;
  int type;
  long long s_i, s_k;
  int planes;
  long long plane_stride;
};

__device__ __forceinline__ float load_elem(const DevOperand& o, long long i, long long k) {
  const long long off = i * o.s_i + k * o.s_k;
  if (o.type == F32) return reinterpret_cast<const float*>(o.ptr)[off];
  if (o.type == U8) return (float)reinterpret_cast<const uint8_t*>(o.ptr)[off];
  float v = 0.f;
  if (o.type == F16) {
    const __half* p = reinterpret_cast<const __half*>(o.ptr);
    for (int pl = 0; pl < o.planes; ++pl) v += __half2float(p[off + pl * o.plane_stride]);
    return v;
  }
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(o.ptr);
  for (int pl = 0; pl < o.planes; ++pl) v += __bfloat162float(p[off + pl * o.plane_stride]);
  return v;
}

struct KParams {
  DevOperand A, B;
  long long M, N, K;
  float* C;
  long long ldc, split_stride;
  long long k_per_split;
  const float* bias;
  int relu;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const KParams p) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * TM, n0 = (long long)blockIdx.x * TN;
  const long long k0 = (long long)blockIdx.z * p.k_per_split;
  const long long k1 = min(p.K, k0 + p.k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (p.A.s_k == 1), b_kfast = (p.B.s_k == 1);
  for (long long kk = k0; kk < k1; kk += TK) {
    for (int e = tid; e < TM * TK; e += 256) {
      int mm, k;
      if (a_kfast) { mm = e / TK; k = e % TK; } else { k = e / TM; mm = e % TM; }
      const long long m = m0 + mm, kq = kk + k;
      As[k][mm] = (m < p.M && kq < k1) ? load_elem(p.A, m, kq) : 0.f;
    }
    for (int e = tid; e < TN * TK; e += 256) {
      int nn, k;
      if (b_kfast) { nn = e / TK; k = e % TK; } else { k = e / TN; nn = e % TN; }
      const long long n = n0 + nn, kq = kk + k;
      Bs[k][nn] = (n < p.N && kq < k1) ? load_elem(p.B, n, kq) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = p.C + (long long)blockIdx.z * p.split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty + 16 * i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx + 16 * j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      C[m * p.ldc + n] = v;
    }
  }
}

static DevOperand to_dev(const Operand& o) { return DevOperand{o.ptr, o.type, o.s_i, o.s_k, o.planes, o.plane_stride}; }

int gemm(const GemmDesc& g, cudaStream_t stream) {
  VS_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0 && g.C, VS_ERR_INVALID, "simt gemm: bad shape");
  int splits = g.splits > 0 ? g.splits : 1;
  long long kps = round_up(ceil_div(g.K > 0 ? g.K : 1, splits), TK);
  splits = (int)ceil_div(g.K > 0 ? g.K : 1, kps);
  VS_REQUIRE(splits == 1 || (!g.bias && !g.relu), VS_ERR_INVALID, "simt gemm: bias/relu only without split-K");
  KParams p{to_dev(g.A), to_dev(g.B), g.M, g.N, g.K, g.C, g.ldc, g.split_stride, kps, g.bias, g.relu};
  dim3 grid((unsigned)ceil_div(g.N, TN), (unsigned)ceil_div(g.M, TM), (unsigned)splits);
  VS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, VS_ERR_UNSUPPORTED, "simt gemm: grid too large");
  VS_LAUNCH(gemm_simt_kernel, grid, 256, 0, stream, p);
  return VS_OK;
}

}  // namespace simt

// ------------------------------------------------------------------------------------------
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long long split_stride, long long ld_o,
                                     long long ld_b, const float* __restrict__ bias, float* __restrict__ y, long long batch,
                                     long long out_dim, int relu) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * out_dim) return;
  const long long b = idx / out_dim, o = idx % out_dim;
  float s = 0.f;
  for (int i = 0; i < splits; ++i) s += part[(long long)i * split_stride + o * ld_o + b * ld_b];  // fixed order
  if (bias) s += bias[o];
  if (relu) s = fmaxf(s, 0.f);
  y[b * out_dim + o] = s;
}

int splitk_reduce_bias_act(const float* part, int splits, long long split_stride, long long ld_o, long long ld_b,
                           const float* bias, float* y, long long batch, long long out_dim, int relu, cudaStream_t stream) {
  const long long n = batch * out_dim;
  VS_LAUNCH(splitk_reduce_kernel, (unsigned)ceil_div(n, 256), 256, 0, stream, part, splits, split_stride, ld_o, ld_b, bias,
            y, batch, out_dim, relu);
  return VS_OK;
}

}  // namespace vs
