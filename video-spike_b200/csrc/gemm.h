// Internal GEMM interfaces shared by the translation units of libvs_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vs {

constexpr int kMaxPass = 6;   // plane products accumulated into one tile (3 planes -> 6 products)
constexpr int kMaxBN = 448;   // widest accumulator tile the tcgen05 kernel keeps in TMEM

// Layout of the fp32 copy of U that the dense-forward generator streams (rank 3): for every 64-feature block kb and neuron n
// a 768-byte record of 6 x 8 float4; float4 (i, q) holds values 4i..4i+3 of the 24 = 8 features x 3 ranks that generator
// lane q (features 8q..8q+7 of the block) needs, so the 8 lanes of a neuron read 128 contiguous bytes per load.
__host__ __device__ inline long long u32g_index(long long n, long long c, int j, long long N) {
  const long long flat = (c & 7) * 3 + j;
  return (((((c >> 6) * N + n) * 6 + (flat >> 2)) * 8 + ((c & 63) >> 3)) << 2) + (flat & 3);
}

namespace tc {
// K-major operand: rows x k elements, pitch ld (elements), optionally several residual planes
struct Operand {
  const void* ptr = nullptr;
  long long rows = 0, k = 0, ld = 0;
  int planes = 1;
  long long plane_stride = 0;  // elements
};
struct GemmDesc {
  Operand A, B;            // C[M,N] = sum_pass A_pa * B_pb^T
  long long M = 0, N = 0, K = 0;
  float* C = nullptr;
  long long ldc = 0;
  long long split_stride = 0;  // elements between split-K partials
  int splits = 1;              // requested split-K factor
  int* splits_out = nullptr;   // actual factor used
  int BN = 0;                  // 0 = choose
  bool tf32 = false;
  bool f16 = false;            // 16-bit operands are IEEE half instead of bf16 (ignored with tf32)
  int n_pass = 1;
  int pa[kMaxPass] = {0, 0, 0, 0, 0, 0};
  int pb[kMaxPass] = {0, 0, 0, 0, 0, 0};
  // Tail-wave balancing (only when splits == 1): with balance_ws (balance_ws_bytes() bytes) the tiles of the
  // last, partially filled wave are split along K over the idle SMs and summed back into C by a small
  // ordered reduction, instead of costing a whole wave.
  void* balance_ws = nullptr;
};
// Dense-per-time-bin RRR backward (rrr_bwd_dense_kernel): G[c, j*Npad + n] = sum_t V[j,t] sum_k Xb[c, t*Kp + k] R[n, t*Kp + k]
struct DenseBwdDesc {
  const void* Xb = nullptr;   // (C1, ldr) 16-bit, column t*Kp + k, zero for K <= k < Kp
  const void* R = nullptr;    // (Npad, ldr) 16-bit, same columns
  long long C1 = 0, K = 0, Kp = 0, T = 0, Npad = 0, ldr = 0;
  int r = 0;
  bool f16 = false;
  const double* V = nullptr;  // (r, T) device
  float* G = nullptr;         // (C1, ldg)
  long long ldg = 0;
  // exact-operand mode (vs_rrr_closure_exact): Xb holds the EXACT integers frame - round(mean) and R is split into
  // r_planes = 2 residual planes (hi, lo) at r_plane_stride elements; the z-score scale 1/std[t,c] is applied per
  // (feature row, time bin) to the rank-one weights: G_j += V[j,t] * scaleT[c*ldt + t] * D_t
  int r_planes = 1;
  long long r_plane_stride = 0;
  const float* scaleT = nullptr;
  long long ldt = 0;
  // dV pass (dense-forward mode, where no Z exists): the same D_t tiles contracted with U instead of accumulated with V:
  //   dvpart[((cta * 8 + w) * T + t) * 3 + j], cta < rrr_bwd_dense_ctas(C1), w < 8  (G is not written)
  const float* dv_U32 = nullptr;  // fp32 copy of U, u32g_index layout
  long long dv_ldu = 0, dv_N = 0;
  double* dvpart = nullptr;
};
inline long long rrr_bwd_dense_ctas(long long C1) { return 2 * ((C1 + 255) / 256); }
// Dense-per-time-bin RRR forward (rrr_fwd_dense_pair_kernel): Y[(t,k), n] = sum_c Xc[(t,k), c] * (U[n,c,:] . V[:,t]) * isd[t,c],
// the coefficient tiles generated on chip from U (hi + lo half planes), three time bins per CTA pair
struct DenseFwdDesc {
  const void* Xc = nullptr;     // (K*T, ldc) IEEE half, EXACT integers frame - round(mean), row t*K + k
  long long K = 0, T = 0, C1 = 0, N = 0, Npad = 0, ldc = 0;
  const float* U32 = nullptr;   // fp32 copy of U in the u32g_index layout (zero for C1 <= c < ldu)
  const float* isd = nullptr;   // [T][ldu] fp32, zero for C1 <= c < ldu
  long long ldu = 0;            // multiple of 64, >= C1
  const double* V = nullptr;    // (3, T)
  const float* bscale = nullptr;  // [T] power-of-two scales of the generated tiles
  float* Y = nullptr;           // (K*T, ldy) fp32; with splits > 1: `splits` partial matrices at split_stride elements
  long long ldy = 0;
  int splits = 1;               // accumulation runs over the feature dimension (bounds the fp32 TMEM truncation drift)
  long long split_stride = 0;
  int* splits_out = nullptr;
};
bool rrr_fwd_dense_supported(const DenseFwdDesc& g);
int rrr_fwd_dense(const DenseFwdDesc& g, cudaStream_t stream);
bool rrr_bwd_dense_supported(const DenseBwdDesc& g);
int rrr_bwd_dense(const DenseBwdDesc& g, cudaStream_t stream);
size_t balance_ws_bytes();
bool gemm_supported(const GemmDesc& g);
int pick_bn(long long N);
int gemm_tn(const GemmDesc& g, cudaStream_t stream);
}  // namespace tc

namespace simt {
enum ElemType { F32 = 0, U8 = 1, BF16 = 2, F16 = 3 };
// generic strided operand: element (i, k) at ptr[i*s_i + k*s_k] (+ plane*plane_stride summed over planes)
struct Operand {
  const void* ptr = nullptr;
  int type = F32;
  long long s_i = 0, s_k = 0;
  int planes = 1;
  long long plane_stride = 0;
};
struct GemmDesc {
  Operand A, B;            // C[m,n] = sum_k A(m,k) * B(n,k)
  long long M = 0, N = 0, K = 0;
  float* C = nullptr;
  long long ldc = 0;
  long long split_stride = 0;
  int splits = 1;
  const float* bias = nullptr;  // per-n bias, only when splits == 1
  int relu = 0;                 // only when splits == 1
};
int gemm(const GemmDesc& g, cudaStream_t stream);
}  // namespace simt

// y[b,o] = act(sum_s part[s][o*ld_o + b*ld_b] + bias[o]) -- deterministic split-K reduction
int splitk_reduce_bias_act(const float* part, int splits, long long split_stride, long long ld_o, long long ld_b,
                           const float* bias, float* y, long long batch, long long out_dim, int relu, cudaStream_t stream);

}  // namespace vs
