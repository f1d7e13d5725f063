// Reduced-rank regression closure (src/model/rrr.py:79-155,165-175) on device.
//
// Factorised evaluation (DESIGN.md "RRR closure"; oracle/rrr_oracle.py loss_and_grad_lowrank):
//   Z[(k,t),(j,n)] = sum_c X[k,t,c] U[n,c,j]                    GEMM-F  (tcgen05, bf16 planes)
//   yhat = sum_j V[j,t] Z[..(j,n)] + xl * b[n,t];  R = yhat - y   epilogue-F (+ RV = R (x) V, dV partials)
//   Gacc[c,(j,n)] = sum_(k,t) X[k,t,c] RV[(j,n),(k,t)]           GEMM-B  (tcgen05, bf16 planes)
//   dU = 2 Gacc + 2 l2 U (V V^T);  dV = 2 sum R Z + 2 l2 (U^T U) V;  db = 2 sum_k xl R + 2 l2 b
// beta = cat(U@V, b) (rrr.py:79-96) and its gradient are never materialised.
// Operand rows are TIME-MAJOR: row d = t*K + k, so the K trials of one time bin are contiguous and every
// reduction over trials (db, SSE, dV) is a reduction over adjacent rows done inside the epilogue block.
// All reductions are ordered (no floating-point atomics): results are bit-reproducible run to run.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "gemm.h"

namespace vs {
namespace rrr {

constexpr int kMaxR = 8;
constexpr double kExactUScale = 256.0;   // exact-operand mode: U planes are stored as 256 * U (see closure_impl)

// 16-bit operand encodings (vs_rrr_dims.fmt): bf16 or IEEE half, round to nearest even
__device__ __forceinline__ uint16_t enc16(float v, int fmt) {
  return fmt == VS_OPERAND_F16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float dec16(uint16_t b, int fmt) {
  return fmt == VS_OPERAND_F16 ? __half2float(__ushort_as_half(b)) : __uint_as_float((uint32_t)b << 16);
}

// split v into `planes` 16-bit residual planes: v ~= p0 + p1 + p2
__device__ __forceinline__ void split_planes(double v, int planes, int fmt, uint16_t out[3]) {
  if (planes == 1) {  // fast path: one rounding, no fp64 arithmetic
    out[0] = enc16((float)v, fmt); out[1] = 0; out[2] = 0;
    return;
  }
  double rem = v;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    if (p < planes) {
      const uint16_t b = enc16((float)rem, fmt);
      out[p] = b;
      rem -= (double)dec16(b, fmt);
    } else {
      out[p] = 0;
    }
  }
}

// ------------------------------------------------------------------ pack X (R0 tail / RRRGD input)
// One tile of 32 trials (of ONE time bin t) x 32 features per block.  Source: X[(k*T + t), c] (fp64 path) or
// frames[k, sorted_idx[t], c] (uint8 path).  Destination rows are time-major d = t*K + k: Xa[d][c] is written
// directly (coalesced in c) and Xb[c][t*Kp + k] through a shared-memory transpose (coalesced in k).  In Xb every
// time bin is padded to Kp = K rounded up to 16 trials (zeros, pad_zero_kernel): a bin then starts on a 32-byte
// boundary (TMA boxes must start 16-byte aligned) and ends on a UMMA K-step.
template <bool kFromU8>
__global__ void __launch_bounds__(256) pack_kernel(const double* __restrict__ X, const uint8_t* __restrict__ frames,
                                                   const int32_t* __restrict__ sorted_idx, const double* __restrict__ mean,
                                                   const double* __restrict__ sd, long long Tf, long long k_begin, long long k_end,
                                                   long long K, long long T, long long C1, int planes, int fmt, long long ldc,
                                                   long long ldr, uint16_t* __restrict__ Xa, uint16_t* __restrict__ Xb,
                                                   float* __restrict__ xl, int* __restrict__ overflow) {
  __shared__ uint16_t tile[3][32][33];
  const long long kblocks = (k_end - k_begin + 31) / 32;
  const long long t = blockIdx.x / kblocks, k0 = k_begin + (blockIdx.x % kblocks) * 32;
  const long long c0 = (long long)blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const long long pa = K * T * ldc, pb = C1 * ldr, Kp = (K + 15) / 16 * 16;
  long long f = 0;
  if constexpr (kFromU8) f = sorted_idx[t];
  // single-plane fast path from uint8 frames: the block's 32 columns share (mean, 1/std), taken to fp32 ONCE per block --
  // per-element fp64 subtract + divide is what bounded this kernel on the narrow fp64 pipe (2.7 ms -> HBM-bound)
  __shared__ float s_mean[32], s_istd[32];
  const bool fast = kFromU8 && planes == 1;
  if (fast) {
    if (threadIdx.x < 32) {
      const long long c = c0 + threadIdx.x;
      const long long col = f * C1 + (c < C1 ? c : C1 - 1);
      s_mean[threadIdx.x] = (float)mean[col];
      s_istd[threadIdx.x] = (float)(1.0 / sd[col]);
    }
    __syncthreads();
  }
  for (int rr = ly; rr < 32; rr += 8) {
    const long long k = k0 + rr, c = c0 + lx;
    uint16_t pl[3] = {0, 0, 0};
    if (fast) {
      if (k < k_end && c < C1) {
        const float vf = ((float)frames[(k * Tf + f) * C1 + c] - s_mean[lx]) * s_istd[lx];
        if (fmt == VS_OPERAND_F16 && overflow && !(fabsf(vf) <= 65504.f)) atomicOr(overflow, 1);
        pl[0] = enc16(vf, fmt);
        Xa[(t * K + k) * ldc + c] = pl[0];
      }
    } else if (k < k_end && c < C1) {
      double v;
      if constexpr (kFromU8) {
        const long long col = f * C1 + c;
        v = ((double)frames[(k * Tf + f) * C1 + c] - mean[col]) / sd[col];
      } else {
        v = X[((k - k_begin) * T + t) * (C1 + 1) + c];   // X points at trial k_begin
      }
      if (fmt == VS_OPERAND_F16 && overflow && !(fabs(v) <= 65504.0)) atomicOr(overflow, 1);
      split_planes(v, planes, fmt, pl);
      const long long d = t * K + k;
      for (int p = 0; p < planes; ++p) Xa[p * pa + d * ldc + c] = pl[p];
    }
    for (int p = 0; p < planes; ++p) tile[p][rr][lx] = pl[p];
  }
  __syncthreads();
  for (int cc = ly; cc < 32; cc += 8) {
    const long long c = c0 + cc, k = k0 + lx;
    if (c < C1 && k < k_end)
      for (int p = 0; p < planes; ++p) Xb[p * pb + c * ldr + t * Kp + k] = tile[p][lx][cc];
  }
  if (blockIdx.y == 0 && threadIdx.x < 32) {
    const long long k = k0 + threadIdx.x;
    if (k < k_end) xl[t * K + k] = kFromU8 ? 1.0f : (float)X[((k - k_begin) * T + t) * (C1 + 1) + C1];
  }
}

// Single-plane pack from uint8 frames, wide tile: 32 trials (of ONE time bin) x 128 features per block.  Every thread
// reads 4 consecutive pixels as ONE 32-bit load (128 B per trial row), writes them as 4 packed 16-bit values of Xa
// (256 B per row) and, after a shared-memory transpose, 2 trials of one feature row of Xb per thread (64 B per feature).
// Same arithmetic as the fast path of pack_kernel ((x - mean) * (1/std) in fp32); that kernel moved 32-byte row segments
// with byte loads and ran at 1.5 TB/s.
__global__ void __launch_bounds__(256) pack_u8_wide_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ sorted_idx,
                                                           const double* __restrict__ mean, const double* __restrict__ sd, long long Tf,
                                                           long long K, long long T, long long C1, int fmt, long long ldc, long long ldr,
                                                           uint16_t* __restrict__ Xa, uint16_t* __restrict__ Xb, float* __restrict__ xl,
                                                           int* __restrict__ overflow) {
  __shared__ uint16_t tile[128][34];          // [feature][trial], 2 trials per 32-bit word, odd word pitch
  __shared__ float s_mean[128], s_istd[128];
  const long long kblocks = (K + 31) / 32;
  const long long t = blockIdx.x / kblocks, k0 = (blockIdx.x % kblocks) * 32;
  const long long c0 = (long long)blockIdx.y * 128, Kp = (K + 15) / 16 * 16;
  const long long f = sorted_idx[t];
  if (threadIdx.x < 128) {
    const long long c = c0 + threadIdx.x;
    const long long col = f * C1 + (c < C1 ? c : C1 - 1);
    s_mean[threadIdx.x] = (float)mean[col];
    s_istd[threadIdx.x] = (float)(1.0 / sd[col]);
  }
  __syncthreads();
  const int cx = (threadIdx.x & 31) * 4, ry = threadIdx.x >> 5;      // 4 features per thread, 8 trial rows per pass
  bool ovf = false;
#pragma unroll
  for (int rr = ry; rr < 32; rr += 8) {
    const long long k = k0 + rr, c = c0 + cx;
    uint16_t v[4] = {0, 0, 0, 0};
    if (k < K && c < C1) {
      const uint8_t* src = frames + (k * Tf + f) * C1 + c;
      uint32_t w;
      if (c + 4 <= C1 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
        w = __ldg(reinterpret_cast<const uint32_t*>(src));
      } else {
        w = 0;
        for (int i = 0; i < 4; ++i)
          if (c + i < C1) w |= (uint32_t)src[i] << (8 * i);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float vf = ((float)((w >> (8 * i)) & 0xff) - s_mean[cx + i]) * s_istd[cx + i];
        if (fmt == VS_OPERAND_F16 && !(fabsf(vf) <= 65504.f)) ovf = true;
        v[i] = (c + i < C1) ? enc16(vf, fmt) : (uint16_t)0;
      }
      uint16_t* dst = Xa + (t * K + k) * ldc + c;
      if (c + 4 <= C1) {
        *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));   // ldc % 8 == 0, c % 4 == 0
      } else {
        for (int i = 0; i < 4; ++i)
          if (c + i < C1) dst[i] = v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) tile[cx + i][rr] = v[i];
  }
  if (ovf && overflow) atomicOr(overflow, 1);
  __syncthreads();
  // Xb[c][t*Kp + k]: 16 threads per feature row, 2 trials (one 32-bit store) each; 16 feature rows per pass
  const int kp = (threadIdx.x & 15) * 2, cr = threadIdx.x >> 4;
#pragma unroll
  for (int cc = cr; cc < 128; cc += 16) {
    const long long c = c0 + cc, k = k0 + kp;
    if (c < C1 && k < K) {
      uint16_t* dst = Xb + c * ldr + t * Kp + k;
      if (k + 1 < K) *reinterpret_cast<uint32_t*>(dst) = (uint32_t)tile[cc][kp] | ((uint32_t)tile[cc][kp + 1] << 16);   // Kp, k0, kp even
      else dst[0] = tile[cc][kp];
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < 32) {
    const long long k = k0 + threadIdx.x;
    if (k < K) xl[t * K + k] = 1.0f;
  }
}


// ------------------------------------------------------------------ exact-operand pack (vs_rrr_pack_u8_exact)
// The whole-fit parity of the reference's un-line-searched L-BFGS needs ~22 significant bits on every tensor-core
// operand (profiles/r02_precision_sim_full.txt).  Two facts make that affordable for uint8 frames:
//   * frame - round(mean) is an INTEGER of magnitude <= 255: exact in IEEE half.  The backward operand Xi holds those
//     integers (layout of Xb); the z-score 1/std[t,c] moves into the rank-one weights of the dense backward and the
//     fractional part of the mean into a rank-T correction of its result (epi_b_kernel), so dbeta_t = X_t^T R_t is formed
//     from an exact A operand and only the small operand R needs two planes;
//   * the forward operand Xa stores z = (frame - mean)/std as hi + lo half planes (the second plane is the rounding
//     residual of the first), which the factorised GEMM consumes as three plane products.
// isdT[c*ldt + t] = 1/std[t,c];  qT[c*ldt + t] = (mean - round(mean))[t,c] / std[t,c]   (zeros for T <= t < ldt)
__global__ void __launch_bounds__(256) exact_stats_kernel(const int32_t* __restrict__ sorted_idx, const double* __restrict__ mean,
                                                          const double* __restrict__ sd, long long T, long long C1, long long ldt,
                                                          float* __restrict__ isdT, float* __restrict__ qT) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C1 * ldt) return;
  const long long c = i / ldt, t = i % ldt;
  float a = 0.f, q = 0.f;
  if (t < T) {
    const long long col = (long long)sorted_idx[t] * C1 + c;
    const double m = mean[col], is = 1.0 / sd[col];
    a = (float)is;
    q = (float)((m - rint(m)) * is);
  }
  isdT[i] = a;
  qT[i] = q;
}

// Tables of the dense forward (VS_RRR_MODE_DENSE), one thread per (t, c), rows padded to Tq = T rounded up to 16:
//   isd[t*ldc + c] = 1/std[t,c] (0 past C1 and past T);  qh: hi + lo half planes of q[t,c] = (mean - round(mean))/std (the A
//   operand of the small GEMM that forms the constant term);  isdmax[t] = max_c isd[t,c] (caller zeroes it).
// Features that are constant in the train split (std clipped at 1e-8) get scale 0 in every table: their exact
// operand is identically zero on the train split, and the pack kernel flags any other split on which it is not.
__global__ void __launch_bounds__(256) exact_stats2_kernel(const int32_t* __restrict__ sorted_idx, const double* __restrict__ mean,
                                                           const double* __restrict__ sd, long long T, long long Tq, long long C1,
                                                           long long ldc, float* __restrict__ isd, uint16_t* __restrict__ qh,
                                                           unsigned* __restrict__ isdmax_bits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Tq * ldc) return;
  const long long t = i / ldc, c = i % ldc;
  float a = 0.f, q = 0.f;
  if (t < T && c < C1) {
    const long long col = (long long)sorted_idx[t] * C1 + c;
    const double m = mean[col], sdv = sd[col];
    if (sdv > 1e-8) {
      a = (float)(1.0 / sdv);
      q = (float)((m - rint(m)) / sdv);
    }
  }
  if (t < T) isd[i] = a;
  const __half h = __float2half_rn(q);
  qh[i] = __half_as_ushort(h);
  qh[Tq * ldc + i] = __half_as_ushort(__float2half_rn(q - __half2float(h)));
  if (a > 0.f) atomicMax(isdmax_bits + t, __float_as_uint(a));       // non-negative floats order like their bit patterns
}

// 32 trials (of ONE time bin) x 128 features per block, like pack_u8_wide_kernel.  Xa (may be NULL): two half planes of z
// (row d = t*K + k); Xc (may be NULL): the exact integers in the same row layout (dense forward operand); Xi (may be NULL:
// forward-only splits): the exact integers, transposed through shared memory (backward operand).
__global__ void __launch_bounds__(256) pack_u8_exact_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ sorted_idx,
                                                            const double* __restrict__ mean, const double* __restrict__ sd, long long Tf,
                                                            long long K, long long T, long long C1, long long ldc, long long ldr,
                                                            uint16_t* __restrict__ Xa, uint16_t* __restrict__ Xc, uint16_t* __restrict__ Xi,
                                                            float* __restrict__ xl, int* __restrict__ overflow) {
  __shared__ uint16_t tile[128][34];          // [feature][trial]
  __shared__ float s_m[128], s_dl[128], s_istd[128];
  __shared__ bool s_const[128];
  const long long kblocks = (K + 31) / 32;
  const long long t = blockIdx.x / kblocks, k0 = (blockIdx.x % kblocks) * 32;
  const long long c0 = (long long)blockIdx.y * 128, Kp = (K + 15) / 16 * 16;
  const long long f = sorted_idx[t];
  const long long pa = K * T * ldc;
  if (threadIdx.x < 128) {
    const long long c = c0 + threadIdx.x;
    const long long col = f * C1 + (c < C1 ? c : C1 - 1);
    const double m = mean[col], mi = rint(m);
    s_m[threadIdx.x] = (float)mi;
    s_dl[threadIdx.x] = (float)(m - mi);
    s_istd[threadIdx.x] = (float)(1.0 / sd[col]);
    s_const[threadIdx.x] = !(sd[col] > 1e-8);
  }
  __syncthreads();
  const int cx = (threadIdx.x & 31) * 4, ry = threadIdx.x >> 5;
  bool ovf = false;
#pragma unroll
  for (int rr = ry; rr < 32; rr += 8) {
    const long long k = k0 + rr, c = c0 + cx;
    uint16_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0}, xi[4] = {0, 0, 0, 0};
    if (k < K && c < C1) {
      const uint8_t* src = frames + (k * Tf + f) * C1 + c;
      uint32_t w;
      if (c + 4 <= C1 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
        w = __ldg(reinterpret_cast<const uint32_t*>(src));
      } else {
        w = 0;
        for (int i = 0; i < 4; ++i)
          if (c + i < C1) w |= (uint32_t)src[i] << (8 * i);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (c + i >= C1) continue;
        const float xc = (float)((w >> (8 * i)) & 0xff) - s_m[cx + i];          // exact integer, |xc| <= 255
        const float zf = (xc - s_dl[cx + i]) * s_istd[cx + i];
        // a value outside the half range -- or, for the integer operands, a feature that is constant in the train split but
        // not here (the reference then multiplies by 1e8) -- is reported, never silently mangled
        if (Xa ? !(fabsf(zf) <= 65504.f) : (s_const[cx + i] && xc != 0.f)) ovf = true;
        const __half h = __float2half_rn(zf);
        hi[i] = __half_as_ushort(h);
        lo[i] = __half_as_ushort(__float2half_rn(zf - __half2float(h)));
        xi[i] = __half_as_ushort(__float2half_rn(xc));
      }
      if (Xa) {
        uint16_t* dst = Xa + (t * K + k) * ldc + c;
        if (c + 4 <= C1) {
          *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16));
          *reinterpret_cast<uint2*>(dst + pa) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
        } else {
          for (int i = 0; i < 4; ++i)
            if (c + i < C1) { dst[i] = hi[i]; dst[pa + i] = lo[i]; }
        }
      }
      if (Xc) {
        uint16_t* dst = Xc + (t * K + k) * ldc + c;
        if (c + 4 <= C1) {
          *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)xi[0] | ((uint32_t)xi[1] << 16), (uint32_t)xi[2] | ((uint32_t)xi[3] << 16));
        } else {
          for (int i = 0; i < 4; ++i)
            if (c + i < C1) dst[i] = xi[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) tile[cx + i][rr] = xi[i];
  }
  if (ovf && overflow) atomicOr(overflow, 1);
  __syncthreads();
  if (Xi) {
    const int kp = (threadIdx.x & 15) * 2, cr = threadIdx.x >> 4;
#pragma unroll
    for (int cc = cr; cc < 128; cc += 16) {
      const long long c = c0 + cc, k = k0 + kp;
      if (c < C1 && k < K) {
        uint16_t* dst = Xi + c * ldr + t * Kp + k;
        if (k + 1 < K) *reinterpret_cast<uint32_t*>(dst) = (uint32_t)tile[cc][kp] | ((uint32_t)tile[cc][kp + 1] << 16);
        else dst[0] = tile[cc][kp];
      }
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < 32) {
    const long long k = k0 + threadIdx.x;
    if (k < K) xl[t * K + k] = 1.0f;
  }
}

// ------------------------------------------------------------------ fused R0 loader: statistics + pack, ONE read of the frames
// Block (t, 128-feature chunk): the K trials of one selected frame's chunk (K x 128 bytes, 51 KB at K = 400) are staged in
// shared memory with 128-bit global loads; the exact integer column sums give mean / clipped std (src/utils/utils.py:107-112,
// kStats: train split) or the caller's statistics are read (other splits); then every operand of the split is produced from
// the staged bytes with 128-bit stores:
//   Xa : z as hi + lo half planes          (EXACT mode)     row t*K + k, 8 features per store
//   Xc : the exact integers, same layout   (DENSE mode)
//   Xi : the exact integers transposed     (train split)    row c, 8 trials per store (a feature row receives its whole
//                                                           bin -- K contiguous values -- from one block), pad trials zeroed
// The old path read the frames twice (colstats_kernel at 0.26 of the HBM peak, then the pack at 0.49) and wrote the
// transposed operand in 64-byte pieces.  Requires C1 % 4 == 0 (row starts are then 4-byte aligned: the 16-byte global words
// are shifted into place by whole 32-bit lanes) and K * 128 bytes of shared memory; other shapes use the two-kernel path.
constexpr int kPackRow = 136;
template <bool kStats>
__global__ void __launch_bounds__(256, 3) pack_fused_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ sorted_idx,
                                                         double* __restrict__ mean, double* __restrict__ sd, long long Tf, long long K,
                                                         long long T, long long C1, long long ldc, long long ldr, uint16_t* __restrict__ Xa,
                                                         uint16_t* __restrict__ Xc, uint16_t* __restrict__ Xi, float* __restrict__ xl,
                                                         int* __restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t raw[];            // [K][kPackRow]: 128 staged bytes per trial, rows 136 bytes apart
  //                                                           (two banks of skew per row: the 16-byte staging stores of a warp span 3-4 rows)
  __shared__ float s_m[128], s_dl[128], s_istd[128];
  __shared__ unsigned s_s1[8][128], s_s2[8][128];
  __shared__ unsigned s_vmin[8][32], s_vmax[8][32];
  // chunk index fastest: consecutive blocks read neighbouring 128-byte pieces of the same frame rows, so the 128-byte DRAM
  // lines a piece straddles (row starts are only 4-byte aligned) are shared through L2 instead of being fetched twice
  const long long t = blockIdx.y, c0 = (long long)blockIdx.x * 128, Kp = (K + 15) / 16 * 16;
  const long long f = sorted_idx[t];
  const long long pa = K * T * ldc;
  // ---- phase 1: stage the chunk.  9 aligned 16-byte words cover 128 bytes at any 4-byte shift; 28 rows per pass.
  // The loads of kStageU passes are issued back to back before the first one is stored: one load per pass left the DRAM
  // latency of every pass exposed (43 % of the kernel's stall samples sat on the first shared-memory store of the loop).
  {
    constexpr int kStageU = 8;
    const int wq = threadIdx.x % 9, rr = threadIdx.x / 9;
    if (rr < 28) {
      for (long long kbase = rr; kbase < K; kbase += 28 * kStageU) {
        uint32_t v[kStageU][4];
        int offs[kStageU];
#pragma unroll
        for (int u = 0; u < kStageU; ++u) {
          const long long k = kbase + 28 * u;
          offs[u] = -64;                                         // "no word": nothing is stored
          v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0u;
          if (k < K) {
            const uint8_t* rowp = frames + (k * Tf + f) * C1 + c0;
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(rowp) & ~(uintptr_t)15;
            const uint8_t* wp = reinterpret_cast<const uint8_t*>(a0) + 16 * wq;
            const long long off = wp - rowp;                     // position of this word's first byte inside the chunk: multiple of 4
            if (off > -16 && off < 128) {
              offs[u] = (int)off;
              // the word may reach before the chunk / past the end of the row (never past the allocation for interior rows;
              // the last row of the buffer is guarded by reading only words that start inside the row)
              const bool safe = (c0 + off >= 0) && (c0 + off + 16 <= C1);
              if (safe) {
                const uint4 q = ld_stream_u4(reinterpret_cast<const uint4*>(wp));
                v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const long long cc = c0 + off + 4 * i;
                  if (cc >= 0 && cc + 4 <= C1) v[u][i] = __ldg(reinterpret_cast<const uint32_t*>(wp + 4 * i));
                }
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kStageU; ++u) {
          const long long k = kbase + 28 * u;
          uint32_t* dst = reinterpret_cast<uint32_t*>(raw + k * kPackRow);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int p = offs[u] + 4 * i;
            if (p >= 0 && p < 128) dst[p >> 2] = v[u][i];
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- phase 1b: statistics of the 128 columns (integer sums are exact).
  // The column extremes decide the overflow flag here, once per column, instead of per element in the store loops: z is a
  // non-decreasing function of the byte value, so its largest magnitude is at the smallest or the largest byte.
  {
    // thread = (4 adjacent columns, one of 8 trial phases): one 32-bit shared-memory word per trial; 32-bit sums are exact
    // for K/8 * 65025 < 2^32; the byte-wise extremes come from the SIMD-in-word min / max
    const int q4 = threadIdx.x & 31, grp = threadIdx.x >> 5;
    // first moments in two packed accumulators of 16-bit lanes (columns 0|2 and 1|3: at most K/8 * 255 < 65536 per lane,
    // K <= 1505 by the shared-memory bound); second moments one dp4a per column (the other three bytes masked off)
    unsigned p02 = 0u, p13 = 0u, a2[4] = {0u, 0u, 0u, 0u};
    unsigned wmin = 0xffffffffu, wmax = 0u;
    const int Ks = (int)K;
    for (int k = grp; k < Ks; k += 8) {
      const unsigned wv = *reinterpret_cast<const unsigned*>(raw + k * kPackRow + 4 * q4);
      if constexpr (kStats) {
        p02 += wv & 0x00ff00ffu;
        p13 += (wv >> 8) & 0x00ff00ffu;
        a2[0] = __dp4a(wv & 0x000000ffu, wv, a2[0]);
        a2[1] = __dp4a(wv & 0x0000ff00u, wv, a2[1]);
        a2[2] = __dp4a(wv & 0x00ff0000u, wv, a2[2]);
        a2[3] = __dp4a(wv & 0xff000000u, wv, a2[3]);
      }
      wmin = __vminu4(wmin, wv); wmax = __vmaxu4(wmax, wv);
    }
    const unsigned a1[4] = {p02 & 0xffffu, p13 & 0xffffu, p02 >> 16, p13 >> 16};
#pragma unroll
    for (int i = 0; i < 4; ++i) { s_s1[grp][4 * q4 + i] = a1[i]; s_s2[grp][4 * q4 + i] = a2[i]; }
    s_vmin[grp][q4] = wmin; s_vmax[grp][q4] = wmax;
    __syncthreads();
    const int c = threadIdx.x & 127;
    if (threadIdx.x < 128) {
      const long long cg = c0 + c;
      double m = 0.0, sdev = 1.0;
      if (cg < C1) {
        const long long col = f * C1 + cg;
        if constexpr (kStats) {
          unsigned long long t1 = 0, t2 = 0;
#pragma unroll
          for (int g = 0; g < 8; ++g) { t1 += s_s1[g][c]; t2 += s_s2[g][c]; }
          const double s1 = (double)t1, s2 = (double)t2;
          m = s1 / (double)K;
          const double var = ((double)K * s2 - s1 * s1) / ((double)K * (double)K);     // exact integer moments
          sdev = sqrt(var > 0.0 ? var : 0.0);
          sdev = sdev < 1e-8 ? 1e-8 : sdev;
          mean[col] = m; sd[col] = sdev;
        } else {
          m = mean[col]; sdev = sd[col];
        }
      }
      const double mi = rint(m);
      const float fm = (float)mi, fdl = (float)(m - mi), fis = (float)(1.0 / sdev);
      const bool cst = !(sdev > 1e-8);
      s_m[c] = fm; s_dl[c] = fdl; s_istd[c] = fis;
      if (cg < C1) {
        unsigned bmin = 255u, bmax = 0u;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          bmin = min(bmin, (s_vmin[g][c >> 2] >> (8 * (c & 3))) & 0xffu);
          bmax = max(bmax, (s_vmax[g][c >> 2] >> (8 * (c & 3))) & 0xffu);
        }
        const float x0 = (float)bmin - fm, x1 = (float)bmax - fm;
        const float z0 = (x0 - fdl) * fis, z1 = (x1 - fdl) * fis;
        const bool zbad = !(fabsf(z0) <= 65504.f) || !(fabsf(z1) <= 65504.f);      // a half plane of z would overflow
        const bool cbad = cst && (x0 != 0.f || x1 != 0.f);                          // a constant train column that varies here
        if (overflow && ((Xa && zbad) || (!Xa && Xc && cbad) || (Xi && cbad))) atomicOr(overflow, 1);
      }
    }
    __syncthreads();
  }
  // ---- phase 2a: row-major operands, 8 features (one 128-bit store) per thread: 16 threads per trial row.
  // byte -> float without the conversion pipe: the byte is placed in the mantissa of 2^23 and 2^23 + round(mean) is subtracted
  // (both exact), which is the integer operand; z = (that - frac(mean)) / std in the same two roundings as the reference pack.
  if (Xa || Xc) {
    const int q8 = (threadIdx.x & 15) * 8, r16 = threadIdx.x >> 4;
    const long long c = c0 + q8;
    if (c < C1) {
      float bm[8], dl[8], is[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // columns past C1 (last chunk): staged bytes 0, mean 0, std 1 -> zeros are written
        bm[j] = 8388608.f + s_m[q8 + j]; dl[j] = s_dl[q8 + j]; is[j] = s_istd[q8 + j];
      }
      const int Ki = (int)K;
      for (int k = r16; k < Ki; k += 16) {
        const uint2 b8 = *reinterpret_cast<const uint2*>(raw + k * kPackRow + q8);
        uint32_t hi[4], lo[4], xi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float xc[2], zf[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 2 * i + e;
            const unsigned wv = j < 4 ? b8.x : b8.y;
            const float raw_f = __uint_as_float(__byte_perm(wv, 0x4b000000u, 0x7540 + (j & 3)));   // 2^23 + byte
            xc[e] = raw_f - bm[j];
            zf[e] = (xc[e] - dl[j]) * is[j];
          }
          const __half2 h = __floats2half2_rn(zf[0], zf[1]);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(zf[0] - hf.x, zf[1] - hf.y);
          const __half2 x = __floats2half2_rn(xc[0], xc[1]);
          hi[i] = *reinterpret_cast<const uint32_t*>(&h); lo[i] = *reinterpret_cast<const uint32_t*>(&l); xi[i] = *reinterpret_cast<const uint32_t*>(&x);
        }
        const long long o = (t * K + k) * ldc + c;             // ldc % 64 == 0, c % 8 == 0: 16-byte aligned; the pad columns c >= C1 stay unread
        if (Xa) {
          *reinterpret_cast<uint4*>(Xa + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(Xa + pa + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (Xc) *reinterpret_cast<uint4*>(Xc + o) = make_uint4(xi[0], xi[1], xi[2], xi[3]);
      }
    }
  }
  // ---- phase 2b: transposed integer operand.  Thread = (4 features, 8 trials): 8 conflict-free 32-bit shared-memory reads, four
  // 128-bit stores (one per feature row; the neighbouring 8-trial groups of a row come from the same block, so L2 merges
  // the half sectors before they reach DRAM); pad trials K <= k < Kp are written as zeros
  if (Xi) {
    // thread = (4 features, 16 trials): every store pair covers ONE whole 32-byte sector of a feature row (8-trial pieces
    // left half sectors for L2 to merge; the ones it evicted first came back as DRAM read-modify-writes)
    const int ngrp = (int)(Kp / 16), Ki = (int)K;
    for (int e = threadIdx.x; e < 32 * ngrp; e += 256) {
      const int c4 = (e & 31) * 4;
      const int k16 = (e >> 5) * 16;
      const int ncol = (int)((C1 - c0 - c4) < 4 ? (C1 - c0 - c4) : 4);      // valid features of this thread's four
      if (ncol <= 0) continue;
      uint16_t* dst = Xi + (c0 + c4) * ldr + t * Kp + k16;
      uint32_t wd[16];
      const bool full = k16 + 16 <= Ki;
      if (full) {
#pragma unroll
        for (int j = 0; j < 16; ++j) wd[j] = *reinterpret_cast<const uint32_t*>(raw + (k16 + j) * kPackRow + c4);
      } else {      // the group holding the last trials and the zero pad K <= k < Kp
#pragma unroll
        for (int j = 0; j < 16; ++j) wd[j] = (k16 + j < Ki) ? *reinterpret_cast<const uint32_t*>(raw + (k16 + j) * kPackRow + c4) : 0u;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= ncol) break;
        const float bm = 8388608.f + s_m[c4 + i];
        uint32_t w[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float a0 = __uint_as_float(__byte_perm(wd[2 * jj], 0x4b000000u, 0x7540 + i)) - bm;
          float a1 = __uint_as_float(__byte_perm(wd[2 * jj + 1], 0x4b000000u, 0x7540 + i)) - bm;
          if (!full) {
            if (k16 + 2 * jj >= Ki) a0 = 0.f;
            if (k16 + 2 * jj + 1 >= Ki) a1 = 0.f;
          }
          const __half2 x = __floats2half2_rn(a0, a1);
          w[jj] = *reinterpret_cast<const uint32_t*>(&x);
        }
        uint4* o = reinterpret_cast<uint4*>(dst + (long long)i * ldr);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
  if (blockIdx.x == 0)
    for (long long k = threadIdx.x; k < K; k += 256) xl[t * K + k] = 1.0f;
}

// zeros in the pad trials K <= k < Kp of every time bin of Xb (one thread per (plane, c, t))
__global__ void __launch_bounds__(256) pad_zero_kernel(uint16_t* __restrict__ Xb, long long rows, long long T, long long K, long long Kp,
                                                       long long ldr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * T) return;
  const long long row = i / T, t = i % T;
  for (long long k = K; k < Kp; ++k) Xb[row * ldr + t * Kp + k] = 0;
}

// column mean / population std (clipped at 1e-8) over the K trials: src/utils/utils.py:107-112
__global__ void colstats_kernel(const uint8_t* __restrict__ frames, long long K, long long cols, double* __restrict__ mean,
                                double* __restrict__ sd) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  // integer sums are exact: sum x <= 255 K, sum x^2 <= 65025 K
  unsigned long long s1 = 0, s2 = 0;
  for (long long k = 0; k < K; ++k) {
    const unsigned v = frames[k * cols + c];
    s1 += v;
    s2 += v * v;
  }
  const double m = (double)s1 / (double)K;
  // var = E[(x-m)^2] evaluated from exact integer moments: (K*s2 - s1^2) / K^2
  const double var = ((double)K * (double)s2 - (double)s1 * (double)s1) / ((double)K * (double)K);
  double s = sqrt(var > 0.0 ? var : 0.0);
  mean[c] = m;
  sd[c] = s < 1e-8 ? 1e-8 : s;
}


// ------------------------------------------------------------------ y preprocessing (R0)
// scipy.ndimage.gaussian_filter1d(sigma, axis=1, mode="reflect", truncate=4.0): radius = int(4 sigma + .5),
// weights exp(-x^2 / (2 sigma^2)) normalised to sum 1, "reflect" = (d c b a | a b c d | d c b a).
__global__ void smooth_y_kernel(const float* __restrict__ counts, long long K, long long T, long long N, double sigma,
                                const double* __restrict__ mean, const double* __restrict__ sd, float* __restrict__ out,
                                float* __restrict__ out_lo) {
  __shared__ double wts[65];
  const int radius = (int)(4.0 * sigma + 0.5);
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = -radius; i <= radius; ++i) { wts[i + radius] = exp(-0.5 * (double)i * (double)i / (sigma * sigma)); s += wts[i + radius]; }
    for (int i = 0; i <= 2 * radius; ++i) wts[i] /= s;
  }
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * T * N) return;
  const long long n = idx % N, t = (idx / N) % T, k = idx / (N * T);
  double acc = 0.0;
  for (int i = -radius; i <= radius; ++i) {
    long long tt = t + i;
    // reflect about the edges (period 2T)
    while (tt < 0 || tt >= T) tt = tt < 0 ? -tt - 1 : 2 * T - 1 - tt;
    acc += wts[i + radius] * (double)counts[(k * T + tt) * N + n];
  }
  if (mean) acc = (acc - mean[t * N + n]) / sd[t * N + n];
  const float hi = (float)acc;
  out[idx] = hi;
  if (out_lo) out_lo[idx] = (float)(acc - (double)hi);     // out + out_lo = the float64 value to ~2^-48
}

__global__ void colstats_f32_kernel(const float* __restrict__ x, long long K, long long cols, double* __restrict__ mean,
                                    double* __restrict__ sd) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (long long k = 0; k < K; ++k) s += (double)x[k * cols + c];
  const double m = s / (double)K;
  double v = 0.0;
  for (long long k = 0; k < K; ++k) { const double dlt = (double)x[k * cols + c] - m; v += dlt * dlt; }
  double sdev = sqrt(v / (double)K);
  mean[c] = m;
  sd[c] = sdev < 1e-8 ? 1e-8 : sdev;
}

// ------------------------------------------------------------------ closure stage 0: U -> Ub planes, Gram partials
// Ub[p][j*Npad + n][c] = plane p of U[n][c][j];  Gp[block][i*r+j] = sum over the block's (n,c) of U_i U_j.
// One thread owns kPrepC consecutive features of one neuron: its r*kPrepC doubles are contiguous in U.
constexpr int kPrepC = 4;
template <int RMAX>
__global__ void __launch_bounds__(256) prep_u_kernel(const double* __restrict__ U, long long N, long long Npad, long long C1,
                                                     int r, int planes, int fmt, long long ldc, uint16_t* __restrict__ Ub,
                                                     double* __restrict__ Gp, double uscale, float* __restrict__ U32, long long ldu,
                                                     unsigned* __restrict__ umax_bits) {
  __shared__ double red[8][RMAX * RMAX];
  const long long n = blockIdx.y;
  const long long c0 = ((long long)blockIdx.x * 256 + threadIdx.x) * kPrepC;
  const long long pu = (long long)r * Npad * ldc;
  double g[RMAX * RMAX];
#pragma unroll
  for (int e = 0; e < RMAX * RMAX; ++e) g[e] = 0.0;
  float um = 0.f;
  // fast path (rank 3, two half planes, a full group of 4 features, even C1): the 12 doubles of the group arrive as six
  // 128-bit loads and every (plane, component) row receives its 4 features as ONE 8-byte store
  bool done = false;
  if constexpr (RMAX == 3) {
    if (planes == 2 && fmt == VS_OPERAND_F16 && c0 + kPrepC <= C1 && (C1 & 1) == 0 && (reinterpret_cast<uintptr_t>(U) & 15) == 0) {
      const double2* up = reinterpret_cast<const double2*>(U + (n * C1 + c0) * 3);
      double u[12];
#pragma unroll
      for (int i = 0; i < 6; ++i) { const double2 v2 = __ldg(up + i); u[2 * i] = v2.x; u[2 * i + 1] = v2.y; }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        uint16_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double v = u[3 * q + j] * uscale;
          const __half h = __float2half_rn((float)v);
          hi[q] = __half_as_ushort(h);
          lo[q] = __half_as_ushort(__float2half_rn((float)(v - (double)__half2float(h))));
        }
        const long long o = ((long long)j * Npad + n) * ldc + c0;
        *reinterpret_cast<uint2*>(Ub + o) = make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16));
        *reinterpret_cast<uint2*>(Ub + pu + o) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
      }
      if (U32) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 3; ++j) { const float uf = (float)u[3 * q + j]; U32[u32g_index(n, c0 + q, j, N)] = uf; um = fmaxf(um, fabsf(uf)); }
      }
      if (Gp) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) g[i * 3 + j] = fma(u[3 * q + i], u[3 * q + j], g[i * 3 + j]);
      }
      done = true;
    }
  }
#pragma unroll
  for (int q = 0; q < kPrepC; ++q) {
    if (done) break;
    const long long c = c0 + q;
    if (c >= C1) {
      // dense forward operand (u32g_index layout, rank 3) with zeros in the pad columns: its generator reads whole 64-feature blocks
      if (U32 && c < ldu)
        for (int j = 0; j < r; ++j) U32[u32g_index(n, c, j, N)] = 0.f;
      continue;
    }
    double u[RMAX];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) u[j] = j < r ? U[(n * C1 + c) * r + j] : 0.0;
    if (U32) {
#pragma unroll
      for (int j = 0; j < RMAX; ++j)
        if (j < r) { const float uf = (float)u[j]; U32[u32g_index(n, c, j, N)] = uf; um = fmaxf(um, fabsf(uf)); }
    }
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      if (j >= r) break;
      uint16_t pl[3];
      split_planes(u[j] * uscale, planes, fmt, pl);        // uscale: a power of two (exact), undone by the epilogue of GEMM-F
      for (int p = 0; p < planes; ++p) Ub[p * pu + ((long long)j * Npad + n) * ldc + c] = pl[p];
    }
    if (Gp) {
#pragma unroll
      for (int i = 0; i < RMAX; ++i)
#pragma unroll
        for (int j = i; j < RMAX; ++j)
          if (j < r) g[i * RMAX + j] = fma(u[i], u[j], g[i * RMAX + j]);
    }
  }
  if (umax_bits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) um = fmaxf(um, __shfl_xor_sync(0xffffffffu, um, o));
    // non-negative floats order like their bit patterns: an integer max is exact and order-independent
    if ((threadIdx.x & 31) == 0 && um > 0.f) atomicMax(umax_bits, __float_as_uint(um));
  }
  if (Gp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < RMAX; ++i)
#pragma unroll
      for (int j = i; j < RMAX; ++j) {
        if (j >= r) continue;
        const double sum = warp_sum(g[i * RMAX + j]);
        if (lane == 0) { red[warp][i * r + j] = sum; red[warp][j * r + i] = sum; }
      }
    __syncthreads();
    if (threadIdx.x < r * r) {
      double sum = 0.0;
      for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
      Gp[((long long)blockIdx.y * gridDim.x + blockIdx.x) * (r * r) + threadIdx.x] = sum;
    }
  }
}

// G = sum of Gram partials (block e < r*r: thread i sums partials i, i+256, ... then a fixed tree, so the result does
// not depend on scheduling); the last block forms W = V V^T.
__global__ void __launch_bounds__(256) small_mats_kernel(const double* __restrict__ Gp, long long nblocks, const double* __restrict__ V,
                                                         int r, long long T, double* __restrict__ G, double* __restrict__ W) {
  __shared__ double red[256];
  const int rr = r * r;
  const int e = blockIdx.x;
  if (e < rr) {
    double s = 0.0;
    if (Gp)
      for (long long b = threadIdx.x; b < nblocks; b += 256) s += Gp[b * rr + e];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
      if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
      __syncthreads();
    }
    if (threadIdx.x == 0) G[e] = red[0];
  } else if (threadIdx.x < rr) {
    const int i = threadIdx.x / r, j = threadIdx.x % r;
    double w = 0.0;
    for (long long t = 0; t < T; ++t) w += V[i * T + t] * V[j * T + t];
    W[threadIdx.x] = w;
  }
}

// ------------------------------------------------------------------ closure stage 2: epilogue of GEMM-F
constexpr int kEpiRows = 64;

// block (kb, t, nt) = 64 trials of time bin t x 32 neurons.  Phase 1 (lanes over n): yhat, residual, per-block
// partials of SSE, db and dV.  Phase 2 (lanes over trial pairs): the B operand of the backward GEMM -- R (x) V planes
// for the factorised route, the plain residual R for the dense per-time-bin route (one plane, or hi + lo half planes in
// the exact-operand mode) -- written as 16-bit pairs.  The residual itself never goes to HBM.
// AT = arithmetic of the epilogue: float on the plain 16-bit path (the operands carry ~1e-3 of rounding anyway), double in
// the exact-operand mode (Z comes out of TMEM as fp32; V, b and the targets y = y + y_lo enter at full precision, and
// the SSE / db / dV partial sums are wide: dV is a small difference of large sums, DESIGN.md "RRR precision").
template <bool kPredict, int RMAX, typename AT>
__global__ void __launch_bounds__(256) epi_f_kernel(const float* __restrict__ Z, long long ldz, int splits, long long split_stride,
                                                    const float* __restrict__ y, const float* __restrict__ y_lo, const float* __restrict__ xl,
                                                    const double* __restrict__ V, const double* __restrict__ b, long long K,
                                                    long long T, long long N, long long Npad, int r, int planes, int fmt, long long ldr,
                                                    uint16_t* __restrict__ RV, AT* __restrict__ sse_part,
                                                    AT* __restrict__ db_part, AT* __restrict__ pv_part,
                                                    double* __restrict__ yhat, long long Kp, int dense, double zscale) {
  __shared__ AT Rs[kEpiRows][33];
  __shared__ AT red[8][32][2];
  __shared__ float xls[kEpiRows];
  __shared__ AT pvs[8][RMAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long kb = blockIdx.x, t = blockIdx.y, nt = blockIdx.z, KB = gridDim.x, NT = gridDim.z;
  const long long k0 = kb * kEpiRows, d0 = t * K + k0, n0 = nt * 32;
  AT vt[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) vt[j] = j < r ? (AT)(V[(long long)j * T + t] * zscale) : (AT)0;     // Z = zscale^-1 * X U (see prep_u_kernel)
  if (threadIdx.x < kEpiRows) xls[threadIdx.x] = (k0 + threadIdx.x < K) ? xl[d0 + threadIdx.x] : 0.f;
  __syncthreads();
  AT pvacc[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) pvacc[j] = (AT)0;
  const long long n = n0 + lane;
  const AT bn = n < N ? (AT)b[n * T + t] : (AT)0;
  AT sse = (AT)0, sdb = (AT)0;
  constexpr int RPW = kEpiRows / 8;   // rows per warp
  // all loads of the block's tile are issued before the first use
  AT z[RPW][RMAX];
  AT yv[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int rl = w * RPW + i;
    const long long k = k0 + rl, d = d0 + rl;
    const bool ok = k < K && n < N;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      z[i][j] = (AT)0;
      if (ok && j < r) {
        const float* zp = Z + d * ldz + (long long)j * Npad + n;
        if (splits == 1) {
          z[i][j] = (AT)__ldg(zp);
        } else {  // split-K partials (high-precision mode): ordered sum, wide accumulator; the loads are independent
          double zs = 0.0;
#pragma unroll 4
          for (int sp = 0; sp < splits; ++sp) zs += (double)__ldg(zp + (long long)sp * split_stride);
          z[i][j] = (AT)zs;
        }
      }
    }
    if constexpr (!kPredict) {
      yv[i] = (AT)0;
      if (ok) {
        yv[i] = (AT)__ldg(y + (k * T + t) * N + n);
        if (y_lo) yv[i] += (AT)__ldg(y_lo + (k * T + t) * N + n);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int rl = w * RPW + i;
    const long long k = k0 + rl;
    AT res = (AT)0;
    if (k < K && n < N) {
      AT acc = (AT)xls[rl] * bn;
#pragma unroll
      for (int j = 0; j < RMAX; ++j)
        if (j < r) acc = fma(vt[j], z[i][j], acc);
      if constexpr (kPredict) {
        yhat[(k * T + t) * N + n] = (double)acc;
      } else {
        res = acc - yv[i];
        sse = fma(res, res, sse);
        sdb = fma((AT)xls[rl], res, sdb);
#pragma unroll
        for (int j = 0; j < RMAX; ++j)
          if (j < r) pvacc[j] = fma(res, z[i][j], pvacc[j]);
      }
    }
    if constexpr (!kPredict) Rs[rl][lane] = res;
  }
  if constexpr (!kPredict) {
    red[w][lane][0] = sse;
    red[w][lane][1] = sdb;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      const AT sj = warp_sum(pvacc[j]);
      if (lane == 0) pvs[w][j] = sj;
    }
    __syncthreads();
    // phase 2: lane = trial pair (2*lane, 2*lane+1), warp w covers neurons w*4 .. w*4+3 of the tile.  Columns follow Xb:
    // t*Kp + k with zeros in the pad K <= k < Kp (Rs is zero there); Kp and k0 are even, so pair stores are aligned.
    const long long ka = k0 + 2 * lane;
    const long long dcol = t * Kp + ka;
    if (dense) {
      // dense backward (tc::rrr_bwd_dense): the plain residual R[n][t*Kp + k], `planes` residual planes of Npad rows
      const long long prd = Npad * ldr;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nl = w * 4 + i;
        const long long nn = n0 + nl;
        if (nn >= N || ka >= Kp) continue;
        float ra = (float)Rs[2 * lane][nl], rb = (float)Rs[2 * lane + 1][nl];
        for (int pl = 0; pl < planes && pl < 2; ++pl) {
          const uint32_t lo = enc16(ra, fmt), hi = enc16(rb, fmt);
          *reinterpret_cast<uint32_t*>(RV + pl * prd + nn * ldr + dcol) = lo | (hi << 16);
          ra -= dec16((uint16_t)lo, fmt); rb -= dec16((uint16_t)hi, fmt);
        }
      }
    } else {
    const long long prv = (long long)r * Npad * ldr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nl = w * 4 + i;
      const long long nn = n0 + nl;
      if (nn >= N || ka >= Kp) continue;
      const float r0v = (float)Rs[2 * lane][nl], r1v = (float)Rs[2 * lane + 1][nl];
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        if (j >= r) continue;
        uint16_t pl0[3], pl1[3];
        if (planes == 1) {   // plain bf16 operand: no fp64 on this path (the vector fp64 pipe is narrow)
          pl0[0] = enc16((float)vt[j] * r0v, fmt); pl1[0] = enc16((float)vt[j] * r1v, fmt);
          pl0[1] = pl0[2] = pl1[1] = pl1[2] = 0;
        } else {
          split_planes((double)((float)vt[j] * r0v), planes, fmt, pl0);
          split_planes((double)((float)vt[j] * r1v), planes, fmt, pl1);
        }
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
          if (pq >= planes) continue;
          *reinterpret_cast<uint32_t*>(RV + pq * prv + ((long long)j * Npad + nn) * ldr + dcol) = (uint32_t)pl0[pq] | ((uint32_t)pl1[pq] << 16);
        }
      }
    }
    }
    // phase 3: per-(t, kb, n) partials, the 8 warps' sums added in a fixed order
    if (w == 0 && n < N) {
      AT s0 = (AT)0, s1 = (AT)0;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) { s0 += red[w2][lane][0]; s1 += red[w2][lane][1]; }
      sse_part[(t * KB + kb) * N + n] = s0;
      db_part[(t * KB + kb) * N + n] = s1;
    }
    if (threadIdx.x < r) {
      AT sj = (AT)0;
      for (int w2 = 0; w2 < 8; ++w2) sj += pvs[w2][threadIdx.x];
      pv_part[((t * KB + kb) * NT + nt) * r + threadIdx.x] = sj * (AT)zscale;
    }
  }
}

// The same epilogue for the exact-operand training closure (rank 3, N % 4 == 0, dense backward, two IEEE-half residual
// planes), laid out for memory-level parallelism: epi_f_kernel<.., double> above keeps 128 registers per thread and reads Z in
// 128-byte pieces -- 0.31 ms at 19 % of the HBM peak for 0.46 GB (profiles/r02_ncu_full_exact.csv).  Here a block owns 64
// trials of one time bin and ALL neurons (32 at a time in shared memory); a thread item is (trial, 4 neurons): 3 * splits independent 128-bit loads of the
// Z partials (one 576-byte run per trial and j at N = 144) summed in float64, then yhat, the residual, the dV partial.
// The residual tile goes through shared memory (float64) for the per-neuron SSE / db sums (ordered over the trials) and
// the transposed 128-byte stores of the two residual planes.
constexpr int kFxRows = 64;
constexpr int kFxHalf = 32;                             // trials staged in shared memory at a time
__global__ void __launch_bounds__(256) epi_fx_kernel(const float* __restrict__ Z, long long ldz, int splits, long long split_stride,
                                                     const float* __restrict__ y, const float* __restrict__ y_lo, const float* __restrict__ xl,
                                                     const double* __restrict__ V, const double* __restrict__ b, long long K, long long T,
                                                     long long N, long long Npad, long long ldr, uint16_t* __restrict__ RV,
                                                     double* __restrict__ sse_part, double* __restrict__ db_part, double* __restrict__ pv_part,
                                                     long long Kp, double zscale) {
  extern __shared__ double fx_rs[];                     // [kFxHalf][N + 1]
  __shared__ double bs[160];
  __shared__ double pvs[8][3];
  __shared__ float xls[kFxRows];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long kb = blockIdx.x, t = blockIdx.y, KB = gridDim.x;
  const long long k0 = kb * kFxRows, d0 = t * K + k0;
  const int NQ = (int)(N >> 2), NS = (int)N + 1;
  const double v0 = V[t] * zscale, v1 = V[T + t] * zscale, v2 = V[2 * T + t] * zscale;      // Z = zscale^-1 * X U (prep_u_kernel)
  if (threadIdx.x < kFxRows) xls[threadIdx.x] = (k0 + threadIdx.x < K) ? xl[d0 + threadIdx.x] : 0.f;
  for (int n = threadIdx.x; n < N; n += 256) bs[n] = b[(long long)n * T + t];
  double pv0 = 0.0, pv1 = 0.0, pv2 = 0.0;
  double cs0 = 0.0, cs1 = 0.0;                          // thread n < N: SSE and sum xl * R of neuron n over the block's trials
  const long long prd = Npad * ldr;
  for (int half = 0; half < kFxRows / kFxHalf; ++half) {
    __syncthreads();                                    // xls / bs ready (first pass); the previous half's tile fully consumed
    for (int item = threadIdx.x; item < kFxHalf * NQ; item += 256) {
      const int rl = item / NQ, n4 = item - rl * NQ, row = half * kFxHalf + rl;
      const long long k = k0 + row;
      double res[4] = {0.0, 0.0, 0.0, 0.0};
      if (k < K) {
        const float* zp = Z + (d0 + row) * ldz + 4 * n4;
        double z[3][4];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) z[j][e] = 0.0;
#pragma unroll 4
        for (int sp = 0; sp < splits; ++sp) {           // ordered sum of the accumulation runs, wide accumulator
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(zp + (long long)sp * split_stride + (long long)j * Npad));
            z[j][0] += (double)q.x; z[j][1] += (double)q.y; z[j][2] += (double)q.z; z[j][3] += (double)q.w;
          }
        }
        const long long yo = (k * T + t) * N + 4 * n4;
        const float4 yh = __ldg(reinterpret_cast<const float4*>(y + yo));
        const float4 yl = __ldg(reinterpret_cast<const float4*>(y_lo + yo));
        const double yy[4] = {(double)yh.x + (double)yl.x, (double)yh.y + (double)yl.y, (double)yh.z + (double)yl.z, (double)yh.w + (double)yl.w};
        const double xv = (double)xls[row];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          double acc = xv * bs[4 * n4 + e];
          acc = fma(v0, z[0][e], acc); acc = fma(v1, z[1][e], acc); acc = fma(v2, z[2][e], acc);
          const double rr = acc - yy[e];
          res[e] = rr;
          pv0 = fma(rr, z[0][e], pv0); pv1 = fma(rr, z[1][e], pv1); pv2 = fma(rr, z[2][e], pv2);
        }
      }
      double* dst = fx_rs + rl * NS + 4 * n4;
      dst[0] = res[0]; dst[1] = res[1]; dst[2] = res[2]; dst[3] = res[3];
    }
    __syncthreads();
    // per-neuron sums, trials added in order
    for (int n = threadIdx.x; n < N; n += 256) {       // N <= 160: at most one neuron per thread
#pragma unroll 8
      for (int rl = 0; rl < kFxHalf; ++rl) {
        const double rr = fx_rs[rl * NS + n];
        cs0 = fma(rr, rr, cs0);
        cs1 = fma((double)xls[half * kFxHalf + rl], rr, cs1);
      }
    }
    // residual planes R[n][t*Kp + k] (hi, then lo = R - hi): a half warp = the 16 trial pairs of this half (64 contiguous
    // bytes per neuron and plane), the two half warps of warp w take neurons 2w and 2w + 1 (+16, +32, ...); columns follow
    // Xb, zeros in the pad K <= k < Kp (the tile is zero there); Kp and k0 are even, so the pair stores are aligned
    const int pr = lane & 15;
    const long long ka = k0 + half * kFxHalf + 2 * pr;
    if (ka < Kp) {
      uint16_t* base = RV + t * Kp + ka;
      for (int n = 2 * w + (lane >> 4); n < N; n += 16) {
        const float ra = (float)fx_rs[(2 * pr) * NS + n], rb = (float)fx_rs[(2 * pr + 1) * NS + n];
        const __half2 h = __floats2half2_rn(ra, rb);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(ra - hf.x, rb - hf.y);
        *reinterpret_cast<uint32_t*>(base + (long long)n * ldr) = *reinterpret_cast<const uint32_t*>(&h);
        *reinterpret_cast<uint32_t*>(base + prd + (long long)n * ldr) = *reinterpret_cast<const uint32_t*>(&l);
      }
    }
  }
  for (int n = threadIdx.x; n < N; n += 256) {
    sse_part[(t * KB + kb) * N + n] = cs0;
    db_part[(t * KB + kb) * N + n] = cs1;
  }
  pv0 = warp_sum(pv0); pv1 = warp_sum(pv1); pv2 = warp_sum(pv2);
  if (lane == 0) { pvs[w][0] = pv0; pvs[w][1] = pv1; pvs[w][2] = pv2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double sj = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) sj += pvs[w2][threadIdx.x];
    pv_part[(t * KB + kb) * 3 + threadIdx.x] = sj * zscale;
  }
}

// per (t, n): db and SSE from the per-block partials (ordered sum over the trial blocks); sr (may be NULL) receives
// sum_k xl*R = the column sums of the residual the exact-operand backward's mean correction needs
template <typename AT>
__global__ void __launch_bounds__(128) reduce_part_kernel(const AT* __restrict__ sse_part, const AT* __restrict__ db_part,
                                                          const double* __restrict__ b, long long KB, long long T, long long N,
                                                          double l2, double* __restrict__ db, double* __restrict__ sse_tn,
                                                          float* __restrict__ sr, long long Npad) {
  const long long t = blockIdx.y;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double sdb = 0.0, sse = 0.0;
  for (long long kb = 0; kb < KB; ++kb) {
    sse += (double)sse_part[(t * KB + kb) * N + n];
    sdb += (double)db_part[(t * KB + kb) * N + n];
  }
  if (db) db[n * T + t] = 2.0 * sdb + 2.0 * l2 * b[n * T + t];
  if (sr) sr[t * Npad + n] = (float)sdb;
  sse_tn[t * N + n] = sse;
}

// final scalars: sse_n, loss, dV
template <typename AT>
__global__ void __launch_bounds__(1024) finalize_kernel(const double* __restrict__ sse_tn, const AT* __restrict__ pv_part,
                                                       const double* __restrict__ G, const double* __restrict__ W,
                                                       const double* __restrict__ V, const double* __restrict__ b, long long KB,
                                                       long long NT, long long T, long long N, int r, double l2, double* __restrict__ sse_n,
                                                       double* __restrict__ loss, double* __restrict__ dV) {
  __shared__ double red[1024];      // one block of 1024 threads: every loop below is a latency chain, so more lanes = shorter chains
  // per-neuron SSE (src/model/rrr.py:151) and its total
  double tot = 0.0;
  if (N <= 512) {
    // P threads per neuron, each over every P-th time bin; the P partial sums are then added in a fixed order
    __shared__ double sp[1024];
    const int P = 1024 / (int)N;
    const int part = threadIdx.x / (int)N, n = threadIdx.x % (int)N;
    if (part < P) {
      double s = 0.0;
#pragma unroll 4
      for (long long t = part; t < T; t += P) s += sse_tn[t * N + n];
      sp[part * N + n] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
      double s = 0.0;
      for (int q = 0; q < P; ++q) s += sp[q * N + threadIdx.x];
      if (sse_n) sse_n[threadIdx.x] = s;
      tot += s;
    }
  } else {
    for (long long n = threadIdx.x; n < N; n += 1024) {
      double s = 0.0;
#pragma unroll 10
      for (long long t = 0; t < T; ++t) s += sse_tn[t * N + n];      // independent loads: unrolled so they overlap
      if (sse_n) sse_n[n] = s;
      tot += s;
    }
  }
  // sum b^2 (the intercept is part of beta: SURVEY A14)
  double bsq = 0.0;
#pragma unroll 8
  for (long long i = threadIdx.x; i < N * T; i += 1024) bsq += b[i] * b[i];
  red[threadIdx.x] = tot + l2 * bsq;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss) {
    double gw = 0.0;
    for (int e = 0; e < r * r; ++e) gw += G[e] * W[e];
    *loss = red[0] + l2 * gw;
  }
  if (dV) {
    for (long long e = threadIdx.x; e < (long long)r * T; e += 1024) {
      const long long j = e / T, t = e % T;
      double s = 0.0;
#pragma unroll 7
      for (long long q = 0; q < KB * NT; ++q) s += (double)pv_part[(t * KB * NT + q) * r + j];
      double gv = 0.0;
      for (int jj = 0; jj < r; ++jj) gv += G[j * r + jj] * V[(long long)jj * T + t];
      dV[e] += 2.0 * s + 2.0 * l2 * gv;  // accumulated: V is shared across sessions (rrr.py:49)
    }
  }
}

// ------------------------------------------------------------------ closure stage 4: epilogue of GEMM-B
// dU[n][c][j] = 2 Gacc[c][(j,n)] + 2 l2 sum_j' U[n][c][j'] W[j'][j]; 32 c x 32 n tile, smem transpose.
// kFast (plain bf16 operands, one split): the whole combination runs in fp32 and is widened once per output -- the
// vector fp64 pipe of the B200 is narrow and the gradient carries the ~1e-3 operand rounding anyway.  Otherwise
// (residual planes / split-K partials) everything is summed in fp64.
// kCorr (exact-operand mode): the backward contracted the integers frame - round(mean) instead of frame - mean; the
// missing part is rank T:  Gacc[c,(j,n)] -= sum_t V[j,t] * qT[c,t] * SR[t,n]   with qT = (mean - round(mean))/std and
// SR[t,n] = sum_k R[k,t,n].  |mean - round(mean)| <= 1/2, so the correction is < 1 % of Gacc; it is formed in fp32.
template <bool kFast, int RMAX, bool kCorr>
__global__ void __launch_bounds__(256) epi_b_kernel(const float* __restrict__ Gacc, long long ldg, int splits,
                                                    long long split_stride, const double* __restrict__ U,
                                                    const double* __restrict__ W, long long C1, long long N, long long Npad, int r,
                                                    double l2, double* __restrict__ dU, const float* __restrict__ qT, long long ldt,
                                                    const float* __restrict__ SR, const double* __restrict__ V, long long T) {
  __shared__ float Gs[RMAX][32][33];
  __shared__ double Ws[RMAX * RMAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long c0 = (long long)blockIdx.x * 32, n0 = (long long)blockIdx.y * 32;
  if (threadIdx.x < r * r) Ws[threadIdx.x] = W[threadIdx.x];
  float corr[4][RMAX];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < RMAX; ++j) corr[i][j] = 0.f;
  if constexpr (kCorr) {
    __shared__ float qs[32][33], srs[32][33], vsj[RMAX][32];
    for (long long t0 = 0; t0 < T; t0 += 32) {
      __syncthreads();
      // qs[c][t], srs[t][n], vsj[j][t] for 32 time bins (zero past T)
      for (int e = threadIdx.x; e < 32 * 32; e += 256) {
        const int a = e >> 5, bq = e & 31;
        const long long c = c0 + a, tt = t0 + bq;
        qs[a][bq] = (c < C1 && tt < T) ? __ldg(qT + c * ldt + tt) : 0.f;
        const long long t2 = t0 + a, nn = n0 + bq;
        srs[a][bq] = (t2 < T && nn < N) ? __ldg(SR + t2 * Npad + nn) : 0.f;
      }
      for (int e = threadIdx.x; e < RMAX * 32; e += 256) {
        const int j = e >> 5, tq = e & 31;
        vsj[j][tq] = (j < r && t0 + tq < T) ? (float)V[(long long)j * T + t0 + tq] : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int tq = 0; tq < 32; ++tq) {
        const float sv = srs[tq][lane];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float qv = qs[w + 8 * i][tq] * sv;
#pragma unroll
          for (int j = 0; j < RMAX; ++j) corr[i][j] = fmaf(vsj[j][tq], qv, corr[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = w + 8 * i;
    const long long c = c0 + cl, n = n0 + lane;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      if (j >= r) break;
      float gv = 0.f;
      if (c < C1 && n < N) {
        const float* gp = Gacc + c * ldg + (long long)j * Npad + n;
        if constexpr (kFast) {
          gv = __ldg(gp);
        } else {
          double gsum = 0.0;
          for (int sp = 0; sp < splits; ++sp) gsum += (double)gp[(long long)sp * split_stride];
          gv = (float)(gsum - (double)corr[i][j]);
        }
      }
      Gs[j][cl][lane] = gv;
    }
  }
  __syncthreads();
  const long long c = c0 + lane;
  if (c >= C1) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nl = w + 8 * i;
    const long long n = n0 + nl;
    if (n >= N) continue;
    const double* up = U + (n * C1 + c) * r;
    double* gp = dU + (n * C1 + c) * r;
    if constexpr (kFast) {
      float u[RMAX];
#pragma unroll
      for (int j = 0; j < RMAX; ++j) u[j] = j < r ? (float)up[j] : 0.f;
      const float l2f = (float)(2.0 * l2);
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        if (j >= r) break;
        float reg = 0.f;
#pragma unroll
        for (int jj = 0; jj < RMAX; ++jj)
          if (jj < r) reg = fmaf(u[jj], (float)Ws[jj * r + j], reg);
        gp[j] = (double)fmaf(l2f, reg, 2.f * Gs[j][lane][nl]);
      }
    } else {
      double u[RMAX];
      for (int j = 0; j < r; ++j) u[j] = up[j];
      for (int j = 0; j < r; ++j) {
        double reg = 0.0;
        for (int jj = 0; jj < r; ++jj) reg += u[jj] * Ws[jj * r + j];
        gp[j] = 2.0 * (double)Gs[j][lane][nl] + 2.0 * l2 * reg;
      }
    }
  }
}

// epi_b_kernel<false, 4, true> for the exact-operand closure at rank 3 with a register-tiled correction: the rank-T
// correction is a small GEMM  corr[c, (j,n)] = sum_t qT[c,t] * (V[j,t] SR[t,n])  (789 M fused multiply-adds at C1 = 18,260,
// N = 144, T = 100); the kernel above spends ~3 shared-memory reads per 4 of them (0.178 ms, issue-bound at 69 %:
// profiles/r02_ncu_full_exact.csv).  Here a block owns 64 features x 32 neurons and a thread 8 features x 3 components of
// one neuron: per four time bins 8 broadcast 128-bit reads of q and 12 reads of the products P = V SR feed 96 FMAs.
__global__ void __launch_bounds__(256) epi_bx_kernel(const float* __restrict__ Gacc, long long ldg, const double* __restrict__ U,
                                                     const double* __restrict__ W, long long C1, long long N, long long Npad, double l2,
                                                     double* __restrict__ dU, const float* __restrict__ qT, long long ldt,
                                                     const float* __restrict__ SR, const double* __restrict__ V, long long T) {
  __shared__ __align__(16) float qs[64][36];
  __shared__ float ps[32][3][32];
  __shared__ float Gs[3][64][33];
  __shared__ double Ws[9];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long c0 = (long long)blockIdx.x * 64, n0 = (long long)blockIdx.y * 32;
  if (threadIdx.x < 9) Ws[threadIdx.x] = W[threadIdx.x];
  float corr[8][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) corr[i][0] = corr[i][1] = corr[i][2] = 0.f;
  for (long long t0 = 0; t0 < T; t0 += 32) {
    __syncthreads();
    // fill: thread = (row a = w + 8 i, column lane) of both tiles; V[j, t] is read once per row and broadcast
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int a = w + 8 * i;
      const long long c = c0 + a, tt = t0 + lane;
      qs[a][lane] = (c < C1 && tt < T) ? __ldg(qT + c * ldt + tt) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int tq = w + 8 * i;
      const long long tt = t0 + tq, nn = n0 + lane;
      const bool ok = tt < T && nn < N;
      const float sv = ok ? __ldg(SR + tt * Npad + nn) : 0.f;
      const long long tc = tt < T ? tt : T - 1;
      ps[tq][0][lane] = (float)V[tc] * sv;
      ps[tq][1][lane] = (float)V[T + tc] * sv;
      ps[tq][2][lane] = (float)V[2 * T + tc] * sv;
    }
    __syncthreads();
#pragma unroll 2
    for (int t4 = 0; t4 < 8; ++t4) {
      float pv[4][3];
#pragma unroll
      for (int tt = 0; tt < 4; ++tt)
#pragma unroll
        for (int j = 0; j < 3; ++j) pv[tt][j] = ps[4 * t4 + tt][j][lane];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q4 = *reinterpret_cast<const float4*>(&qs[w + 8 * i][4 * t4]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          corr[i][j] = fmaf(q4.x, pv[0][j], corr[i][j]);
          corr[i][j] = fmaf(q4.y, pv[1][j], corr[i][j]);
          corr[i][j] = fmaf(q4.z, pv[2][j], corr[i][j]);
          corr[i][j] = fmaf(q4.w, pv[3][j], corr[i][j]);
        }
      }
    }
  }
  {
    const long long n = n0 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cl = w + 8 * i;
      const long long c = c0 + cl;
      const bool ok = c < C1 && n < N;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Gs[j][cl][lane] = ok ? (float)((double)__ldg(Gacc + c * ldg + (long long)j * Npad + n) - (double)corr[i][j]) : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int cl = 32 * cc + lane;
    const long long c = c0 + cl;
    if (c >= C1) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nl = w + 8 * i;
      const long long n = n0 + nl;
      if (n >= N) continue;
      const double* up = U + (n * C1 + c) * 3;
      double* gp = dU + (n * C1 + c) * 3;
      const double u0 = up[0], u1 = up[1], u2 = up[2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double reg = u0 * Ws[j] + u1 * Ws[3 + j] + u2 * Ws[6 + j];
        gp[j] = 2.0 * (double)Gs[j][cl][nl] + 2.0 * l2 * reg;
      }
    }
  }
}

// ------------------------------------------------------------------ dense-forward mode (VS_RRR_MODE_DENSE): small kernels
// power-of-two scale of time bin t's generated coefficient tiles: |beta'_t| <= isdmax[t] * umax * sum_j |V[j,t]| =: bound;
// scale = 2^floor(log2(16384 / bound)) keeps the hi plane below the half maximum and the lo plane (2^-11 of the value)
// in the normal range for every coefficient within ~2^-9 of the largest one
__global__ void __launch_bounds__(128) bscale_kernel(const unsigned* __restrict__ umax_bits, const float* __restrict__ isdmax,
                                                     const double* __restrict__ V, int r, long long T, float* __restrict__ bscale) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float umax = __uint_as_float(*umax_bits);
  float vs1 = 0.f;
  for (int j = 0; j < r; ++j) vs1 += fabsf((float)V[(long long)j * T + t]);
  const float bound = isdmax[t] * umax * vs1;
  float sc = 1.f;
  if (bound > 0.f && isfinite(bound)) {
    int e = (int)floorf(log2f(16384.f / bound));
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    sc = exp2f((float)e);
  }
  bscale[t] = sc;
}

// M1[t][(j,n)] = (1/uscale) * ordered sum of the split-K partials of the small GEMM  q (T x C1) * Ub^T (3 Npad x C1):
// M1[t][j][n] = sum_c q[t,c] U[n,c,j] with q = (mean - round(mean))/std -- the fractional part of the mean that the exact
// integer operands leave out (constant term of the forward, correction of dV)
__global__ void __launch_bounds__(256) m1_reduce_kernel(const float* __restrict__ part, int splits, long long split_stride, long long n,
                                                        double inv_uscale, double* __restrict__ M1) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int sp = 0; sp < splits; ++sp) s += (double)part[(long long)sp * split_stride + i];
  M1[i] = s * inv_uscale;
}

// Epilogue of the dense forward: block (kb, t, nt) = 64 trials of time bin t x 32 neurons (the tiling of epi_f_kernel).
//   yhat = Y - c0[t,n] + xl * b[n,t],  c0[t,n] = sum_j V[j,t] M1[t][j][n];  R = yhat - y  (float64 throughout)
// outputs: per-block SSE / db partials, the residual as hi + lo half planes R[p][n][t*Kp + k] (B operand of the backward and
// of the dV pass), or the prediction itself (kPredict).
template <bool kPredict>
__global__ void __launch_bounds__(256) epi_d_kernel(const float* __restrict__ Y, long long ldy, int splits, long long split_stride,
                                                    const double* __restrict__ M1, long long ldm,
                                                    const float* __restrict__ y, const float* __restrict__ y_lo, const float* __restrict__ xl,
                                                    const double* __restrict__ V, const double* __restrict__ b, long long K, long long T,
                                                    long long N, long long Npad, int r, long long ldr, uint16_t* __restrict__ RV,
                                                    double* __restrict__ sse_part, double* __restrict__ db_part, double* __restrict__ yhat,
                                                    long long Kp) {
  __shared__ double Rs[kEpiRows][33];
  __shared__ double red[8][32][2];
  __shared__ float xls[kEpiRows];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long kb = blockIdx.x, t = blockIdx.y, nt = blockIdx.z, KB = gridDim.x;
  const long long k0 = kb * kEpiRows, d0 = t * K + k0, n0 = nt * 32;
  if (threadIdx.x < kEpiRows) xls[threadIdx.x] = (k0 + threadIdx.x < K) ? xl[d0 + threadIdx.x] : 0.f;
  __syncthreads();
  const long long n = n0 + lane;
  double c0 = 0.0, bn = 0.0;
  if (n < N) {
    for (int j = 0; j < r; ++j) c0 = fma(V[(long long)j * T + t], M1[t * ldm + (long long)j * Npad + n], c0);
    bn = b[n * T + t];
  }
  double sse = 0.0, sdb = 0.0;
  constexpr int RPW = kEpiRows / 8;
  double yr[RPW];
  double yv[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int rl = w * RPW + i;
    const long long k = k0 + rl;
    const bool ok = k < K && n < N;
    yr[i] = 0.0;
    if (ok)
      for (int sp = 0; sp < splits; ++sp) yr[i] += (double)__ldg(Y + (long long)sp * split_stride + (d0 + rl) * ldy + n);   // ordered sum of the runs
    if constexpr (!kPredict) {
      yv[i] = 0.0;
      if (ok) {
        yv[i] = (double)__ldg(y + (k * T + t) * N + n);
        if (y_lo) yv[i] += (double)__ldg(y_lo + (k * T + t) * N + n);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int rl = w * RPW + i;
    const long long k = k0 + rl;
    double res = 0.0;
    if (k < K && n < N) {
      const double acc = fma((double)xls[rl], bn, yr[i] - c0);
      if constexpr (kPredict) {
        yhat[(k * T + t) * N + n] = acc;
      } else {
        res = acc - yv[i];
        sse = fma(res, res, sse);
        sdb = fma((double)xls[rl], res, sdb);
      }
    }
    if constexpr (!kPredict) Rs[rl][lane] = res;
  }
  if constexpr (!kPredict) {
    red[w][lane][0] = sse;
    red[w][lane][1] = sdb;
    __syncthreads();
    const long long ka = k0 + 2 * lane;
    const long long dcol = t * Kp + ka;
    const long long prd = Npad * ldr;
    if (RV) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nl = w * 4 + i;
        const long long nn = n0 + nl;
        if (nn >= N || ka >= Kp) continue;
        float ra = (float)Rs[2 * lane][nl], rb = (float)Rs[2 * lane + 1][nl];
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          const uint32_t lo = enc16(ra, VS_OPERAND_F16), hi = enc16(rb, VS_OPERAND_F16);
          *reinterpret_cast<uint32_t*>(RV + pl * prd + nn * ldr + dcol) = lo | (hi << 16);
          ra -= dec16((uint16_t)lo, VS_OPERAND_F16); rb -= dec16((uint16_t)hi, VS_OPERAND_F16);
        }
      }
    }
    if (w == 0 && n < N) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) { s0 += red[w2][lane][0]; s1 += red[w2][lane][1]; }
      sse_part[(t * KB + kb) * N + n] = s0;
      db_part[(t * KB + kb) * N + n] = s1;
    }
  }
}

// dV[j,t] += 2 * ( sum over the dV pass's partials  -  sum_n SR[t,n] M1[t][j][n] ):  one block per (j, t); the partials are
// summed in a fixed order (thread i takes partials i, i+128, ..., then a fixed tree)
__global__ void __launch_bounds__(128) dv_reduce_kernel(const double* __restrict__ dvpart, long long nparts, const float* __restrict__ SR,
                                                        const double* __restrict__ M1, long long ldm, long long T, long long N,
                                                        long long Npad, int r, double* __restrict__ dV) {
  __shared__ double sh[128];
  const long long e = blockIdx.x;
  const long long j = e / T, t = e % T;
  double s = 0.0;
  for (long long q = threadIdx.x; q < nparts; q += 128) s += dvpart[(q * T + t) * 3 + j];
  for (long long n = threadIdx.x; n < N; n += 128) s -= (double)SR[t * Npad + n] * M1[t * ldm + j * Npad + n];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int h = 64; h > 0; h >>= 1) {
    if (threadIdx.x < h) sh[threadIdx.x] += sh[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) dV[e] += 2.0 * sh[0];
}

// ------------------------------------------------------------------ host orchestration
struct Ws {
  uint16_t *Ub, *RV;
  float *Z, *Gacc, *SR;
  float *U32, *bscale;            // dense-forward mode
  double* dvpart;
  unsigned* umax;
  double* M1;
  long long ldm, dv_parts;
  void *sse_part, *db_part, *pv_part;   // float, or double in the exact-operand mode
  void* bal;                 // tail-wave split-K scratch of the GEMMs
  double *Gp, *G, *W, *sse_tn;
  long long Npad, ldz, gp_blocks, KB;
  int splits_f, splits_b;  // split-K factors of the two GEMMs
  long long Kp;            // trials per time bin in the backward operands (K rounded up to 16)
  size_t total;
};

// The tensor cores accumulate in fp32 with truncation, so a long contraction drifts by about
// (#MMA steps) * 2^-24.  That is far below bf16 operand noise (planes == 1) but not below the
// ~2^-24 the 3-plane mode is after: there each TMEM accumulation run is limited to 16 k-blocks and
// the partial tiles are summed in fp64 by the epilogue kernels.
static int hp_splits(long long k_elems, int planes, int mode) {
  // experiment hook: VS_RRR_RUN=<k-blocks per TMEM accumulation run> forces split-K in the single-plane mode too
  static int run1 = -1;
  if (run1 < 0) { const char* e = getenv("VS_RRR_RUN"); run1 = e ? atoi(e) : 0; }
  if (mode == VS_RRR_MODE_EXACT || mode == VS_RRR_MODE_DENSE) {
    // exact-operand modes: a few accumulation runs.  The truncation of the fp32 accumulator is the largest error left in a
    // closure evaluation (per-evaluation loss error ~1.3e-8 per MMA step of full magnitude).
    // The drift is mostly a relative SHRINK of the prediction.  Near the V ~ 0 plateau the fit visits after its first steps
    // (|dV| falls from 1e7 to ~10 while sum |R Z| stays ~1e7) the shrink biases the residual by -eps * yhat, i.e. dV by a
    // non-cancelling -2 eps (Z^T Z) V: measured 0.5 (4 runs of the factorised forward would give ~0.25) against |dV| = 10
    // (profiles/r02_mode_diff_probe.txt).  Hence: VS_RRR_RUN_EXACT / VS_RRR_RUN_DENSE = k-blocks per run, defaults 72 (4 runs
    // of the factorised forward at 18,260 features) and 36 (8 runs of the dense forward, whose hi/lo products interleave and
    // truncate twice per k-step; its partial matrices are a third of the size).
    static int runx = -1, rund = -1;
    if (runx < 0) { const char* e = getenv("VS_RRR_RUN_EXACT"); runx = e ? atoi(e) : 72; if (runx <= 0) runx = 1 << 30; }
    if (rund < 0) { const char* e = getenv("VS_RRR_RUN_DENSE"); rund = e ? atoi(e) : 36; if (rund <= 0) rund = 1 << 30; }
    long long sx = ceil_div(ceil_div(k_elems, 64), mode == VS_RRR_MODE_DENSE ? rund : runx);
    return (int)(sx < 1 ? 1 : (sx > 64 ? 64 : sx));
  }
  if (planes < 2) {
    if (run1 <= 0) return 1;
    long long s1 = ceil_div(ceil_div(k_elems, 64), run1);
    return (int)(s1 < 1 ? 1 : (s1 > 256 ? 256 : s1));
  }
  long long s = ceil_div(ceil_div(k_elems, 64), 16);
  return (int)(s < 1 ? 1 : (s > 256 ? 256 : s));
}

static Ws carve(const vs_rrr_dims& d, void* base) {
  Ws w;
  const long long KT = d.K * d.T;
  w.Npad = round_up(d.N, 16);
  w.ldz = d.r * w.Npad;
  w.gp_blocks = ceil_div(d.C1, 256 * kPrepC) * d.N;
  w.KB = ceil_div(d.K, kEpiRows);
  w.splits_f = hp_splits(d.C1, d.planes, d.mode);
  w.splits_b = d.mode == VS_RRR_MODE_CLASSIC ? hp_splits(d.T * round_up(d.K, 16), d.planes, d.mode) : 1;   // exact backward: one bin per run already
  const size_t pe = d.mode != VS_RRR_MODE_CLASSIC ? 8 : 4;     // element size of the epilogue partials
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* q = p ? p + off : nullptr; off += (size_t)round_up((long long)bytes, 1024); return q; };
  w.Ub = (uint16_t*)take((size_t)d.planes * w.ldz * d.ldc * 2);
  w.Kp = round_up(d.K, 16);
  w.RV = (uint16_t*)take((size_t)d.planes * w.ldz * d.ldr * 2);
  w.Z = (float*)take((size_t)w.splits_f * KT * w.ldz * 4);
  w.Gacc = (float*)take((size_t)w.splits_b * d.C1 * w.ldz * 4);
  w.sse_part = take((size_t)d.T * w.KB * d.N * pe);
  w.db_part = take((size_t)d.T * w.KB * d.N * pe);
  w.pv_part = take((size_t)d.T * w.KB * ceil_div(d.N, 32) * d.r * pe);
  w.SR = (float*)take((size_t)d.T * w.Npad * 4);
  w.U32 = nullptr; w.bscale = nullptr; w.dvpart = nullptr; w.umax = nullptr; w.M1 = nullptr;
  w.ldm = w.ldz; w.dv_parts = 8 * tc::rrr_bwd_dense_ctas(d.C1);
  if (d.mode == VS_RRR_MODE_DENSE) {
    w.U32 = (float*)take((size_t)d.N * d.ldc * d.r * 4);
    w.bscale = (float*)take((size_t)d.T * 4);
    w.umax = (unsigned*)take(64);
    w.M1 = (double*)take((size_t)d.T * w.ldm * 8);
    w.dvpart = (double*)take((size_t)w.dv_parts * d.T * 3 * 8);
  }
  w.bal = take(tc::balance_ws_bytes());
  w.Gp = (double*)take((size_t)w.gp_blocks * d.r * d.r * 8);
  w.G = (double*)take(kMaxR * kMaxR * 8);
  w.W = (double*)take(kMaxR * kMaxR * 8);
  w.sse_tn = (double*)take((size_t)d.T * d.N * 8);
  w.total = off;
  return w;
}

static int check_dims(const vs_rrr_dims& d) {
  VS_REQUIRE(d.K > 0 && d.T > 0 && d.C1 > 0 && d.N > 0 && d.r > 0, VS_ERR_INVALID, "rrr: empty dimension");
  VS_REQUIRE(d.r <= kMaxR, VS_ERR_UNSUPPORTED, "rrr: rank %lld > %d", (long long)d.r, kMaxR);
  VS_REQUIRE(d.planes >= 1 && d.planes <= 3, VS_ERR_INVALID, "rrr: planes must be 1..3");
  VS_REQUIRE(d.fmt == VS_OPERAND_BF16 || d.fmt == VS_OPERAND_F16, VS_ERR_INVALID, "rrr: fmt must be VS_OPERAND_BF16 or VS_OPERAND_F16");
  VS_REQUIRE(d.ldc >= d.C1 && d.ldc % 8 == 0 && d.ldr >= d.T * round_up(d.K, 16) && d.ldr % 8 == 0, VS_ERR_INVALID, "rrr: bad pitches");
  VS_REQUIRE(d.K * d.T < (1ll << 31) && d.C1 < (1ll << 31), VS_ERR_UNSUPPORTED, "rrr: dimension exceeds 2^31");
  VS_REQUIRE(d.mode == VS_RRR_MODE_CLASSIC || d.mode == VS_RRR_MODE_EXACT || d.mode == VS_RRR_MODE_DENSE, VS_ERR_INVALID,
             "rrr: mode must be VS_RRR_MODE_*");
  VS_REQUIRE(d.mode == VS_RRR_MODE_CLASSIC || (d.planes == 2 && d.fmt == VS_OPERAND_F16), VS_ERR_INVALID,
             "rrr: the exact-operand modes use two IEEE-half planes (planes = 2, fmt = VS_OPERAND_F16)");
  VS_REQUIRE(d.mode != VS_RRR_MODE_DENSE || d.r == 3, VS_ERR_UNSUPPORTED, "rrr: the dense-forward mode is built for rank 3");
  return VS_OK;
}

static void set_passes(tc::GemmDesc& g, int planes) {
  // plane products kept: everything down to ~2^-8(planes) relative.  Small corrections first, the
  // dominant (0,0) product last: the accumulator is truncated at every MMA, so it should be small
  // for as many of those truncations as possible.
  static const int pa3[6] = {2, 1, 0, 1, 0, 0}, pb3[6] = {0, 1, 2, 0, 1, 0};
  static const int pa2[3] = {1, 0, 0}, pb2[3] = {0, 1, 0};
  if (planes == 1) { g.n_pass = 1; g.pa[0] = g.pb[0] = 0; return; }
  g.n_pass = planes == 2 ? 3 : 6;
  for (int i = 0; i < g.n_pass; ++i) { g.pa[i] = planes == 2 ? pa2[i] : pa3[i]; g.pb[i] = planes == 2 ? pb2[i] : pb3[i]; }
}

// Z = Xa * Ub^T  (M = K*T, N = r*Npad, contraction C1)
static int gemm_f(const vs_rrr_dims& d, const uint16_t* Xa, const Ws& w, int engine, cudaStream_t st, int* splits_used) {
  *splits_used = 1;
  const long long KT = d.K * d.T;
  if (engine == VS_ENGINE_SIMT) {
    simt::GemmDesc g;
    g.A.ptr = Xa; g.A.type = d.fmt == VS_OPERAND_F16 ? simt::F16 : simt::BF16; g.A.s_i = d.ldc; g.A.s_k = 1; g.A.planes = d.planes; g.A.plane_stride = KT * d.ldc;
    g.B.ptr = w.Ub; g.B.type = d.fmt == VS_OPERAND_F16 ? simt::F16 : simt::BF16; g.B.s_i = d.ldc; g.B.s_k = 1; g.B.planes = d.planes; g.B.plane_stride = w.ldz * d.ldc;
    g.M = KT; g.N = w.ldz; g.K = d.C1; g.C = w.Z; g.ldc = w.ldz;
    return simt::gemm(g, st);
  }
  tc::GemmDesc g;
  g.A.ptr = Xa; g.A.rows = KT; g.A.k = d.C1; g.A.ld = d.ldc; g.A.planes = d.planes; g.A.plane_stride = KT * d.ldc;
  g.B.ptr = w.Ub; g.B.rows = w.ldz; g.B.k = d.C1; g.B.ld = d.ldc; g.B.planes = d.planes; g.B.plane_stride = w.ldz * d.ldc;
  g.M = KT; g.N = w.ldz; g.K = d.C1; g.C = w.Z; g.ldc = w.ldz;
  g.splits = w.splits_f; g.split_stride = KT * w.ldz; g.splits_out = splits_used;
  g.balance_ws = w.bal;
  g.f16 = d.fmt == VS_OPERAND_F16;
  set_passes(g, d.planes);
  return tc::gemm_tn(g, st);
}

// Gacc = Xb * RV^T  (M = C1, N = r*Npad, contraction K*T)
static int gemm_b(const vs_rrr_dims& d, const uint16_t* Xb, const Ws& w, int engine, cudaStream_t st, int* splits_used) {
  *splits_used = 1;
  const long long KT = w.Kp * d.T;       // padded time bins
  if (engine == VS_ENGINE_SIMT) {
    simt::GemmDesc g;
    g.A.ptr = Xb; g.A.type = d.fmt == VS_OPERAND_F16 ? simt::F16 : simt::BF16; g.A.s_i = d.ldr; g.A.s_k = 1; g.A.planes = d.planes; g.A.plane_stride = d.C1 * d.ldr;
    g.B.ptr = w.RV; g.B.type = d.fmt == VS_OPERAND_F16 ? simt::F16 : simt::BF16; g.B.s_i = d.ldr; g.B.s_k = 1; g.B.planes = d.planes; g.B.plane_stride = w.ldz * d.ldr;
    g.M = d.C1; g.N = w.ldz; g.K = KT; g.C = w.Gacc; g.ldc = w.ldz;
    return simt::gemm(g, st);
  }
  tc::GemmDesc g;
  g.A.ptr = Xb; g.A.rows = d.C1; g.A.k = KT; g.A.ld = d.ldr; g.A.planes = d.planes; g.A.plane_stride = d.C1 * d.ldr;
  g.B.ptr = w.RV; g.B.rows = w.ldz; g.B.k = KT; g.B.ld = d.ldr; g.B.planes = d.planes; g.B.plane_stride = w.ldz * d.ldr;
  g.M = d.C1; g.N = w.ldz; g.K = KT; g.C = w.Gacc; g.ldc = w.ldz;
  g.splits = w.splits_b; g.split_stride = d.C1 * w.ldz; g.splits_out = splits_used;
  g.balance_ws = w.bal;
  g.f16 = d.fmt == VS_OPERAND_F16;
  set_passes(g, d.planes);
  return tc::gemm_tn(g, st);
}

}  // namespace rrr
}  // namespace vs

using namespace vs;
using namespace vs::rrr;

extern "C" int64_t vs_rrr_ldc(int64_t C1) { return round_up(C1, 64); }
extern "C" int64_t vs_rrr_ldr(int64_t K, int64_t T) { return round_up(T * round_up(K, 16), 64); }

extern "C" size_t vs_rrr_workspace(vs_rrr_dims d) {
  if (d.K <= 0 || d.T <= 0 || d.C1 <= 0 || d.N <= 0 || d.r <= 0 || d.planes < 1) return 0;
  return carve(d, nullptr).total;
}

extern "C" int vs_rrr_pack(const double* X_trials, int64_t k0, int64_t nk, vs_rrr_dims d, uint16_t* Xa, uint16_t* Xb,
                           float* xl, int32_t* overflow_flag, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  VS_REQUIRE(X_trials && Xa && Xb && xl, VS_ERR_INVALID, "vs_rrr_pack: null pointer");
  VS_REQUIRE(k0 >= 0 && nk > 0 && k0 + nk <= d.K, VS_ERR_INVALID, "vs_rrr_pack: trial range outside the matrix");
  dim3 grid((unsigned)(ceil_div(nk, 32) * d.T), (unsigned)ceil_div(d.C1, 32));
  VS_REQUIRE(grid.y <= 65535u, VS_ERR_UNSUPPORTED, "vs_rrr_pack: too many columns");
  VS_LAUNCH((pack_kernel<false>), grid, 256, 0, stream, X_trials, nullptr, nullptr, nullptr, nullptr, 0ll, (long long)k0,
            (long long)(k0 + nk), (long long)d.K, (long long)d.T, (long long)d.C1, d.planes, (int)d.fmt, (long long)d.ldc, (long long)d.ldr,
            Xa, Xb, xl, overflow_flag);
  if (k0 + nk == d.K && round_up(d.K, 16) > d.K)
    VS_LAUNCH(pad_zero_kernel, (unsigned)ceil_div((long long)d.planes * d.C1 * d.T, 256), 256, 0, stream, Xb, (long long)d.planes * d.C1,
              (long long)d.T, (long long)d.K, (long long)round_up(d.K, 16), (long long)d.ldr);
  return VS_OK;
}

extern "C" int vs_rrr_colstats(const uint8_t* frames, int64_t K, int64_t cols, double* mean, double* std_clipped, void* stream) {
  VS_REQUIRE(frames && mean && std_clipped && K > 0 && cols > 0, VS_ERR_INVALID, "vs_rrr_colstats: bad arguments");
  VS_LAUNCH(colstats_kernel, (unsigned)ceil_div(cols, 256), 256, 0, stream, frames, (long long)K, (long long)cols, mean, std_clipped);
  return VS_OK;
}

extern "C" int vs_rrr_pack_u8(const uint8_t* frames, int64_t Tf, const int32_t* sorted_idx, const double* mean,
                              const double* std_clipped, vs_rrr_dims d, uint16_t* Xa, uint16_t* Xb, float* xl,
                              int32_t* overflow_flag, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  VS_REQUIRE(frames && sorted_idx && mean && std_clipped && Xa && Xb && xl && Tf >= d.T, VS_ERR_INVALID, "vs_rrr_pack_u8: bad arguments");
  if (d.planes == 1 && d.ldr % 2 == 0) {
    dim3 gw((unsigned)(ceil_div(d.K, 32) * d.T), (unsigned)ceil_div(d.C1, 128));
    VS_REQUIRE(gw.y <= 65535u, VS_ERR_UNSUPPORTED, "vs_rrr_pack_u8: too many columns");
    VS_LAUNCH(pack_u8_wide_kernel, gw, 256, 0, stream, frames, sorted_idx, mean, std_clipped, (long long)Tf, (long long)d.K, (long long)d.T,
              (long long)d.C1, (int)d.fmt, (long long)d.ldc, (long long)d.ldr, Xa, Xb, xl, overflow_flag);
    if (round_up(d.K, 16) > d.K)
      VS_LAUNCH(pad_zero_kernel, (unsigned)ceil_div((long long)d.planes * d.C1 * d.T, 256), 256, 0, stream, Xb, (long long)d.planes * d.C1,
                (long long)d.T, (long long)d.K, (long long)round_up(d.K, 16), (long long)d.ldr);
    return VS_OK;
  }
  dim3 grid((unsigned)(ceil_div(d.K, 32) * d.T), (unsigned)ceil_div(d.C1, 32));
  VS_REQUIRE(grid.y <= 65535u, VS_ERR_UNSUPPORTED, "vs_rrr_pack_u8: too many columns");
  VS_LAUNCH((pack_kernel<true>), grid, 256, 0, stream, nullptr, frames, sorted_idx, mean, std_clipped, (long long)Tf, 0ll,
            (long long)d.K, (long long)d.K, (long long)d.T, (long long)d.C1, d.planes, (int)d.fmt, (long long)d.ldc, (long long)d.ldr, Xa, Xb,
            xl, overflow_flag);
  if (round_up(d.K, 16) > d.K)
    VS_LAUNCH(pad_zero_kernel, (unsigned)ceil_div((long long)d.planes * d.C1 * d.T, 256), 256, 0, stream, Xb, (long long)d.planes * d.C1,
              (long long)d.T, (long long)d.K, (long long)round_up(d.K, 16), (long long)d.ldr);
  return VS_OK;
}

// exact-operand extras of one split (vs_rrr_closure_exact); all NULL in the classic mode
struct ExactArgs {
  const float* isdT = nullptr;   // (C1, ldt): 1/std[t,c]
  const float* qT = nullptr;     // (C1, ldt): (mean - round(mean))[t,c] / std[t,c]
  long long ldt = 0;
  const float* y_lo = nullptr;   // (K,T,N): y = y + y_lo at float64 precision (may be NULL)
  // dense-forward mode
  const uint16_t* Xc = nullptr;  // (K*T, ldc) half exact integers (forward A operand)
  const float* isd = nullptr;    // (T, ldc)
  const uint16_t* qh = nullptr;  // (2, Tq, ldc) half planes of q
  const float* isdmax = nullptr; // (T)
};

// Forward of the dense mode up to the raw accumulators: U planes + U32 + Gram partials, scales, the small GEMM M1 = q U,
// then Y = Xc beta' (tc::rrr_fwd_dense).  Y lands in w.Z (pitch Npad), M1 in w.M1.
static int dense_forward(const vs_rrr_dims& d, const ExactArgs& ex, const Ws& w, const double* U, const double* V, bool want_gram,
                         cudaStream_t st, int* splits_used) {
  const int r = (int)d.r;
  VS_REQUIRE(ex.Xc && ex.isd && ex.qh && ex.isdmax, VS_ERR_INVALID, "vs_rrr_closure_exact: the dense-forward mode needs Xc, isd, qh and isdmax");
  VS_CHECK_CUDA(cudaMemsetAsync(w.umax, 0, 4, st));
  dim3 g0((unsigned)ceil_div(d.C1, 256 * kPrepC), (unsigned)d.N);
  VS_LAUNCH(prep_u_kernel<3>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub,
            want_gram ? w.Gp : (double*)nullptr, kExactUScale, w.U32, (long long)d.ldc, w.umax);
  VS_LAUNCH(bscale_kernel, (unsigned)ceil_div(d.T, 128), 128, 0, st, w.umax, ex.isdmax, V, r, (long long)d.T, w.bscale);
  // M1 partials (split-K of the (T x C1) x (C1 x 3 Npad) product over ~all SMs) live in the Gacc buffer until reduced
  const long long Tq = round_up(d.T, 16);
  tc::GemmDesc g;
  g.A.ptr = ex.qh; g.A.rows = d.T; g.A.k = d.C1; g.A.ld = d.ldc; g.A.planes = 2; g.A.plane_stride = Tq * d.ldc;
  g.B.ptr = w.Ub; g.B.rows = w.ldz; g.B.k = d.C1; g.B.ld = d.ldc; g.B.planes = 2; g.B.plane_stride = w.ldz * d.ldc;
  g.M = d.T; g.N = w.ldz; g.K = d.C1; g.C = w.Gacc; g.ldc = w.ldz;
  long long want = kNumSMs;
  if (want * d.T > d.C1) want = d.C1 / d.T;             // the partials must fit the Gacc buffer (C1 x ldz floats)
  if (want < 1) want = 1;
  int splits = 1;
  g.splits = (int)want; g.split_stride = d.T * w.ldz; g.splits_out = &splits;
  g.f16 = true;
  set_passes(g, 2);
  int rc = tc::gemm_tn(g, st);
  if (rc) return rc;
  VS_LAUNCH(m1_reduce_kernel, (unsigned)ceil_div(d.T * w.ldz, 256), 256, 0, st, w.Gacc, splits, (long long)(d.T * w.ldz), (long long)(d.T * w.ldz),
            1.0 / kExactUScale, w.M1);
  tc::DenseFwdDesc f;
  f.Xc = ex.Xc; f.K = d.K; f.T = d.T; f.C1 = d.C1; f.N = d.N; f.Npad = w.Npad; f.ldc = d.ldc;
  f.U32 = w.U32; f.isd = ex.isd; f.ldu = d.ldc; f.V = V; f.bscale = w.bscale; f.Y = w.Z; f.ldy = w.Npad;
  f.splits = w.splits_f; f.split_stride = d.K * d.T * w.Npad; f.splits_out = splits_used;
  return tc::rrr_fwd_dense(f, st);
}

// One closure evaluation in the dense-forward mode: every contraction is the reference's own (src/model/rrr.py:105-116 and
// its autograd), one time bin at a time, with the exact integer frames as the large operand:
//   forward  Y_t = Xc_t beta'_t^T            tc::rrr_fwd_dense   (coefficient tiles generated on chip, hi + lo)
//   dU       G_j = sum_t V[j,t]/std D_t      tc::rrr_bwd_dense   (D_t = Xc_t^T R_t, R as hi + lo planes)
//   dV       sum_c,n U/std D_t               tc::rrr_bwd_dense with dvpart (the same D_t tiles contracted with U)
// plus the rank-T terms the fractional part of the mean contributes (M1 = q U).
static int closure_dense(const vs_rrr_dims& d, const uint16_t* Xi, const ExactArgs& ex, const float* xl, const float* y, const double* U,
                         const double* V, const double* b, double l2, double* loss, double* sse_n, double* dU, double* dV, double* db,
                         const Ws& w, cudaStream_t st) {
  const int r = (int)d.r;
  int sf = 1;
  int rc = dense_forward(d, ex, w, U, V, true, st, &sf);
  if (rc) return rc;
  VS_LAUNCH(small_mats_kernel, r * r + 1, 256, 0, st, w.Gp, w.gp_blocks, V, r, (long long)d.T, w.G, w.W);
  dim3 ge((unsigned)w.KB, (unsigned)d.T, (unsigned)ceil_div(d.N, 32));
  VS_LAUNCH((epi_d_kernel<false>), ge, 256, 0, st, w.Z, w.Npad, sf, (long long)(d.K * d.T * w.Npad), w.M1, w.ldm, y, ex.y_lo, xl, V, b, (long long)d.K, (long long)d.T, (long long)d.N,
            w.Npad, r, (long long)d.ldr, (dU || dV) ? w.RV : (uint16_t*)nullptr, (double*)w.sse_part, (double*)w.db_part, (double*)nullptr, w.Kp);
  dim3 g2((unsigned)ceil_div(d.N, 128), (unsigned)d.T);
  VS_LAUNCH(reduce_part_kernel<double>, g2, 128, 0, st, (const double*)w.sse_part, (const double*)w.db_part, b, w.KB, (long long)d.T,
            (long long)d.N, l2, db, w.sse_tn, w.SR, w.Npad);
  // loss, per-neuron SSE, and the penalty part of dV (the data part comes from the dV pass below: no Z exists here)
  VS_LAUNCH(finalize_kernel<double>, 1, 1024, 0, st, w.sse_tn, (const double*)nullptr, w.G, w.W, V, b, 0ll, 0ll, (long long)d.T,
            (long long)d.N, r, l2, sse_n, loss, dV);
  if (!dU && !dV) return VS_OK;
  VS_REQUIRE(Xi && ex.isdT && ex.qT, VS_ERR_INVALID, "vs_rrr_closure_exact: gradients need the backward operand Xi and the scale tables");
  tc::DenseBwdDesc dd;
  dd.Xb = Xi; dd.R = w.RV; dd.C1 = d.C1; dd.K = d.K; dd.Kp = w.Kp; dd.T = d.T; dd.Npad = w.Npad; dd.ldr = d.ldr;
  dd.r = r; dd.f16 = true; dd.V = V; dd.G = w.Gacc; dd.ldg = w.ldz;
  dd.r_planes = 2; dd.r_plane_stride = w.Npad * d.ldr; dd.scaleT = ex.isdT; dd.ldt = ex.ldt;
  VS_REQUIRE(tc::rrr_bwd_dense_supported(dd), VS_ERR_UNSUPPORTED, "vs_rrr_closure_exact: shape outside the dense backward kernel");
  if (dV) {
    tc::DenseBwdDesc dv = dd;
    dv.dv_U32 = w.U32; dv.dv_ldu = d.ldc; dv.dv_N = d.N; dv.dvpart = w.dvpart;
    rc = tc::rrr_bwd_dense(dv, st);
    if (rc) return rc;
    VS_LAUNCH(dv_reduce_kernel, (unsigned)((long long)r * d.T), 128, 0, st, w.dvpart, w.dv_parts, w.SR, w.M1, w.ldm, (long long)d.T,
              (long long)d.N, w.Npad, r, dV);
  }
  if (dU) {
    rc = tc::rrr_bwd_dense(dd, st);
    if (rc) return rc;
    dim3 gx((unsigned)ceil_div(d.C1, 64), (unsigned)ceil_div(d.N, 32));
    VS_LAUNCH(epi_bx_kernel, gx, 256, 0, st, w.Gacc, w.ldz, U, w.W, (long long)d.C1, (long long)d.N, w.Npad, l2, dU, ex.qT, ex.ldt, w.SR, V,
              (long long)d.T);
  }
  return VS_OK;
}

static int closure_impl(vs_rrr_dims d, const uint16_t* Xa, const uint16_t* Xb, const ExactArgs& ex, const float* xl, const float* y,
                        const double* U, const double* V, const double* b, double l2, double* loss, double* sse_n,
                        double* dU, double* dV, double* db, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  const bool dense_fwd = d.mode == VS_RRR_MODE_DENSE;
  const bool exact = d.mode == VS_RRR_MODE_EXACT || dense_fwd;
  VS_REQUIRE((Xa || dense_fwd) && xl && y && U && V && b, VS_ERR_INVALID, "vs_rrr_closure: null pointer");
  VS_REQUIRE(!dU || Xb, VS_ERR_INVALID, "vs_rrr_closure: dU needs Xb");
  VS_REQUIRE(workspace && workspace_bytes >= vs_rrr_workspace(d), VS_ERR_WORKSPACE, "vs_rrr_closure: workspace too small (%zu < %zu)",
             workspace_bytes, vs_rrr_workspace(d));
  VS_REQUIRE(((uintptr_t)workspace & 1023) == 0, VS_ERR_INVALID, "vs_rrr_closure: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = carve(d, workspace);
  const long long KT = d.K * d.T;
  const int r = (int)d.r;
  // half planes: U ~ 1/sqrt(T r) would put its lo plane (2^-11 of the value) into the half subnormals; a power-of-two
  // scale keeps both planes normal and is undone exactly in the epilogue
  const double uscale = exact ? kExactUScale : 1.0;
  float* u32 = nullptr; long long ldu = 0; unsigned* umax = nullptr;
  if (dense_fwd) return closure_dense(d, Xb, ex, xl, y, U, V, b, l2, loss, sse_n, dU, dV, db, w, st);
  // stage 0: U planes + Gram partials, then G and W = V V^T
  dim3 g0((unsigned)ceil_div(d.C1, 256 * kPrepC), (unsigned)d.N);
  if (r == 3) {
    VS_LAUNCH(prep_u_kernel<3>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub, w.Gp, uscale, u32, ldu, umax);
  } else if (r <= 4) {
    VS_LAUNCH(prep_u_kernel<4>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub, w.Gp, uscale, u32, ldu, umax);
  } else {
    VS_LAUNCH(prep_u_kernel<kMaxR>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub, w.Gp, uscale, u32, ldu, umax);
  }
  VS_LAUNCH(small_mats_kernel, r * r + 1, 256, 0, st, w.Gp, w.gp_blocks, V, r, (long long)d.T, w.G, w.W);
  // stage 1: Z
  int sf = 1, sb = 1;
  rc = gemm_f(d, Xa, w, engine, st, &sf);
  if (rc) return rc;
  // stage 2: residuals -> backward operand (R (x) V, or plain R for the dense backward), per-block partials of SSE / db / dV
  tc::DenseBwdDesc dd;
  dd.Xb = Xb; dd.R = w.RV; dd.C1 = d.C1; dd.K = d.K; dd.Kp = w.Kp; dd.T = d.T; dd.Npad = w.Npad; dd.ldr = d.ldr;
  dd.r = r; dd.f16 = d.fmt == VS_OPERAND_F16; dd.V = V; dd.G = w.Gacc; dd.ldg = w.ldz;
  if (exact) { dd.r_planes = 2; dd.r_plane_stride = w.Npad * d.ldr; dd.scaleT = ex.isdT; dd.ldt = ex.ldt; }
  const bool dense_ok = engine != VS_ENGINE_SIMT && w.splits_b == 1 && tc::rrr_bwd_dense_supported(dd);
  if (exact && dU)
    VS_REQUIRE(dense_ok && ex.isdT && ex.qT, VS_ERR_UNSUPPORTED,
               "vs_rrr_closure_exact: the exact-operand backward needs the tcgen05 dense kernel (r = 3, N <= 160) and the scale tables");
  const bool dense = dU && (exact || d.planes == 1) && dense_ok;
  dim3 ge((unsigned)w.KB, (unsigned)d.T, (unsigned)ceil_div(d.N, 32));
#define VS_EPI_F(RM, AT) VS_LAUNCH((epi_f_kernel<false, RM, AT>), ge, 256, 0, st, w.Z, w.ldz, sf, KT * w.ldz, y, ex.y_lo, xl, V, b,       \
                                   (long long)d.K, (long long)d.T, (long long)d.N, w.Npad, r, d.planes, (int)d.fmt, (long long)d.ldr, w.RV,   \
                                   (AT*)w.sse_part, (AT*)w.db_part, (AT*)w.pv_part, nullptr, w.Kp, dense ? 1 : 0, 1.0 / uscale)
  // exact-operand training closure at rank 3: the wide-load epilogue (one block per 64 trials x all neurons: NT = 1)
  static int fx_env = -1;
  if (fx_env < 0) { const char* e = getenv("VS_RRR_EPI_FX"); fx_env = e ? atoi(e) : 1; }
  const size_t fx_smem = (size_t)kFxHalf * (d.N + 1) * 8;
  const bool use_fx = exact && dense && r == 3 && d.N % 4 == 0 && d.N <= 160 && d.planes == 2 && d.fmt == VS_OPERAND_F16 && ex.y_lo && fx_env != 0 &&
                      (((uintptr_t)y | (uintptr_t)ex.y_lo) & 15) == 0 &&
                      kFxRows == kEpiRows && fx_smem <= 96 * 1024;
  if (use_fx) {
    static bool fx_attr = false;
    if (!fx_attr) { VS_CHECK_CUDA(cudaFuncSetAttribute(epi_fx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); fx_attr = true; }
    dim3 gx((unsigned)w.KB, (unsigned)d.T);
    VS_LAUNCH(epi_fx_kernel, gx, 256, fx_smem, st, w.Z, w.ldz, sf, KT * w.ldz, y, ex.y_lo, xl, V, b, (long long)d.K, (long long)d.T, (long long)d.N,
              w.Npad, (long long)d.ldr, w.RV, (double*)w.sse_part, (double*)w.db_part, (double*)w.pv_part, w.Kp, 1.0 / uscale);
  } else if (exact) { if (r <= 4) { VS_EPI_F(4, double); } else { VS_EPI_F(kMaxR, double); } }
  else { if (r <= 4) { VS_EPI_F(4, float); } else { VS_EPI_F(kMaxR, float); } }
#undef VS_EPI_F
  dim3 g2((unsigned)ceil_div(d.N, 128), (unsigned)d.T);
  const long long NT = use_fx ? 1 : ceil_div(d.N, 32);
  if (exact) {
    VS_LAUNCH(reduce_part_kernel<double>, g2, 128, 0, st, (const double*)w.sse_part, (const double*)w.db_part, b, w.KB, (long long)d.T,
              (long long)d.N, l2, db, w.sse_tn, w.SR, w.Npad);
    VS_LAUNCH(finalize_kernel<double>, 1, 1024, 0, st, w.sse_tn, (const double*)w.pv_part, w.G, w.W, V, b, w.KB, NT, (long long)d.T,
              (long long)d.N, r, l2, sse_n, loss, dV);
  } else {
    VS_LAUNCH(reduce_part_kernel<float>, g2, 128, 0, st, (const float*)w.sse_part, (const float*)w.db_part, b, w.KB, (long long)d.T,
              (long long)d.N, l2, db, w.sse_tn, (float*)nullptr, w.Npad);
    VS_LAUNCH(finalize_kernel<float>, 1, 1024, 0, st, w.sse_tn, (const float*)w.pv_part, w.G, w.W, V, b, w.KB, NT, (long long)d.T,
              (long long)d.N, r, l2, sse_n, loss, dV);
  }
  if (dU) {
    // stage 3/4: Gacc and dU
    rc = dense ? tc::rrr_bwd_dense(dd, st) : gemm_b(d, Xb, w, engine, st, &sb);
    if (rc) return rc;
    dim3 g4((unsigned)ceil_div(d.C1, 32), (unsigned)ceil_div(d.N, 32));
    const bool fast = d.planes == 1 && sb == 1;
#define VS_EPI_B(FAST, RM, CORR) VS_LAUNCH((epi_b_kernel<FAST, RM, CORR>), g4, 256, 0, st, w.Gacc, w.ldz, sb, (long long)d.C1 * w.ldz, U, w.W, \
                                           (long long)d.C1, (long long)d.N, w.Npad, r, l2, dU, ex.qT, ex.ldt, w.SR, V, (long long)d.T)
    static int bx_env = -1;
    if (bx_env < 0) { const char* e = getenv("VS_RRR_EPI_BX"); bx_env = e ? atoi(e) : 1; }
    if (exact && r == 3 && sb == 1 && bx_env != 0) {
      dim3 gx((unsigned)ceil_div(d.C1, 64), (unsigned)ceil_div(d.N, 32));
      VS_LAUNCH(epi_bx_kernel, gx, 256, 0, st, w.Gacc, w.ldz, U, w.W, (long long)d.C1, (long long)d.N, w.Npad, l2, dU, ex.qT, ex.ldt, w.SR, V,
                (long long)d.T);
    } else if (exact) { if (r <= 4) { VS_EPI_B(false, 4, true); } else { VS_EPI_B(false, kMaxR, true); } }
    else if (r <= 4) { if (fast) { VS_EPI_B(true, 4, false); } else { VS_EPI_B(false, 4, false); } }
    else { if (fast) { VS_EPI_B(true, kMaxR, false); } else { VS_EPI_B(false, kMaxR, false); } }
#undef VS_EPI_B
  }
  return VS_OK;
}

extern "C" int vs_rrr_closure(vs_rrr_dims d, const uint16_t* Xa, const uint16_t* Xb, const float* xl, const float* y,
                              const double* U, const double* V, const double* b, double l2, double* loss, double* sse_n,
                              double* dU, double* dV, double* db, int engine, void* workspace, size_t workspace_bytes,
                              void* stream) {
  VS_REQUIRE(d.mode == VS_RRR_MODE_CLASSIC, VS_ERR_INVALID, "vs_rrr_closure: exact-operand splits go through vs_rrr_closure_exact");
  return closure_impl(d, Xa, Xb, ExactArgs(), xl, y, U, V, b, l2, loss, sse_n, dU, dV, db, engine, workspace, workspace_bytes, stream);
}

static ExactArgs exact_args(const vs_rrr_exact_ops* o) {
  ExactArgs ex;
  if (o) {
    ex.isdT = o->isdT; ex.qT = o->qT; ex.ldt = o->ldt; ex.y_lo = o->y_lo;
    ex.Xc = o->Xc; ex.isd = o->isd; ex.qh = o->qh; ex.isdmax = o->isdmax;
  }
  return ex;
}

extern "C" int vs_rrr_closure_exact(vs_rrr_dims d, const uint16_t* Xa, const vs_rrr_exact_ops* ops, const float* xl, const float* y,
                                    const double* U, const double* V, const double* b, double l2, double* loss, double* sse_n, double* dU,
                                    double* dV, double* db, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(d.mode == VS_RRR_MODE_EXACT || d.mode == VS_RRR_MODE_DENSE, VS_ERR_INVALID,
             "vs_rrr_closure_exact: dims.mode must be VS_RRR_MODE_EXACT or VS_RRR_MODE_DENSE");
  VS_REQUIRE(ops, VS_ERR_INVALID, "vs_rrr_closure_exact: null operand table");
  return closure_impl(d, Xa, ops->Xi, exact_args(ops), xl, y, U, V, b, l2, loss, sse_n, dU, dV, db, VS_ENGINE_AUTO, workspace, workspace_bytes,
                      stream);
}

extern "C" int64_t vs_rrr_ldt(int64_t T) { return round_up(T, 4); }

// shapes the exact-operand closures cover (their backward is the tcgen05 dense per-time-bin kernel); callers fall back to
// the classic mode with 3 residual planes elsewhere
extern "C" int vs_rrr_exact_supported(int64_t K, int64_t T, int64_t C1, int64_t N, int64_t r) {
  const long long Npad = round_up(N, 16);
  return (r == 3 && Npad <= 160 && C1 > 128 && K > 0 && T > 0 && T * round_up(K, 16) + 64 < (1ll << 31)) ? 1 : 0;
}

extern "C" int vs_rrr_pack_u8_exact(const uint8_t* frames, int64_t Tf, const int32_t* sorted_idx, const double* mean,
                                    const double* std_clipped, vs_rrr_dims d, uint16_t* Xa, const vs_rrr_exact_ops* out, float* xl,
                                    int32_t* overflow_flag, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  VS_REQUIRE(d.mode == VS_RRR_MODE_EXACT || d.mode == VS_RRR_MODE_DENSE, VS_ERR_INVALID, "vs_rrr_pack_u8_exact: dims.mode must be an exact-operand mode");
  VS_REQUIRE(frames && sorted_idx && mean && std_clipped && out && xl && Tf >= d.T, VS_ERR_INVALID, "vs_rrr_pack_u8_exact: bad arguments");
  VS_REQUIRE(d.mode == VS_RRR_MODE_EXACT ? Xa != nullptr : out->Xc != nullptr, VS_ERR_INVALID,
             "vs_rrr_pack_u8_exact: the forward operand of the mode is missing (Xa for EXACT, ops.Xc for DENSE)");
  VS_REQUIRE((out->isdT == nullptr) == (out->qT == nullptr), VS_ERR_INVALID, "vs_rrr_pack_u8_exact: isdT and qT go together");
  VS_REQUIRE(d.ldr % 2 == 0, VS_ERR_INVALID, "vs_rrr_pack_u8_exact: ldr must be even");
  uint16_t* Xi = const_cast<uint16_t*>(out->Xi);
  uint16_t* Xc = d.mode == VS_RRR_MODE_DENSE ? const_cast<uint16_t*>(out->Xc) : nullptr;
  dim3 gw((unsigned)(ceil_div(d.K, 32) * d.T), (unsigned)ceil_div(d.C1, 128));
  VS_REQUIRE(gw.y <= 65535u, VS_ERR_UNSUPPORTED, "vs_rrr_pack_u8_exact: too many columns");
  VS_LAUNCH(pack_u8_exact_kernel, gw, 256, 0, stream, frames, sorted_idx, mean, std_clipped, (long long)Tf, (long long)d.K, (long long)d.T,
            (long long)d.C1, (long long)d.ldc, (long long)d.ldr, d.mode == VS_RRR_MODE_EXACT ? Xa : (uint16_t*)nullptr, Xc, Xi, xl, overflow_flag);
  if (Xi && round_up(d.K, 16) > d.K)
    VS_LAUNCH(pad_zero_kernel, (unsigned)ceil_div((long long)d.C1 * d.T, 256), 256, 0, stream, Xi, (long long)d.C1, (long long)d.T,
              (long long)d.K, (long long)round_up(d.K, 16), (long long)d.ldr);
  if (out->isdT) {
    const long long ldt = round_up(d.T, 4);
    VS_REQUIRE(out->ldt == ldt, VS_ERR_INVALID, "vs_rrr_pack_u8_exact: ops.ldt must be vs_rrr_ldt(T)");
    VS_LAUNCH(exact_stats_kernel, (unsigned)ceil_div(d.C1 * ldt, 256), 256, 0, stream, sorted_idx, mean, std_clipped, (long long)d.T,
              (long long)d.C1, ldt, const_cast<float*>(out->isdT), const_cast<float*>(out->qT));
  }
  if (out->isd) {
    VS_REQUIRE(out->qh && out->isdmax, VS_ERR_INVALID, "vs_rrr_pack_u8_exact: isd, qh and isdmax go together");
    const long long Tq = round_up(d.T, 16);
    VS_CHECK_CUDA(cudaMemsetAsync(const_cast<float*>(out->isdmax), 0, (size_t)d.T * 4, (cudaStream_t)stream));
    VS_LAUNCH(exact_stats2_kernel, (unsigned)ceil_div(Tq * d.ldc, 256), 256, 0, stream, sorted_idx, mean, std_clipped, (long long)d.T, Tq,
              (long long)d.C1, (long long)d.ldc, const_cast<float*>(out->isd), const_cast<uint16_t*>(out->qh),
              reinterpret_cast<unsigned*>(const_cast<float*>(out->isdmax)));
  }
  return VS_OK;
}

// Fused R0 loader (pack_fused_kernel): ONE read of the frames per split.  compute_stats != 0 (train split): mean / std_clipped
// of the SELECTED frames sorted_idx[t] are computed and written (rows of the (Tf, C1) tables that are not selected are left
// untouched); compute_stats == 0: they are read.  Same outputs as vs_rrr_pack_u8_exact.
extern "C" int vs_rrr_pack_u8_fused(const uint8_t* frames, int64_t Tf, const int32_t* sorted_idx, double* mean, double* std_clipped,
                                    int compute_stats, vs_rrr_dims d, uint16_t* Xa, const vs_rrr_exact_ops* out, float* xl,
                                    int32_t* overflow_flag, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  VS_REQUIRE(d.mode == VS_RRR_MODE_EXACT || d.mode == VS_RRR_MODE_DENSE, VS_ERR_INVALID, "vs_rrr_pack_u8_fused: dims.mode must be an exact-operand mode");
  VS_REQUIRE(frames && sorted_idx && mean && std_clipped && out && xl && Tf >= d.T, VS_ERR_INVALID, "vs_rrr_pack_u8_fused: bad arguments");
  VS_REQUIRE(d.mode == VS_RRR_MODE_EXACT ? Xa != nullptr : out->Xc != nullptr, VS_ERR_INVALID,
             "vs_rrr_pack_u8_fused: the forward operand of the mode is missing (Xa for EXACT, ops.Xc for DENSE)");
  const size_t smem = (size_t)d.K * kPackRow;
  VS_REQUIRE(d.C1 % 4 == 0 && ((uintptr_t)frames & 3) == 0 && smem <= 200 * 1024, VS_ERR_UNSUPPORTED,
             "vs_rrr_pack_u8_fused: needs C1 %% 4 == 0, 4-byte aligned frames and K <= 1600 (use vs_rrr_colstats + vs_rrr_pack_u8_exact)");
  uint16_t* Xi = const_cast<uint16_t*>(out->Xi);
  uint16_t* Xc = d.mode == VS_RRR_MODE_DENSE ? const_cast<uint16_t*>(out->Xc) : nullptr;
  uint16_t* Xz = d.mode == VS_RRR_MODE_EXACT ? Xa : nullptr;
  dim3 grid((unsigned)ceil_div(d.C1, 128), (unsigned)d.T);
  VS_REQUIRE(grid.y <= 65535u, VS_ERR_UNSUPPORTED, "vs_rrr_pack_u8_fused: too many time bins");
  if (compute_stats) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(pack_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(pack_fused_kernel<true>, grid, 256, smem, stream, frames, sorted_idx, mean, std_clipped, (long long)Tf, (long long)d.K,
              (long long)d.T, (long long)d.C1, (long long)d.ldc, (long long)d.ldr, Xz, Xc, Xi, xl, overflow_flag);
  } else {
    VS_CHECK_CUDA(cudaFuncSetAttribute(pack_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(pack_fused_kernel<false>, grid, 256, smem, stream, frames, sorted_idx, mean, std_clipped, (long long)Tf, (long long)d.K,
              (long long)d.T, (long long)d.C1, (long long)d.ldc, (long long)d.ldr, Xz, Xc, Xi, xl, overflow_flag);
  }
  if (out->isdT) {
    const long long ldt = round_up(d.T, 4);
    VS_REQUIRE(out->ldt == ldt && out->qT, VS_ERR_INVALID, "vs_rrr_pack_u8_fused: ops.ldt must be vs_rrr_ldt(T); isdT and qT go together");
    VS_LAUNCH(exact_stats_kernel, (unsigned)ceil_div(d.C1 * ldt, 256), 256, 0, stream, sorted_idx, mean, std_clipped, (long long)d.T,
              (long long)d.C1, ldt, const_cast<float*>(out->isdT), const_cast<float*>(out->qT));
  }
  if (out->isd) {
    VS_REQUIRE(out->qh && out->isdmax, VS_ERR_INVALID, "vs_rrr_pack_u8_fused: isd, qh and isdmax go together");
    const long long Tq = round_up(d.T, 16);
    VS_CHECK_CUDA(cudaMemsetAsync(const_cast<float*>(out->isdmax), 0, (size_t)d.T * 4, (cudaStream_t)stream));
    VS_LAUNCH(exact_stats2_kernel, (unsigned)ceil_div(Tq * d.ldc, 256), 256, 0, stream, sorted_idx, mean, std_clipped, (long long)d.T, Tq,
              (long long)d.C1, (long long)d.ldc, const_cast<float*>(out->isd), const_cast<uint16_t*>(out->qh),
              reinterpret_cast<unsigned*>(const_cast<float*>(out->isdmax)));
  }
  return VS_OK;
}

// src/model/rrr.py:105-130 for a split packed in the dense-forward mode (vs_rrr_predict serves the other modes)
extern "C" int vs_rrr_predict_exact(vs_rrr_dims d, const vs_rrr_exact_ops* ops, const float* xl, const double* U, const double* V,
                                    const double* b, double* yhat, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  VS_REQUIRE(d.mode == VS_RRR_MODE_DENSE, VS_ERR_INVALID, "vs_rrr_predict_exact: dims.mode must be VS_RRR_MODE_DENSE");
  VS_REQUIRE(ops && xl && U && V && b && yhat, VS_ERR_INVALID, "vs_rrr_predict_exact: null pointer");
  VS_REQUIRE(workspace && workspace_bytes >= vs_rrr_workspace(d), VS_ERR_WORKSPACE, "vs_rrr_predict_exact: workspace too small");
  VS_REQUIRE(((uintptr_t)workspace & 1023) == 0, VS_ERR_INVALID, "vs_rrr_predict_exact: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = carve(d, workspace);
  int sf = 1;
  rc = dense_forward(d, exact_args(ops), w, U, V, false, st, &sf);
  if (rc) return rc;
  dim3 ge((unsigned)w.KB, (unsigned)d.T, (unsigned)ceil_div(d.N, 32));
  VS_LAUNCH((epi_d_kernel<true>), ge, 256, 0, st, w.Z, w.Npad, sf, (long long)(d.K * d.T * w.Npad), w.M1, w.ldm, (const float*)nullptr, (const float*)nullptr, xl, V, b, (long long)d.K,
            (long long)d.T, (long long)d.N, w.Npad, (int)d.r, (long long)d.ldr, (uint16_t*)nullptr, (double*)nullptr, (double*)nullptr, yhat, 0ll);
  return VS_OK;
}

extern "C" int vs_rrr_predict(vs_rrr_dims d, const uint16_t* Xa, const float* xl, const double* U, const double* V,
                              const double* b, double* yhat, int engine, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  VS_REQUIRE(Xa && xl && U && V && b && yhat, VS_ERR_INVALID, "vs_rrr_predict: null pointer");
  VS_REQUIRE(workspace && workspace_bytes >= vs_rrr_workspace(d), VS_ERR_WORKSPACE, "vs_rrr_predict: workspace too small");
  VS_REQUIRE(((uintptr_t)workspace & 1023) == 0, VS_ERR_INVALID, "vs_rrr_predict: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = carve(d, workspace);
  const long long KT = d.K * d.T;
  dim3 g0((unsigned)ceil_div(d.C1, 256 * kPrepC), (unsigned)d.N);
  VS_REQUIRE(d.mode != VS_RRR_MODE_DENSE, VS_ERR_INVALID, "vs_rrr_predict: dense-forward splits go through vs_rrr_predict_exact");
  const double uscale = d.mode == VS_RRR_MODE_EXACT ? kExactUScale : 1.0;
  float* u32 = nullptr; long long ldu = 0; unsigned* umax = nullptr;
  if (d.r == 3) {
    VS_LAUNCH(prep_u_kernel<3>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, (int)d.r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub,
              (double*)nullptr, uscale, u32, ldu, umax);
  } else if (d.r <= 4) {
    VS_LAUNCH(prep_u_kernel<4>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, (int)d.r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub,
              (double*)nullptr, uscale, u32, ldu, umax);
  } else {
    VS_LAUNCH(prep_u_kernel<kMaxR>, g0, 256, 0, st, U, (long long)d.N, w.Npad, (long long)d.C1, (int)d.r, d.planes, (int)d.fmt, (long long)d.ldc, w.Ub,
              (double*)nullptr, uscale, u32, ldu, umax);
  }
  int sf = 1;
  rc = gemm_f(d, Xa, w, engine, st, &sf);
  if (rc) return rc;
  dim3 ge((unsigned)w.KB, (unsigned)d.T, (unsigned)ceil_div(d.N, 32));
  if (d.r <= 4) {
    VS_LAUNCH((epi_f_kernel<true, 4, float>), ge, 256, 0, st, w.Z, w.ldz, sf, KT * w.ldz, nullptr, nullptr, xl, V, b, (long long)d.K, (long long)d.T,
              (long long)d.N, w.Npad, (int)d.r, d.planes, (int)d.fmt, (long long)d.ldr, nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, yhat, 0ll, 0, 1.0 / uscale);
  } else {
    VS_LAUNCH((epi_f_kernel<true, kMaxR, float>), ge, 256, 0, st, w.Z, w.ldz, sf, KT * w.ldz, nullptr, nullptr, xl, V, b, (long long)d.K, (long long)d.T,
              (long long)d.N, w.Npad, (int)d.r, d.planes, (int)d.fmt, (long long)d.ldr, nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, yhat, 0ll, 0, 1.0 / uscale);
  }
  return VS_OK;
}

extern "C" int vs_rrr_smooth_y(const float* counts, int64_t K, int64_t T, int64_t N, double sigma, const double* mean,
                               const double* std_clipped, float* y_out, void* stream) {
  VS_REQUIRE(counts && y_out && K > 0 && T > 0 && N > 0, VS_ERR_INVALID, "vs_rrr_smooth_y: bad arguments");
  VS_REQUIRE(sigma > 0.0 && (int)(4.0 * sigma + 0.5) <= 32, VS_ERR_UNSUPPORTED, "vs_rrr_smooth_y: sigma out of range");
  VS_REQUIRE((mean == nullptr) == (std_clipped == nullptr), VS_ERR_INVALID, "vs_rrr_smooth_y: mean and std go together");
  const long long n = K * T * N;
  VS_LAUNCH(smooth_y_kernel, (unsigned)ceil_div(n, 256), 256, 0, stream, counts, (long long)K, (long long)T, (long long)N, sigma, mean,
            std_clipped, y_out, (float*)nullptr);
  return VS_OK;
}

extern "C" int vs_rrr_smooth_y2(const float* counts, int64_t K, int64_t T, int64_t N, double sigma, const double* mean,
                                const double* std_clipped, float* y_out, float* y_lo_out, void* stream) {
  VS_REQUIRE(counts && y_out && K > 0 && T > 0 && N > 0, VS_ERR_INVALID, "vs_rrr_smooth_y2: bad arguments");
  VS_REQUIRE(sigma > 0.0 && (int)(4.0 * sigma + 0.5) <= 32, VS_ERR_UNSUPPORTED, "vs_rrr_smooth_y2: sigma out of range");
  VS_REQUIRE((mean == nullptr) == (std_clipped == nullptr), VS_ERR_INVALID, "vs_rrr_smooth_y2: mean and std go together");
  const long long n = K * T * N;
  VS_LAUNCH(smooth_y_kernel, (unsigned)ceil_div(n, 256), 256, 0, stream, counts, (long long)K, (long long)T, (long long)N, sigma, mean,
            std_clipped, y_out, y_lo_out);
  return VS_OK;
}

extern "C" int vs_colstats_f32(const float* x, int64_t K, int64_t cols, double* mean, double* std_clipped, void* stream) {
  VS_REQUIRE(x && mean && std_clipped && K > 0 && cols > 0, VS_ERR_INVALID, "vs_colstats_f32: bad arguments");
  VS_LAUNCH(colstats_f32_kernel, (unsigned)ceil_div(cols, 256), 256, 0, stream, x, (long long)K, (long long)cols, mean, std_clipped);
  return VS_OK;
}
