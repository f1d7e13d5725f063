// HBM-bound kernels of the `Linear` train step: frame loader, Poisson-NLL epilogue, AdamW, and
// the fused first-layer weight-gradient + AdamW update.  All are streaming kernels: 128-bit
// coalesced accesses, L1 bypass for touch-once data, enough loads in flight per SM to cover HBM
// latency, grids sized in multiples of the SM count.
#include <stdlib.h>
#include "common.cuh"
#include "adam.cuh"

namespace vs {

// ------------------------------------------------------------------ loader (L1/L2)
// src/loader/base.py:39,54: uint8 -> float32 WITHOUT scaling; flatten(1) is a view.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long n4,
                                                        long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(in) + i);
    st_stream_f4(reinterpret_cast<float4*>(out) + i,
                 make_float4((float)(w & 0xff), (float)((w >> 8) & 0xff), (float)((w >> 16) & 0xff), (float)(w >> 24)));
  }
  // ragged tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n - n4 * 4)) out[n4 * 4 + threadIdx.x] = (float)in[n4 * 4 + threadIdx.x];
}

__global__ void __launch_bounds__(256) u8_to_bf16_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out,
                                                         long long n8, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(in) + i);
    uint32_t o[4];
    const uint32_t src[2] = {w.x, w.y};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float lo = (float)((src[h] >> (16 * q)) & 0xff), hi = (float)((src[h] >> (16 * q + 8)) & 0xff);
        // 0..255 is exact in bf16: take the upper 16 bits of the fp32 pattern
        o[h * 2 + q] = (__float_as_uint(lo) >> 16) | (__float_as_uint(hi) & 0xffff0000u);
      }
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n - n8 * 8))
    out[n8 * 8 + threadIdx.x] = (uint16_t)(__float_as_uint((float)in[n8 * 8 + threadIdx.x]) >> 16);
}

// ------------------------------------------------------------------ trial windows (L0)
// out[trial][f][:] = frames[start[trial] + f][:] for f < F: F consecutive frames are contiguous in the session video,
// so every window is ONE contiguous copy of F*row bytes.  VEC = bytes per access (16, 4 or 1, chosen by the host from
// the alignment row guarantees).  Frames past the end of the video read as 0.
template <int VEC>
__global__ void __launch_bounds__(256) gather_windows_kernel(const uint8_t* __restrict__ frames, long long n_frames, long long row,
                                                             const long long* __restrict__ start, long long F,
                                                             uint8_t* __restrict__ out) {
  const long long trial = blockIdx.y;
  const long long s0 = start[trial];
  const long long bytes = F * row, units = bytes / VEC;
  const long long src0 = s0 * row, limit = n_frames * row;
  uint8_t* dst = out + trial * bytes;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
    const long long off = src0 + u * VEC;
    const bool in = off >= 0 && off + VEC <= limit;
    if constexpr (VEC == 16) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (in) v = ld_stream_u4(reinterpret_cast<const uint4*>(frames + off));
      *reinterpret_cast<uint4*>(dst + u * 16) = v;
    } else if constexpr (VEC == 4) {
      uint32_t v = 0;
      if (in) v = __ldg(reinterpret_cast<const uint32_t*>(frames + off));
      *reinterpret_cast<uint32_t*>(dst + u * 4) = v;
    } else {
      dst[u] = in ? __ldg(frames + off) : (uint8_t)0;
    }
  }
}

static unsigned stream_grid(long long work_items, int per_block) {
  long long blocks = ceil_div(work_items, per_block);
  const long long cap = (long long)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// ------------------------------------------------------------------ Poisson NLL (C1)
__global__ void __launch_bounds__(256) poisson_nll_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                          double* __restrict__ loss_sum, float* __restrict__ dx, long long n,
                                                          float inv_n) {
  __shared__ double wsum[8];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  double local = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 xv = ld_stream_f4(reinterpret_cast<const float4*>(x) + i);
    const float4 tv = ld_stream_f4(reinterpret_cast<const float4*>(t) + i);
    const float e0 = expf(xv.x), e1 = expf(xv.y), e2 = expf(xv.z), e3 = expf(xv.w);
    float s = (e0 - tv.x * xv.x) + (e1 - tv.y * xv.y);
    s += (e2 - tv.z * xv.z) + (e3 - tv.w * xv.w);
    local += (double)s;
    if (dx)
      st_stream_f4(reinterpret_cast<float4*>(dx) + i,
                   make_float4((e0 - tv.x) * inv_n, (e1 - tv.y) * inv_n, (e2 - tv.z) * inv_n, (e3 - tv.w) * inv_n));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    const float e = expf(x[i]);
    local += (double)(e - t[i] * x[i]);
    if (dx) dx[i] = (e - t[i]) * inv_n;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? wsum[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v);
  }
}

// ------------------------------------------------------------------ AdamW (O1)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, const AdamConsts c) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = ld_f4(reinterpret_cast<float4*>(p) + i);
    float4 mv = ld_f4(reinterpret_cast<float4*>(m) + i);
    float4 vv = ld_f4(reinterpret_cast<float4*>(v) + i);
    const float4 gv = ld_stream_f4(reinterpret_cast<const float4*>(g) + i);
    adamw_elem(pv.x, mv.x, vv.x, gv.x, c);
    adamw_elem(pv.y, mv.y, vv.y, gv.y, c);
    adamw_elem(pv.z, mv.z, vv.z, gv.z, c);
    adamw_elem(pv.w, mv.w, vv.w, gv.w, c);
    st_stream_f4(reinterpret_cast<float4*>(p) + i, pv);
    st_stream_f4(reinterpret_cast<float4*>(m) + i, mv);
    st_stream_f4(reinterpret_cast<float4*>(v) + i, vv);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    adamw_elem(p[i], m[i], v[i], g[i], c);
  }
}

// ------------------------------------------------------------------ fused dW + AdamW (G1 + O1)
// W, m, v are (out, in) row-major.  A CTA owns a strip of 4*256 input columns and walks all
// `out` rows; each thread keeps x[0..BT) for its 4 columns in registers (frames are read ONCE),
// forms g = sum_b dy[b,o]*x[b,i] with BT FMAs per weight and applies AdamW on the spot.  HBM
// traffic is exactly read p,m,v + write p,m,v (24 B/weight) + the frames; dW never exists.
template <int BT, bool kU8, int UNROLL>
__global__ void __launch_bounds__(256, ((BT <= 16 && UNROLL <= 2) ? 2 : 1))
dw_adamw_kernel(const float* __restrict__ dy, const float* __restrict__ xf, const uint8_t* __restrict__ xu,
                float* __restrict__ W, float* __restrict__ M, float* __restrict__ V, float* __restrict__ bias,
                float* __restrict__ mb, float* __restrict__ vb, int batch, long long in_dim, int out_dim, const AdamConsts c) {
  extern __shared__ float dys[];  // [out][BT], zero padded past batch
  for (int e = threadIdx.x; e < out_dim * BT; e += 256) {
    const int o = e / BT, b = e % BT;
    dys[e] = b < batch ? dy[(long long)b * out_dim + o] : 0.f;
  }
  const long long col = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  const bool active = col < in_dim;  // in_dim % 4 == 0 is required by the host wrapper
  float x[BT][4];
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    if (active && b < batch) {
      if constexpr (kU8) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(xu + (long long)b * in_dim + col));
        x[b][0] = (float)(w & 0xff); x[b][1] = (float)((w >> 8) & 0xff);
        x[b][2] = (float)((w >> 16) & 0xff); x[b][3] = (float)(w >> 24);
      } else {
        const float4 w = ld_stream_f4(reinterpret_cast<const float4*>(xf + (long long)b * in_dim + col));
        x[b][0] = w.x; x[b][1] = w.y; x[b][2] = w.z; x[b][3] = w.w;
      }
    } else {
      x[b][0] = x[b][1] = x[b][2] = x[b][3] = 0.f;
    }
  }
  __syncthreads();
  // bias of the layer: dbias[o] = sum_b dy[b,o] -> AdamW, by the first column strip
  if (blockIdx.x == 0 && bias != nullptr) {
    for (int o = threadIdx.x; o < out_dim; o += 256) {
      float gb = 0.f;
#pragma unroll
      for (int b = 0; b < BT; ++b) gb += dys[o * BT + b];
      float p = bias[o], m = mb[o], v = vb[o];
      adamw_elem(p, m, v, gb, c);
      bias[o] = p; mb[o] = m; vb[o] = v;
    }
  }
  if (!active) return;
  for (int o0 = 0; o0 < out_dim; o0 += UNROLL) {
    float4 pv[UNROLL], mv[UNROLL], vv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (o0 + u < out_dim) {
        const long long off = (long long)(o0 + u) * in_dim + col;
        pv[u] = ld_f4(reinterpret_cast<float4*>(W + off));
        mv[u] = ld_f4(reinterpret_cast<float4*>(M + off));
        vv[u] = ld_f4(reinterpret_cast<float4*>(V + off));
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (o0 + u < out_dim) {
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        const float4* d4 = reinterpret_cast<const float4*>(dys + (o0 + u) * BT);
#pragma unroll
        for (int b4 = 0; b4 < BT / 4; ++b4) {
          const float4 d = d4[b4];  // broadcast read
          const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int b = b4 * 4 + q;
            g0 = fmaf(dd[q], x[b][0], g0); g1 = fmaf(dd[q], x[b][1], g1);
            g2 = fmaf(dd[q], x[b][2], g2); g3 = fmaf(dd[q], x[b][3], g3);
          }
        }
        adamw_elem(pv[u].x, mv[u].x, vv[u].x, g0, c);
        adamw_elem(pv[u].y, mv[u].y, vv[u].y, g1, c);
        adamw_elem(pv[u].z, mv[u].z, vv[u].z, g2, c);
        adamw_elem(pv[u].w, mv[u].w, vv[u].w, g3, c);
        const long long off = (long long)(o0 + u) * in_dim + col;
        st_stream_f4(reinterpret_cast<float4*>(W + off), pv[u]);
        st_stream_f4(reinterpret_cast<float4*>(M + off), mv[u]);
        st_stream_f4(reinterpret_cast<float4*>(V + off), vv[u]);
      }
    }
  }
}

// ------------------------------------------------------------------ fused dW + AdamW, bulk-copy ring (G1 + O1)
// Same arithmetic as dw_adamw_kernel, different data movement.  The register-only version keeps ~49 KB of loads in
// flight per SM (2 CTAs x 256 threads x 2 rows x 48 B) and stops at 0.79 of the copy peak; here ONE producer thread
// streams the p/m/v rows of the CTA's column strip into a shared-memory ring with cp.async.bulk (the 1-D bulk-copy path
// of the TMA unit, completion on an mbarrier), 12 stages x 12 KB = 144 KB in flight per SM independent of register
// pressure, and 256 consumer threads do the math and write the results straight back with 128-bit stores.
// Persistent: CTA c walks strips c, c + gridDim.x, ...; the uint8 frame rows of a strip (BT x 1 KB) arrive through
// their own double-buffered bulk copies.  Requires in_dim % 1024 == 0 and 16-byte aligned buffers (the host wrapper
// falls back to dw_adamw_kernel otherwise).
constexpr int kRingStages = 12;
constexpr int kStripCols = 1024;

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ring_mbar_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < 20000; ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000) : "memory");
    if (ok) return;
  }
  printf("vs_b200 dw_adamw ring: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void ring_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ring_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int BT>
__global__ void __launch_bounds__(288, 1)
dw_adamw_ring_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ xu, float* __restrict__ W, float* __restrict__ M,
                     float* __restrict__ V, float* __restrict__ bias, float* __restrict__ mb, float* __restrict__ vb, int batch,
                     long long in_dim, int out_dim, int n_strips, const AdamConsts c) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  // layout: barriers (256 B) | dys [out][BT] fp32 | xs [2][BT][1024] u8 | ring [stages][3][1024] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_smem);
  float* dys = reinterpret_cast<float*>(ring_smem + 256);
  uint8_t* xs = ring_smem + 256 + (size_t)out_dim * BT * sizeof(float);
  float* ring = reinterpret_cast<float*>(xs + 2 * BT * kStripCols);
  const uint32_t bar_full0 = sm_u32(bars), bar_empty0 = sm_u32(bars + kRingStages);
  const uint32_t bar_xfull0 = sm_u32(bars + 2 * kRingStages), bar_xempty0 = sm_u32(bars + 2 * kRingStages + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kRingStages; ++s) { ring_mbar_init(bar_full0 + 8u * s, 1); ring_mbar_init(bar_empty0 + 8u * s, 8); }
    for (int b = 0; b < 2; ++b) { ring_mbar_init(bar_xfull0 + 8u * b, 1); ring_mbar_init(bar_xempty0 + 8u * b, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < out_dim * BT; e += 288) {
    const int o = e / BT, b = e % BT;
    dys[e] = b < batch ? dy[(long long)b * out_dim + o] : 0.f;
  }
  __syncthreads();
  if (warp == 8) {
    // ===== producer: one thread issues every bulk copy =====
    if (tid == 256) {
      int it = 0, xs_it = 0;
      for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x, ++xs_it) {
        const long long col0 = (long long)strip * kStripCols;
        const int xb = xs_it & 1;
        ring_mbar_wait(bar_xempty0 + 8u * xb, ((uint32_t)(xs_it >> 1) & 1u) ^ 1u);
        ring_expect_tx(bar_xfull0 + 8u * xb, (uint32_t)(batch * kStripCols));
        for (int b = 0; b < batch; ++b)
          bulk_g2s(sm_u32(xs + ((size_t)xb * BT + b) * kStripCols), xu + (long long)b * in_dim + col0, kStripCols, bar_xfull0 + 8u * xb);
        for (int o = 0; o < out_dim; ++o, ++it) {
          const int s = it % kRingStages;
          ring_mbar_wait(bar_empty0 + 8u * s, ((uint32_t)(it / kRingStages) & 1u) ^ 1u);
          const uint32_t full = bar_full0 + 8u * s;
          ring_expect_tx(full, 3u * kStripCols * 4u);
          const long long off = (long long)o * in_dim + col0;
          const uint32_t dst = sm_u32(ring + (size_t)s * 3 * kStripCols);
          bulk_g2s(dst, W + off, kStripCols * 4, full);
          bulk_g2s(dst + kStripCols * 4, M + off, kStripCols * 4, full);
          bulk_g2s(dst + 2 * kStripCols * 4, V + off, kStripCols * 4, full);
        }
      }
    }
    return;
  }
  // ===== consumers: 256 threads, 4 columns each =====
  const int lane = tid & 31;
  int it = 0, xs_it = 0;
  for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x, ++xs_it) {
    const long long col = (long long)strip * kStripCols + tid * 4;
    const int xb = xs_it & 1;
    ring_mbar_wait(bar_xfull0 + 8u * xb, (uint32_t)(xs_it >> 1) & 1u);
    float x[BT][4];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      uint32_t w = 0;
      if (b < batch) w = *reinterpret_cast<const uint32_t*>(xs + ((size_t)xb * BT + b) * kStripCols + tid * 4);
      x[b][0] = (float)(w & 0xff); x[b][1] = (float)((w >> 8) & 0xff);
      x[b][2] = (float)((w >> 16) & 0xff); x[b][3] = (float)(w >> 24);
    }
    __syncwarp();
    if (lane == 0) ring_arrive(bar_xempty0 + 8u * xb);
    // bias of the layer: dbias[o] = sum_b dy[b,o] -> AdamW, by whoever owns strip 0
    if (strip == 0 && bias != nullptr) {
      for (int o = tid; o < out_dim; o += 256) {
        float gb = 0.f;
#pragma unroll
        for (int b = 0; b < BT; ++b) gb += dys[o * BT + b];
        float p = bias[o], m = mb[o], v = vb[o];
        adamw_elem(p, m, v, gb, c);
        bias[o] = p; mb[o] = m; vb[o] = v;
      }
    }
    for (int o = 0; o < out_dim; ++o, ++it) {
      const int s = it % kRingStages;
      ring_mbar_wait(bar_full0 + 8u * s, (uint32_t)(it / kRingStages) & 1u);
      const float* st = ring + (size_t)s * 3 * kStripCols + tid * 4;
      float4 pv = *reinterpret_cast<const float4*>(st);
      float4 mv = *reinterpret_cast<const float4*>(st + kStripCols);
      float4 vv = *reinterpret_cast<const float4*>(st + 2 * kStripCols);
      __syncwarp();
      if (lane == 0) ring_arrive(bar_empty0 + 8u * s);      // the slot's data is in registers: hand it back at once
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
      const float4* d4 = reinterpret_cast<const float4*>(dys + o * BT);
#pragma unroll
      for (int b4 = 0; b4 < BT / 4; ++b4) {
        const float4 d = d4[b4];  // broadcast read
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int b = b4 * 4 + q;
          g0 = fmaf(dd[q], x[b][0], g0); g1 = fmaf(dd[q], x[b][1], g1);
          g2 = fmaf(dd[q], x[b][2], g2); g3 = fmaf(dd[q], x[b][3], g3);
        }
      }
      adamw_elem(pv.x, mv.x, vv.x, g0, c); adamw_elem(pv.y, mv.y, vv.y, g1, c);
      adamw_elem(pv.z, mv.z, vv.z, g2, c); adamw_elem(pv.w, mv.w, vv.w, g3, c);
      const long long off = (long long)o * in_dim + col;
      st_stream_f4(reinterpret_cast<float4*>(W + off), pv);
      st_stream_f4(reinterpret_cast<float4*>(M + off), mv);
      st_stream_f4(reinterpret_cast<float4*>(V + off), vv);
    }
  }
}

template <int BT>
static int launch_dw_adamw_ring(const float* dy, const uint8_t* xu, float* W, float* m, float* v, float* bias, float* mb, float* vb,
                                int batch, long long in_dim, int out_dim, const AdamConsts& c, cudaStream_t st) {
  const size_t smem = 256 + (size_t)out_dim * BT * sizeof(float) + 2 * (size_t)BT * kStripCols + (size_t)kRingStages * 3 * kStripCols * 4;
  const int n_strips = (int)(in_dim / kStripCols);
  const unsigned grid = (unsigned)(n_strips < kNumSMs ? n_strips : kNumSMs);
  VS_CHECK_CUDA(cudaFuncSetAttribute(dw_adamw_ring_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH((dw_adamw_ring_kernel<BT>), grid, 288, smem, st, dy, xu, W, m, v, bias, mb, vb, batch, in_dim, out_dim, n_strips, c);
  return VS_OK;
}

// ring engine applies when the frames are uint8, the strip grid is exact and everything is 16-byte aligned
static bool dw_ring_ok(const uint8_t* xu, const float* W, const float* m, const float* v, int batch, long long in_dim, int out_dim) {
  static int enabled = -1;
  // measured on B200 (profiles/r01_linear_dw_variants.txt): 2.45 ms against 2.34 ms for the register-only kernel, so
  // latency hiding is not what separates that kernel from the copy peak; the ring stays selectable (VS_DW_RING=1) only
  if (enabled < 0) { const char* e = getenv("VS_DW_RING"); enabled = e ? atoi(e) : 0; }
  if (!enabled || !xu) return false;
  if (in_dim % kStripCols != 0 || in_dim < kStripCols) return false;
  if ((((uintptr_t)xu | (uintptr_t)W | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return false;
  const int BT = batch <= 8 ? 8 : (batch <= 16 ? 16 : 32);
  const size_t smem = 256 + (size_t)out_dim * BT * 4 + 2 * (size_t)BT * kStripCols + (size_t)kRingStages * 3 * kStripCols * 4;
  return smem <= 227 * 1024;
}

// ------------------------------------------------------------------ small helpers for backward
// dy_masked = dy * (y > 0)   (threshold backward of ReLU)
__global__ void relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// dbias[o] = sum_b dy[b,o]
__global__ void colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, long long batch, long long out_dim) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out_dim) return;
  float s = 0.f;
  for (long long b = 0; b < batch; ++b) s += dy[b * out_dim + o];
  db[o] = s;
}

// y[b,o] = act(y[b,o] + bias[o]) in place (row-parallel first layer: the pre-activation arrives all-reduced)
__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long n, long long out_dim, int relu) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = y[i] + (bias ? bias[i % out_dim] : 0.f);
  y[i] = relu ? fmaxf(v, 0.f) : v;
}
int launch_bias_act(float* y, const float* bias, long long batch, long long out_dim, int relu, cudaStream_t st) {
  const long long n = batch * out_dim;
  VS_LAUNCH(bias_act_kernel, (unsigned)ceil_div(n, 256), 256, 0, st, y, bias, n, out_dim, relu);
  return VS_OK;
}

int launch_relu_mask(const float* dy, const float* y, float* out, long long n, cudaStream_t st) {
  VS_LAUNCH(relu_mask_kernel, (unsigned)ceil_div(n, 256), 256, 0, st, dy, y, out, n);
  return VS_OK;
}
int launch_colsum(const float* dy, float* db, long long batch, long long out_dim, cudaStream_t st) {
  VS_LAUNCH(colsum_kernel, (unsigned)ceil_div(out_dim, 128), 128, 0, st, dy, db, batch, out_dim);
  return VS_OK;
}

// rows of W in flight per thread: 2 keeps two CTAs per SM resident, 8 trades occupancy for 96 KB of loads in flight
// per SM (measured on B200: see DESIGN.md "Linear step")
static int dw_unroll() {
  static int u = -1;
  if (u < 0) {
    const char* e = getenv("VS_DW_UNROLL");
    u = e ? atoi(e) : 2;
    if (u != 2 && u != 4 && u != 8) u = 2;
  }
  return u;
}

template <int BT, bool kU8, int UNROLL>
static int launch_dw_adamw_u(const float* dy, const float* xf, const uint8_t* xu, float* W, float* m, float* v, float* bias,
                             float* mb, float* vb, int batch, long long in_dim, int out_dim, const AdamConsts& c, cudaStream_t st) {
  const size_t smem = (size_t)out_dim * BT * sizeof(float);
  const unsigned grid = (unsigned)ceil_div(in_dim, 1024);
  VS_CHECK_CUDA(cudaFuncSetAttribute(dw_adamw_kernel<BT, kU8, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH((dw_adamw_kernel<BT, kU8, UNROLL>), grid, 256, smem, st, dy, xf, xu, W, m, v, bias, mb, vb, batch, in_dim, out_dim, c);
  return VS_OK;
}

template <int BT>
static int launch_dw_adamw(const float* dy, const float* xf, const uint8_t* xu, float* W, float* m, float* v, float* bias,
                           float* mb, float* vb, int batch, long long in_dim, int out_dim, const AdamConsts& c, cudaStream_t st) {
  const int u = dw_unroll();
#define VS_DW_CASE(U)                                                                                                   \
  if (u == U) return xu ? launch_dw_adamw_u<BT, true, U>(dy, xf, xu, W, m, v, bias, mb, vb, batch, in_dim, out_dim, c, st) \
                        : launch_dw_adamw_u<BT, false, U>(dy, xf, xu, W, m, v, bias, mb, vb, batch, in_dim, out_dim, c, st);
  VS_DW_CASE(2)
  VS_DW_CASE(4)
  VS_DW_CASE(8)
#undef VS_DW_CASE
  return VS_ERR_INVALID;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_u8_to_f32(const uint8_t* frames, float* out, int64_t n, void* stream) {
  VS_REQUIRE(frames && out && n >= 0, VS_ERR_INVALID, "vs_u8_to_f32: null pointer or negative size");
  if (n == 0) return VS_OK;
  VS_REQUIRE(((uintptr_t)frames & 3) == 0 && ((uintptr_t)out & 15) == 0, VS_ERR_INVALID, "vs_u8_to_f32: misaligned buffers");
  const long long n4 = n / 4;
  VS_LAUNCH(u8_to_f32_kernel, stream_grid(n4, 256 * 4), 256, 0, stream, frames, out, n4, (long long)n);
  return VS_OK;
}

extern "C" int vs_u8_to_bf16(const uint8_t* frames, uint16_t* out, int64_t n, void* stream) {
  VS_REQUIRE(frames && out && n >= 0, VS_ERR_INVALID, "vs_u8_to_bf16: null pointer or negative size");
  if (n == 0) return VS_OK;
  VS_REQUIRE(((uintptr_t)frames & 7) == 0 && ((uintptr_t)out & 15) == 0, VS_ERR_INVALID, "vs_u8_to_bf16: misaligned buffers");
  const long long n8 = n / 8;
  VS_LAUNCH(u8_to_bf16_kernel, stream_grid(n8, 256 * 4), 256, 0, stream, frames, out, n8, (long long)n);
  return VS_OK;
}

extern "C" int vs_h2d_select_frames(const uint8_t* host_frames, int64_t K, int64_t Tf, int64_t row_bytes, const int32_t* idx_host,
                                    int64_t T, uint8_t* dev_out, void* stream) {
  VS_REQUIRE(host_frames && idx_host && dev_out && K > 0 && Tf > 0 && row_bytes > 0 && T > 0 && T <= Tf, VS_ERR_INVALID,
             "vs_h2d_select_frames: bad arguments");
  for (int64_t t = 0; t < T; ++t)
    VS_REQUIRE(idx_host[t] >= 0 && idx_host[t] < Tf && (t == 0 || idx_host[t] > idx_host[t - 1]), VS_ERR_INVALID,
               "vs_h2d_select_frames: indices must be strictly increasing and inside [0, %lld)", (long long)Tf);
  int64_t t = 0;
  while (t < T) {                       // one 2-D copy per run of consecutive frame indices: K rows of (run * row_bytes) bytes
    int64_t e = t + 1;
    while (e < T && idx_host[e] == idx_host[e - 1] + 1) ++e;
    VS_CHECK_CUDA(cudaMemcpy2DAsync(dev_out + t * row_bytes, (size_t)(T * row_bytes), host_frames + (int64_t)idx_host[t] * row_bytes,
                                    (size_t)(Tf * row_bytes), (size_t)((e - t) * row_bytes), (size_t)K, cudaMemcpyHostToDevice,
                                    (cudaStream_t)stream));
    t = e;
  }
  return VS_OK;
}

extern "C" int vs_gather_windows(const uint8_t* frames, int64_t n_frames, int64_t row_bytes, const int64_t* start_idx,
                                 int64_t n_trials, int64_t frames_per_trial, uint8_t* out, void* stream) {
  VS_REQUIRE(frames && start_idx && out && n_frames > 0 && row_bytes > 0 && frames_per_trial > 0 && n_trials >= 0, VS_ERR_INVALID,
             "vs_gather_windows: bad arguments");
  if (n_trials == 0) return VS_OK;
  VS_REQUIRE(n_trials <= 65535, VS_ERR_UNSUPPORTED, "vs_gather_windows: more than 65535 trials per call");
  const long long bytes = frames_per_trial * row_bytes;
  const uintptr_t al = (uintptr_t)frames | (uintptr_t)out;
  if (row_bytes % 16 == 0 && (al & 15) == 0) {
    dim3 grid(stream_grid(bytes / 16, 256 * 4) > 64 ? 64 : stream_grid(bytes / 16, 256 * 4), (unsigned)n_trials);
    VS_LAUNCH(gather_windows_kernel<16>, grid, 256, 0, stream, frames, (long long)n_frames, (long long)row_bytes,
              reinterpret_cast<const long long*>(start_idx), (long long)frames_per_trial, out);
  } else if (row_bytes % 4 == 0 && (al & 3) == 0) {
    dim3 grid(stream_grid(bytes / 4, 256 * 4) > 64 ? 64 : stream_grid(bytes / 4, 256 * 4), (unsigned)n_trials);
    VS_LAUNCH(gather_windows_kernel<4>, grid, 256, 0, stream, frames, (long long)n_frames, (long long)row_bytes,
              reinterpret_cast<const long long*>(start_idx), (long long)frames_per_trial, out);
  } else {
    dim3 grid(stream_grid(bytes, 256 * 4) > 64 ? 64 : stream_grid(bytes, 256 * 4), (unsigned)n_trials);
    VS_LAUNCH(gather_windows_kernel<1>, grid, 256, 0, stream, frames, (long long)n_frames, (long long)row_bytes,
              reinterpret_cast<const long long*>(start_idx), (long long)frames_per_trial, out);
  }
  return VS_OK;
}

extern "C" int vs_poisson_nll(const float* logits, const float* target, double* loss_sum, float* dlogits, int64_t n,
                              void* stream) {
  VS_REQUIRE(logits && target && loss_sum && n > 0, VS_ERR_INVALID, "vs_poisson_nll: null pointer or empty input");
  VS_REQUIRE(((uintptr_t)logits & 15) == 0 && ((uintptr_t)target & 15) == 0 && ((uintptr_t)dlogits & 15) == 0, VS_ERR_INVALID,
             "vs_poisson_nll: buffers must be 16-byte aligned");
  VS_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), (cudaStream_t)stream));
  VS_LAUNCH(poisson_nll_kernel, stream_grid(n / 4 + 1, 256 * 2), 256, 0, stream, logits, target, loss_sum, dlogits,
            (long long)n, (float)(1.0 / (double)n));
  return VS_OK;
}

extern "C" int vs_adamw(float* p, const float* g, float* m, float* v, int64_t n, vs_adamw_hyper h, void* stream) {
  VS_REQUIRE(p && g && m && v && n > 0, VS_ERR_INVALID, "vs_adamw: null pointer or empty tensor");
  VS_REQUIRE(h.step >= 1, VS_ERR_INVALID, "vs_adamw: step must be >= 1");
  VS_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, VS_ERR_INVALID,
             "vs_adamw: buffers must be 16-byte aligned");
  VS_LAUNCH(adamw_kernel, stream_grid(n / 4 + 1, 256 * 2), 256, 0, stream, p, g, m, v, (long long)n, make_consts(h));
  return VS_OK;
}

extern "C" int vs_dw_adamw_fused(const float* dy, const float* x_f32, const uint8_t* x_u8, float* W, float* m, float* v,
                                 float* bias, float* mb, float* vb, int64_t batch, int64_t in_dim, int64_t out_dim,
                                 vs_adamw_hyper h, void* stream) {
  VS_REQUIRE(dy && (x_f32 || x_u8) && W && m && v, VS_ERR_INVALID, "vs_dw_adamw_fused: null pointer");
  VS_REQUIRE(!bias || (mb && vb), VS_ERR_INVALID, "vs_dw_adamw_fused: bias needs its moment buffers");
  VS_REQUIRE(h.step >= 1, VS_ERR_INVALID, "vs_dw_adamw_fused: step must be >= 1");
  VS_REQUIRE(batch >= 1 && batch <= 32, VS_ERR_UNSUPPORTED, "vs_dw_adamw_fused: batch %lld outside 1..32", (long long)batch);
  VS_REQUIRE(in_dim % 4 == 0, VS_ERR_UNSUPPORTED, "vs_dw_adamw_fused: in_dim must be a multiple of 4");
  VS_REQUIRE(out_dim >= 1 && out_dim * 32 * 4 <= 200 * 1024, VS_ERR_UNSUPPORTED, "vs_dw_adamw_fused: out_dim too large");
  VS_REQUIRE((((uintptr_t)W | (uintptr_t)m | (uintptr_t)v | (uintptr_t)x_f32) & 15) == 0 && ((uintptr_t)x_u8 & 3) == 0,
             VS_ERR_INVALID, "vs_dw_adamw_fused: misaligned buffers");
  const AdamConsts c = make_consts(h);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PROF_DW_ADAMW, st);
  int rc;
  if (dw_ring_ok(x_f32 ? nullptr : x_u8, W, m, v, (int)batch, in_dim, (int)out_dim)) {
    if (batch <= 8) rc = launch_dw_adamw_ring<8>(dy, x_u8, W, m, v, bias, mb, vb, (int)batch, in_dim, (int)out_dim, c, st);
    else if (batch <= 16) rc = launch_dw_adamw_ring<16>(dy, x_u8, W, m, v, bias, mb, vb, (int)batch, in_dim, (int)out_dim, c, st);
    else rc = launch_dw_adamw_ring<32>(dy, x_u8, W, m, v, bias, mb, vb, (int)batch, in_dim, (int)out_dim, c, st);
    prof_end(PROF_DW_ADAMW, st);
    return rc;
  }
  if (batch <= 8) rc = launch_dw_adamw<8>(dy, x_f32, x_u8, W, m, v, bias, mb, vb, (int)batch, in_dim, (int)out_dim, c, st);
  else if (batch <= 16) rc = launch_dw_adamw<16>(dy, x_f32, x_u8, W, m, v, bias, mb, vb, (int)batch, in_dim, (int)out_dim, c, st);
  else rc = launch_dw_adamw<32>(dy, x_f32, x_u8, W, m, v, bias, mb, vb, (int)batch, in_dim, (int)out_dim, c, st);
  prof_end(PROF_DW_ADAMW, st);
  return rc;
}
