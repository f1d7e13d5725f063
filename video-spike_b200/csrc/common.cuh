// Shared helpers for libvs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/vs_b200.h"

namespace vs {

// ---- thread-local error string ------------------------------------------------
char* err_buf();
int set_err(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define VS_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return vs::set_err(VS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                 \
  } while (0)

#define VS_REQUIRE(cond, code, ...)                 \
  do {                                              \
    if (!(cond)) return vs::set_err(code, __VA_ARGS__); \
  } while (0)

// every kernel launch goes through this so the launch counter stays honest
#define VS_LAUNCH(kernel, grid, block, smem, stream, ...)                                     \
  do {                                                                                        \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                 \
    vs::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess)                                                                    \
      return vs::set_err(VS_ERR_CUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_e)); \
  } while (0)

// ---- optional per-kernel timing (bench.py roofline): event pairs on the launching stream -------
enum { PROF_GEMM_TC = 0, PROF_DW_ADAMW = 1, PROF_RRR_BWD = 2, PROF_RRR_FWD = 3, PROF_RRR_DV = 4, PROF_NUM_TAGS = 5 };
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- device helpers -------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit global accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_f4(const float4* p) {  // read-modify-write data: plain load
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace vs
