// TN GEMM on the 5th-gen tensor cores:  C[M,N] (fp32) = sum over passes of A_pa[M,K] * B_pb[N,K]^T
//
//   * operands bf16 (kind::f16) or fp32 read as tf32 (kind::tf32), both K-major, fetched by TMA
//     (cp.async.bulk.tensor.3d, SWIZZLE_128B) into a multi-stage shared-memory ring
//   * one elected thread issues tcgen05.mma (UMMA 128 x N x 32B), accumulators live in TMEM
//   * 4 epilogue warps read TMEM with tcgen05.ld.32x32b and store the fp32 tile
//   * split-K: gridDim.y splits write partial tiles at C + split * split_stride
//   * "planes": an operand may be stored as several bf16 residual planes (X = X0 + X1 + X2);
//     the pass list (pa[i], pb[i]) names which plane products are accumulated
//
// Used for: first Linear layer forward (W0 x frames, split-K over pixels; src/model/linear.py:26),
// the two RRR contractions (src/model/rrr.py:113 and its autograd), and vs_gemm_tn.
#include <cuda.h>
#include "common.cuh"
#include "gemm.h"

namespace vs {
namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int KB_BYTES = 128;    // one 128B swizzle atom along K per k-block
constexpr int A_TILE_BYTES = BM * KB_BYTES;
constexpr int NUM_THREADS = 192; // warp0 TMA, warp1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int CTRL_BYTES = 1024; // barriers + tmem pointer

struct KParams {
  int M, N;         // logical extent of C
  int BN;           // tile width (multiple of 16, <= kMaxBN)
  int tb_rows;      // rows per TMA box of B
  int n_tb;         // TMA boxes per B tile
  int num_kb;       // k-blocks per pass
  int kb_per_split; // k-blocks handled by one split
  int n_pass;
  int pa[kMaxPass], pb[kMaxPass];
  int stages;
  int kb_group;     // consecutive k-blocks per pipeline stage (1, or 4 for skinny tiles: longer contiguous reads per operand row)
  int tmem_cols;
  int f16;          // 16-bit operand format: 0 = bf16, 1 = IEEE half
  float* C;
  long long ldc, split_stride;
  // tail-wave balancing: CTAs with blockIdx.x >= tail_cta0 work on tile tail_cta0 + u / tail_splits,
  // k-range u % tail_splits, and write dense (BM x BN) partial tiles to tail_ws
  int tail_cta0, tail_splits, tail_kb_per_split;
  float* tail_ws;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  // no suspend-time hint: with one, ptxas emits NANOSLEEP.SYNCS after a failed check and the wake-up latency (~0.4 us)
  // dominates pipelines whose stages hold only ~0.1 us of tensor-core work (profiles/r01_ncu_dense_bwd.txt)
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 1024; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (global_timer_ns() - t0 > 8000000000ull) break;   // 8 s
  }
  printf("vs_b200 gemm_tn: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
// Spinning wait for the single-thread roles (TMA producer, MMA issuer): try_wait parks the thread (NANOSLEEP.SYNCS) after a
// failed check and its wake-up latency lands on the critical path of every stage hand-off; test_wait never sleeps.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint64_t t0 = 0;
  for (int outer = 0;; ++outer) {
#pragma unroll 1
    for (int i = 0; i < 4096; ++i) {
      uint32_t ok;
      asm volatile(
          "{\n\t"
          ".reg .pred P1;\n\t"
          "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, P1;\n\t"
          "}"
          : "=r"(ok)
          : "r"(bar), "r"(parity)
          : "memory");
      if (ok) return;
    }
    if (outer == 0) t0 = global_timer_ns();
    else if (global_timer_ns() - t0 > 8000000000ull) break;   // 8 s
  }
  printf("vs_b200 gemm_tn: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kTF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// start address >> 4 in [0,14), LBO unused (single atom along K), SBO = 8 rows * 128 B = 1024 B
// in [32,46), version 1 in [46,48), layout type SWIZZLE_128B (= 2) in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// UMMA instruction descriptor: D fp32 (c_format 1 @ [4,6)), A/B format @ [7,10)/[10,13)
// (f16 = 0, bf16 = 1, tf32 = 2), both K-major (bits 15,16 = 0), N>>3 @ [17,23), M>>4 @ [24,29).
__device__ __forceinline__ uint32_t make_idesc(bool tf32, bool f16, int n) {
  uint32_t fmt = tf32 ? 2u : (f16 ? 0u : 1u);
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ kernel
template <bool kTF32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t kb_bytes = A_TILE_BYTES + (uint32_t)p.BN * KB_BYTES;   // one k-block: A tile + B tile
  const uint32_t stage_bytes = kb_bytes * (uint32_t)p.kb_group;         // a stage holds kb_group consecutive k-blocks
  const uint32_t tiles0 = base + CTRL_BYTES;
  const uint32_t bar_full0 = base;                 // stages x 8 B
  const uint32_t bar_empty0 = base + 8u * 16;      // stages x 8 B (stages <= 16)
  const uint32_t bar_tmem = base + 8u * 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 8 * 33);

  const int n_tiles = (p.N + p.BN - 1) / p.BN;
  int tile = blockIdx.x, split = blockIdx.y, kbps = p.kb_per_split;
  const bool is_tail = p.tail_splits > 1 && (int)blockIdx.x >= p.tail_cta0;
  if (is_tail) {
    const int u = blockIdx.x - p.tail_cta0;
    tile = p.tail_cta0 + u / p.tail_splits;
    split = u % p.tail_splits;
    kbps = p.tail_kb_per_split;
  }
  const int m_tile = tile / n_tiles, n_tile = tile % n_tiles;
  const int kb0 = split * kbps;
  const int kb1 = min(p.num_kb, kb0 + kbps);
  const int groups = kb1 > kb0 ? (kb1 - kb0 + p.kb_group - 1) / p.kb_group : 0;
  const int iters = groups * p.n_pass;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full0 + 8u * s, 1);
      mbar_init(bar_empty0 + 8u * s, 1);
    }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 8u * 33), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Both issuing roles run with the whole warp converged and elect one lane per instruction group (see gemm_tn_pair_kernel):
  // no ELECT / BRA.U.ANY retry loops around the uniform-datapath instructions, no integer division per stage.  This matters
  // most for the first Linear layer (N = 16: ~30 clk of tensor work per stage).
  if (warp == 0) {
    // ===== TMA producer =====
    const int elems_per_kb = kTF32 ? 32 : 64;
    int s = 0;
    uint32_t ph = 0;
    for (int ps = 0; ps < p.n_pass; ++ps) {
      const int pa = p.pa[ps], pb = p.pb[ps];
      for (int kb = kb0; kb < kb1; kb += p.kb_group) {
        mbar_wait(bar_empty0 + 8u * s, ph ^ 1u);
        if (elect_one()) {
          const uint32_t full = bar_full0 + 8u * s;
          mbar_arrive_expect_tx(full, stage_bytes);
          // consecutive k-blocks of the same rows back to back: each operand row is read in kb_group * 128 contiguous bytes
          // (k-blocks past the end of the split are zero-filled by TMA and add nothing)
          for (int g = 0; g < p.kb_group; ++g) {
            const uint32_t a_dst = tiles0 + s * stage_bytes + (uint32_t)g * kb_bytes;
            const int kcol = (kb + g < kb1 ? kb + g : p.num_kb) * elems_per_kb;
            tma_load_3d(a_dst, &tmA, full, kcol, m_tile * BM, pa);
          }
          for (int g = 0; g < p.kb_group; ++g) {
            const uint32_t b_dst = tiles0 + s * stage_bytes + (uint32_t)g * kb_bytes + A_TILE_BYTES;
            const int kcol = (kb + g < kb1 ? kb + g : p.num_kb) * elems_per_kb;
            for (int t = 0; t < p.n_tb; ++t)
              tma_load_3d(b_dst + (uint32_t)(t * p.tb_rows) * KB_BYTES, &tmB, full, kcol, n_tile * p.BN + t * p.tb_rows, pb);
          }
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const int n0 = p.BN > 256 ? 256 : p.BN;      // first UMMA N chunk
    const int n1 = p.BN - n0;                    // second chunk (0 or a multiple of 16)
    const uint32_t idesc0 = make_idesc(kTF32, p.f16 != 0, n0);
    const uint32_t idesc1 = make_idesc(kTF32, p.f16 != 0, n1 > 0 ? n1 : 16);
    const uint64_t desc0 = make_smem_desc(tiles0);
    const uint32_t stage16 = stage_bytes >> 4, kb16 = kb_bytes >> 4, b16 = A_TILE_BYTES >> 4, b1_16 = (256u * KB_BYTES) >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_full0 + 8u * s, ph);
      tc_fence_after();
      if (elect_one()) {
        for (int g = 0; g < p.kb_group; ++g) {
          const uint64_t da = desc0 + (uint64_t)(s * stage16 + g * kb16);
          const uint64_t db0 = da + (uint64_t)b16;
          const uint64_t db1 = db0 + (uint64_t)b1_16;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 UMMA K-steps of 32 B inside the 128 B atom
            const uint32_t acc = (it > 0 || g > 0 || k > 0) ? 1u : 0u;
            tc_mma<kTF32>(tmem_base, da + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc0, acc);
            if (n1 > 0) tc_mma<kTF32>(tmem_base + 256u, da + (uint64_t)(k * 2), db1 + (uint64_t)(k * 2), idesc1, acc);
          }
        }
        tc_commit(bar_empty0 + 8u * s);  // frees the smem slot when these MMAs retire
        if (it == iters - 1) tc_commit(bar_tmem);   // accumulator complete
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global =====
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int row = m_tile * BM + q * 32 + lane;
    float* crow = p.C + (long long)split * p.split_stride + (long long)row * p.ldc;
    if (is_tail) {
      // dense partial tile [tail tile][split][BM][BN]
      float* trow = p.tail_ws + ((long long)((tile - p.tail_cta0) * p.tail_splits + split) * BM + (q * 32 + lane)) * p.BN;
      for (int c = 0; c < p.BN; c += 16) {
        float v[16];
        if (iters > 0) {
          if (c == 0) { mbar_wait(bar_tmem, 0); tc_fence_after(); }
          tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(trow + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      tc_fence_before();
    } else {
    if (iters > 0) {
      mbar_wait(bar_tmem, 0);
      tc_fence_after();
    }
    for (int c = 0; c < p.BN; c += 16) {
      float v[16];
      if (iters > 0) {
        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      const int col = n_tile * p.BN + c;
      if (row < p.M) {
        if (col + 16 <= p.N && ((p.ldc & 3) == 0)) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(crow + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col + i < p.N) crow[col + i] = v[i];
        }
      }
    }
    tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// Same contract as gemm_tn_kernel for 16-bit operands, but two CTAs of one TPC share every B tile: CTA r of the
// pair owns m-tile 2*pair + r (its own 128 rows of A) and HALF of each B chunk; the leader's tcgen05.mma.cta_group::2
// (UMMA 256 x N) reads both halves, accumulators land in each CTA's own TMEM.  Per k-block a CTA now fetches
// 16 KB + BN/2 * 128 B instead of 16 KB + BN * 128 B, so the ring holds 5 stages instead of 3 at BN = 432 and the
// L2 -> shared-memory traffic drops by 39 %: the single-CTA kernel was latency-bound on its 3-stage ring
// (tensor pipe 67-79 % active, profiles/r01_ncu_full_prof_rrr_gemm.csv).
struct KParams2 {
  int M, N, BN, n_tiles, m_tiles;
  int nchunk;          // UMMA instructions per K-step (1 or 2)
  int cn[2];           // UMMA N of each chunk (whole pair)
  int coff[2];         // first tile column of each chunk
  int num_kb, kb_per_split, n_pass;
  int pa[kMaxPass], pb[kMaxPass];
  int stages, tmem_cols, f16;
  float* C;
  long long ldc, split_stride;
  int tail_unit0, tail_splits, tail_kb_per_split;   // pair units >= tail_unit0 are split along K (n_tiles == 1 only)
  float* tail_ws;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_pair(bool f16, int n) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                    const __grid_constant__ CUtensorMap tmB1, const KParams2 p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int h0 = p.cn[0] >> 1, h1 = p.nchunk > 1 ? (p.cn[1] >> 1) : 0;   // B rows this CTA holds per chunk
  const uint32_t stage_bytes = A_TILE_BYTES + (uint32_t)(h0 + h1) * KB_BYTES;
  const uint32_t tiles0 = base + CTRL_BYTES;
  const uint32_t bar_full0 = base;
  const uint32_t bar_empty0 = base + 8u * 16;
  const uint32_t bar_tmem = base + 8u * 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 8 * 33);

  int unit = blockIdx.x >> 1, split = blockIdx.y, kbps = p.kb_per_split;
  const bool is_tail = p.tail_splits > 1 && unit >= p.tail_unit0;
  if (is_tail) {
    const int u = unit - p.tail_unit0;
    unit = p.tail_unit0 + u / p.tail_splits;
    split = u % p.tail_splits;
    kbps = p.tail_kb_per_split;
  }
  const int m_pair = unit / p.n_tiles, n_tile = unit % p.n_tiles;
  const int m_tile = 2 * m_pair + rank;                              // may be one past the end (odd tile count)
  const int m_load = m_tile < p.m_tiles ? m_tile : p.m_tiles - 1;    // keep the TMA box inside the tensor
  const int kb0 = split * kbps;
  const int kb1 = min(p.num_kb, kb0 + kbps);
  const int iters = (kb1 > kb0 ? kb1 - kb0 : 0) * p.n_pass;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full0 + 8u * s, 1);
      mbar_init(bar_empty0 + 8u * s, 1);
    }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 8u * 33), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers are initialised before either signals the other
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Both issuing roles run with the whole warp converged and elect one lane per instruction group: under `if (lane == 0)`
  // ptxas wraps every uniform-datapath instruction (UTMALDG / UTCHMMA) in an ELECT / BRA.U.ANY retry loop, and `it % stages`
  // with a run-time divisor costs ~40 dependent instructions per stage on the single issuing thread.
  if (warp == 0) {
    // ===== TMA producer: own A rows + own halves of the B chunks, counted on the leader's barrier =====
    int s = 0;
    uint32_t ph = 0;
    for (int ps = 0; ps < p.n_pass; ++ps) {
      const int pa = p.pa[ps], pb = p.pb[ps];
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty0 + 8u * s, ph ^ 1u);
        if (elect_one()) {
          const uint32_t a_dst = tiles0 + s * stage_bytes;
          const uint32_t b_dst = a_dst + A_TILE_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(bar_full0 + 8u * s, 2u * stage_bytes);
          const uint32_t full = mapa_cluster(bar_full0 + 8u * s, 0);
          tma_load_3d_pair(a_dst, &tmA, full, kb * 64, m_load * BM, pa);
          tma_load_3d_pair(b_dst, &tmB0, full, kb * 64, n_tile * p.BN + p.coff[0] + rank * h0, pb);
          if (h1 > 0)
            tma_load_3d_pair(b_dst + (uint32_t)h0 * KB_BYTES, &tmB1, full, kb * 64, n_tile * p.BN + p.coff[1] + rank * h1, pb);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one lane of the leader CTA drives the tensor cores of both SMs =====
    if (rank == 0) {
      const uint32_t idesc0 = make_idesc_pair(p.f16 != 0, p.cn[0]);
      const uint32_t idesc1 = make_idesc_pair(p.f16 != 0, p.nchunk > 1 ? p.cn[1] : 16);
      const uint64_t desc0 = make_smem_desc(tiles0);
      const uint32_t stage16 = stage_bytes >> 4, b16 = A_TILE_BYTES >> 4, b1_16 = ((uint32_t)h0 * KB_BYTES) >> 4;
      const uint32_t tm0 = tmem_base + (uint32_t)p.coff[0], tm1 = tmem_base + (uint32_t)p.coff[1];
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(bar_full0 + 8u * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = desc0 + (uint64_t)(s * stage16);
          const uint64_t db0 = da + (uint64_t)b16;
          const uint64_t db1 = db0 + (uint64_t)b1_16;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            tc_mma_pair(tm0, da + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc0, acc);
            if (h1 > 0) tc_mma_pair(tm1, da + (uint64_t)(k * 2), db1 + (uint64_t)(k * 2), idesc1, acc);
          }
          tc_commit_pair(bar_empty0 + 8u * s);   // frees this slot in both CTAs
          if (it == iters - 1) tc_commit_pair(bar_tmem);   // accumulators complete in both CTAs
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global (each CTA drains its own 128 rows) =====
    const int q = warp & 3;
    const int row = m_tile * BM + q * 32 + lane;
    if (iters > 0) {
      mbar_wait(bar_tmem, 0);
      tc_fence_after();
    }
    if (is_tail) {
      float* trow = p.tail_ws + ((long long)(((unit - p.tail_unit0) * 2 + rank) * p.tail_splits + split) * BM + (q * 32 + lane)) * p.BN;
      for (int c = 0; c < p.BN; c += 16) {
        float v[16];
        if (iters > 0) {
          tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(trow + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
      float* crow = p.C + (long long)split * p.split_stride + (long long)row * p.ldc;
      for (int c = 0; c < p.BN; c += 16) {
        float v[16];
        if (iters > 0) {
          tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        const int col = n_tile * p.BN + c;
        if (row < p.M) {
          if (col + 16 <= p.N && ((p.ldc & 3) == 0)) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(crow + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (col + i < p.N) crow[col + i] = v[i];
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();            // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ RRR backward, dense per time bin
// G[c, j*Npad + n] = sum_t V[j,t] * D_t[c,n],   D_t[c,n] = sum_k X[k,t,c] * R[k,t,n]         (src/model/rrr.py:113 autograd)
// The factorised backward feeds the tensor cores the operand R (x) V (r*N columns: r x the flops of the reference's
// einsum).  Here each time bin's D_t = X_t^T R_t (the reference's dbeta_t tile) is formed in TMEM with the contraction over
// that bin's K trials only, and the r rank-one updates G_j += V[j,t] * D_t run on the CUDA cores, in REGISTERS, while the
// tensor cores work on the next bin (D is double-buffered in TMEM).  TMEM reads run at 64 B/clk/SM, so the only per-bin
// TMEM traffic is D_t itself; to make the r accumulators fit the register file a CTA owns 128 rows of c and HALF of the
// neurons (the two CTAs of a c-tile are neighbours in the grid and share the A tile through L2).
//   A = Xb (C1 rows) and B = R (Npad rows), both with column t*Kp + k, Kp = K rounded up to 16 and zeros in the pad:
//       the bin's boxes start at t*Kp + 64 i (32-byte aligned, TMA needs 16), the last one is consumed only up to the
//       UMMA K-steps that still hold trials of the bin
// warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4).
constexpr int DENSE_THREADS = 320;
constexpr int DENSE_R = 3;

struct DenseParams {
  int C1, Npad, T, K, Kp, nb;
  int nc;            // neurons per CTA (Npad / 2)
  int mma_n;         // nc rounded up to 16
  int stages, f16, vbytes;
  int flags;         // experiment switches (VS_DENSE_FLAGS): 1 = spinning waits in the issuing warps, 2 = L2 prefetch of A
  int prefetch;      // prefetch distance in stages
  const double* V;
  float* G;
  long long ldg;
};

__device__ __forceinline__ void mbar_wait_sel(uint32_t bar, uint32_t parity, bool spin) {
  if (spin) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity);
}

__device__ __forceinline__ void tc_ld4_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int W4>   // 4-column chunks per epilogue thread (Npad / 16)
__global__ void __launch_bounds__(DENSE_THREADS, 1)
rrr_bwd_dense_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DenseParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = A_TILE_BYTES + (uint32_t)p.mma_n * KB_BYTES;
  const uint32_t bar_full0 = base, bar_empty0 = base + 8u * 16;
  const uint32_t bar_dfull0 = base + 8u * 32, bar_dempty0 = base + 8u * 34;     // two D buffers
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 8 * 36);
  float* vsm = reinterpret_cast<float*>(base_ptr + CTRL_BYTES);                 // V as float, (r, T)
  const uint32_t tiles0 = base + CTRL_BYTES + (uint32_t)p.vbytes;
  const int m_tile = blockIdx.x >> 1, half = blockIdx.x & 1;
  const bool spin = (p.flags & 1) != 0;
  constexpr uint32_t kDStride = 128;                                             // TMEM columns between the two D buffers

  for (int e = threadIdx.x; e < DENSE_R * p.T; e += DENSE_THREADS) vsm[e] = (float)p.V[e];
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full0 + 8u * s, 1);
      mbar_init(bar_empty0 + 8u * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull0 + 8u * i, 1);
      mbar_init(bar_dempty0 + 8u * i, 8);      // the 8 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 8u * 36), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The two issuing roles run with the WHOLE warp converged and elect one lane per instruction: a loop under `if (lane == 0)`
  // makes ptxas wrap every uniform-datapath instruction (UTMALDG / UTCMMA operands live in uniform registers) in an
  // ELECT / BRA.U.ANY retry loop, and with only ~160 clk of tensor work per stage that scalar overhead (plus the integer
  // division of `it % stages`) was the bottleneck of the first version (tensor pipe 17 % active, profiles/r01_ncu_dense_bwd.txt).
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    // A comes from HBM (each box is used by the two CTAs of a c-tile, once): the even CTA asks for its boxes kPrefetch
    // stages ahead of the ring with an L2 prefetch, so that the 8-stage ring only has to cover L2 latency
    const int kPrefetch = p.prefetch;
    int pt = 0, pi = 0;
    const bool do_pf = (half == 0) && (p.flags & 2);
    if (do_pf && elect_one()) {
      for (int n = 0; n < kPrefetch && pt < p.T; ++n) {
        tma_prefetch_3d(&tmA, pt * p.Kp + pi * 64, m_tile * BM, 0);
        if (++pi == p.nb) { pi = 0; ++pt; }
      }
    }
    {   // every lane tracks the prefetch cursor (the elected lane may change)
      int n = kPrefetch; pt = 0; pi = 0;
      while (n-- > 0 && pt < p.T) { if (++pi == p.nb) { pi = 0; ++pt; } }
    }
    __syncwarp();
    for (int t = 0; t < p.T; ++t) {
      int col = t * p.Kp;
      for (int i = 0; i < p.nb; ++i, col += 64) {
        mbar_wait_sel(bar_empty0 + 8u * s, ph ^ 1u, spin);
        if (elect_one()) {
          const uint32_t a_dst = tiles0 + s * stage_bytes;
          const uint32_t full = bar_full0 + 8u * s;
          mbar_arrive_expect_tx(full, stage_bytes);
          tma_load_3d(a_dst, &tmA, full, col, m_tile * BM, 0);
          tma_load_3d(a_dst + A_TILE_BYTES, &tmB, full, col, half * p.nc, 0);
          if (do_pf && pt < p.T) tma_prefetch_3d(&tmA, pt * p.Kp + pi * 64, m_tile * BM, 0);
        }
        __syncwarp();
        if (pt < p.T && ++pi == p.nb) { pi = 0; ++pt; }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(false, p.f16 != 0, p.mma_n);
    const int last_ksteps = ((p.K - (p.nb - 1) * 64 + 15) >> 4) > 4 ? 4 : ((p.K - (p.nb - 1) * 64 + 15) >> 4);
    const uint64_t desc0 = make_smem_desc(tiles0);
    const uint32_t stage16 = stage_bytes >> 4, b16 = A_TILE_BYTES >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int t = 0; t < p.T; ++t) {
      const int buf = t & 1;
      if (t >= 2) {                   // the epilogue has drained bin t-2 from this buffer
        mbar_wait_sel(bar_dempty0 + 8u * buf, (uint32_t)((t >> 1) - 1) & 1u, spin);
        tc_fence_after();
      }
      const uint32_t dcol = tmem_base + (uint32_t)buf * kDStride;
      for (int i = 0; i < p.nb; ++i) {
        mbar_wait_sel(bar_full0 + 8u * s, ph, spin);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = desc0 + (uint64_t)(s * stage16);
          const uint64_t db = da + (uint64_t)b16;
          const int ksteps = (i == p.nb - 1) ? last_ksteps : 4;
          tc_mma<false>(dcol, da, db, idesc, i > 0 ? 1u : 0u);
          if (ksteps > 1) tc_mma<false>(dcol, da + 2, db + 2, idesc, 1u);
          if (ksteps > 2) tc_mma<false>(dcol, da + 4, db + 4, idesc, 1u);
          if (ksteps > 3) tc_mma<false>(dcol, da + 6, db + 6, idesc, 1u);
          tc_commit(bar_empty0 + 8u * s);
          if (i == p.nb - 1) tc_commit(bar_dfull0 + 8u * buf);     // D_t complete
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3, hsel = (warp - 2) >> 2;
    constexpr int W = W4 * 4;                 // columns per thread
    const uint32_t t_d = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hsel * W);
    float g0[W], g1[W], g2[W];
#pragma unroll
    for (int i = 0; i < W; ++i) g0[i] = g1[i] = g2[i] = 0.f;
    for (int t = 0; t < p.T; ++t) {
      const int buf = t & 1;
      mbar_wait(bar_dfull0 + 8u * buf, (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const float v0 = vsm[t], v1 = vsm[p.T + t], v2 = vsm[2 * p.T + t];
      const uint32_t td = t_d + (uint32_t)buf * kDStride;
#pragma unroll
      for (int c0 = 0; c0 < W4; c0 += 3) {    // 12 columns per TMEM round trip
        uint32_t d[12];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c0 + c < W4) tc_ld4_issue(td + (uint32_t)((c0 + c) * 4), d + c * 4);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          if (c0 * 4 + i < W) {
            const float dv = __uint_as_float(d[i]);
            g0[c0 * 4 + i] = fmaf(v0, dv, g0[c0 * 4 + i]);
            g1[c0 * 4 + i] = fmaf(v1, dv, g1[c0 * 4 + i]);
            g2[c0 * 4 + i] = fmaf(v2, dv, g2[c0 * 4 + i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dempty0 + 8u * buf);     // the tensor cores may overwrite this D buffer
    }
    // write the tile: G[c, j*Npad + n]
    const int row = m_tile * BM + q * 32 + lane;
    if (row < p.C1) {
      float* grow = p.G + (long long)row * p.ldg + half * p.nc + hsel * W;
#pragma unroll
      for (int c = 0; c < W4; ++c) {
        if (half * p.nc + hsel * W + c * 4 < p.Npad) {
          *reinterpret_cast<float4*>(grow + c * 4) = make_float4(g0[c * 4], g0[c * 4 + 1], g0[c * 4 + 2], g0[c * 4 + 3]);
          *reinterpret_cast<float4*>(grow + p.Npad + c * 4) = make_float4(g1[c * 4], g1[c * 4 + 1], g1[c * 4 + 2], g1[c * 4 + 3]);
          *reinterpret_cast<float4*>(grow + 2 * p.Npad + c * 4) = make_float4(g2[c * 4], g2[c * 4 + 1], g2[c * 4 + 2], g2[c * 4 + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ------------------------------------------------------------------ RRR backward, dense per time bin, CTA pairs
// Same contraction as rrr_bwd_dense_kernel, other decomposition: the L2 -> SM path delivers ~45 B/clk/SM, and a CTA that
// owns 128 rows of c and HALF of the neurons ingests 26 KB per 64-trial box for a 128 x 72 tile -- that kernel runs at the
// ingest limit (5.3 GB per launch, profiles/r01_ncu_dense_bwd.txt).  Here a CTA PAIR (cta_group::2, UMMA 256 x Npad) owns
// 256 rows of c and ALL neurons: each CTA ingests its 128 rows of A and half of R (25 KB per box for a 128 x 144 tile, half
// the bytes per flop).  The price is accumulator space: 128 x 3*Npad fp32 per CTA.  TMEM reads run at ~64 B/clk/SM, so as
// much as possible lives in registers: the issuing warpgroup gives its registers away (setmaxnreg 40 / 232), the epilogue
// threads keep G_1 and G_2 (2 x Npad/2 registers each); G_0 lives in TMEM next to the double-buffered D_t
// (3*Npad <= 512 columns) and is updated with tcgen05.ld / fma / tcgen05.st.
struct DensePairParams {
  int C1, m_tiles, Npad, T, K, Kp, nb;
  int stages, tmem_cols, f16, vbytes;
  int r_planes;              // 1, or 2: R = hi + lo (each plane is a separate B tile, both accumulate into the same D_t)
  const double* V;
  float* G;
  long long ldg;
  const float* scaleT;       // per (feature row, time bin) weight of the rank-one updates (1/std of the z-score), or NULL
  long long ldt;
  // kDV variant: instead of accumulating G_j += V[j,t] D_t, contract D_t with the U slab of the CTA's rows:
  //   dvpart[((cta * 8 + epilogue warp) * T + t) * 3 + j] = sum over the warp's 32 rows c and its columns n of U[n,c,j] * scaleT[c,t] * D_t[c,n]
  const float* U32;          // [N][ldu][3] fp32
  long long ldu;
  int N;
  double* dvpart;
};

__device__ __forceinline__ void tc_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

constexpr int DENSE_PAIR_THREADS = 384;   // warpgroup 0: TMA + MMA (+2 idle warps), warpgroups 1-2: epilogue

template <int HW8, bool kDV>   // 8-column chunks per epilogue thread (Npad / 16); kDV: the dV contraction instead of the dU accumulation
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DENSE_PAIR_THREADS, 1)
rrr_bwd_dense_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DensePairParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int h = p.Npad >> 1;                                   // B rows per CTA = columns per epilogue thread
  const uint32_t b_bytes = (uint32_t)h * KB_BYTES;             // one residual plane of the B tile
  const uint32_t stage_bytes = A_TILE_BYTES + (uint32_t)p.r_planes * b_bytes;
  const uint32_t bar_full0 = base, bar_empty0 = base + 8u * 16;
  const uint32_t bar_dfull0 = base + 8u * 32, bar_dempty0 = base + 8u * 34;     // two D buffers
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 8 * 36);
  float* vsm = reinterpret_cast<float*>(base_ptr + CTRL_BYTES);
  const uint32_t tiles0 = base + CTRL_BYTES + (uint32_t)p.vbytes;
  const int m_tile = 2 * (int)(blockIdx.x >> 1) + rank;
  const int m_load = m_tile < p.m_tiles ? m_tile : p.m_tiles - 1;
  const uint32_t Np = (uint32_t)p.Npad;                        // TMEM columns: D buffers at 0 and Npad, G_0 at 2*Npad

  for (int e = threadIdx.x; e < DENSE_R * p.T; e += DENSE_PAIR_THREADS) vsm[e] = (float)p.V[e];
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full0 + 8u * s, 1);
      mbar_init(bar_empty0 + 8u * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull0 + 8u * i, 1);
      mbar_init(bar_dempty0 + 8u * i, 16);     // the 8 epilogue warps of both CTAs (only the leader's copies are waited on)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 8u * 36), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ---- warpgroup 0 hands most of its registers to the epilogue warpgroups (two of the three accumulators live there)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < p.T; ++t) {
        int col = t * p.Kp;
        for (int i = 0; i < p.nb; ++i, col += 64) {
          mbar_wait(bar_empty0 + 8u * s, ph ^ 1u);
          if (elect_one()) {
            const uint32_t a_dst = tiles0 + s * stage_bytes;
            if (rank == 0) mbar_arrive_expect_tx(bar_full0 + 8u * s, 2u * stage_bytes);
            const uint32_t full = mapa_cluster(bar_full0 + 8u * s, 0);
            tma_load_3d_pair(a_dst, &tmA, full, col, m_load * BM, 0);
            tma_load_3d_pair(a_dst + A_TILE_BYTES, &tmB, full, col, rank * h, 0);
            if (p.r_planes > 1) tma_load_3d_pair(a_dst + A_TILE_BYTES + b_bytes, &tmB, full, col, rank * h, 1);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 1 && rank == 0) {
      const uint32_t idesc = make_idesc_pair(p.f16 != 0, p.Npad);
      const int lk = (p.K - (p.nb - 1) * 64 + 15) >> 4;
      const int last_ksteps = lk > 4 ? 4 : lk;
      const uint64_t desc0 = make_smem_desc(tiles0);
      const uint32_t stage16 = stage_bytes >> 4, b16 = A_TILE_BYTES >> 4, lo16 = b_bytes >> 4;
      const bool two = p.r_planes > 1;
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < p.T; ++t) {
        const int buf = t & 1;
        if (t >= 2) {                   // every epilogue warp of the pair has drained bin t-2 from this D buffer
          mbar_wait(bar_dempty0 + 8u * buf, (uint32_t)((t >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t dcol = tmem_base + (uint32_t)buf * Np;
        for (int i = 0; i < p.nb; ++i) {
          mbar_wait(bar_full0 + 8u * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = desc0 + (uint64_t)(s * stage16);
            const uint64_t db = da + (uint64_t)b16;
            const int ksteps = (i == p.nb - 1) ? last_ksteps : 4;
            if (two) {
              // residual planes of R: the small lo product first, then hi, both into the same D_t
              const uint64_t dl = db + (uint64_t)lo16;
              tc_mma_pair(dcol, da, dl, idesc, i > 0 ? 1u : 0u);
              tc_mma_pair(dcol, da, db, idesc, 1u);
              if (ksteps > 1) { tc_mma_pair(dcol, da + 2, dl + 2, idesc, 1u); tc_mma_pair(dcol, da + 2, db + 2, idesc, 1u); }
              if (ksteps > 2) { tc_mma_pair(dcol, da + 4, dl + 4, idesc, 1u); tc_mma_pair(dcol, da + 4, db + 4, idesc, 1u); }
              if (ksteps > 3) { tc_mma_pair(dcol, da + 6, dl + 6, idesc, 1u); tc_mma_pair(dcol, da + 6, db + 6, idesc, 1u); }
            } else {
            tc_mma_pair(dcol, da, db, idesc, i > 0 ? 1u : 0u);
            if (ksteps > 1) tc_mma_pair(dcol, da + 2, db + 2, idesc, 1u);
            if (ksteps > 2) tc_mma_pair(dcol, da + 4, db + 4, idesc, 1u);
            if (ksteps > 3) tc_mma_pair(dcol, da + 6, db + 6, idesc, 1u);
            }
            tc_commit_pair(bar_empty0 + 8u * s);
            if (i == p.nb - 1) tc_commit_pair(bar_dfull0 + 8u * buf);      // D_t complete in both CTAs
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    constexpr int W = HW8 * 8;              // columns per thread
    const uint32_t t_d = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hsel * W);
    const uint32_t t_g0 = t_d + 2u * Np;
    const uint32_t dempty_leader0 = mapa_cluster(bar_dempty0, 0);
    float g1[W], g2[W];
#pragma unroll
    for (int i = 0; i < W; ++i) g1[i] = g2[i] = 0.f;
    // exact-operand mode: this thread's feature row carries its own z-score scale per time bin (rows past C1 are zero
    // tiles: any in-range scale will do); four bins per 128-bit load
    const int srow = (m_tile * BM + q * 32 + lane) < p.C1 ? (m_tile * BM + q * 32 + lane) : p.C1 - 1;
    if constexpr (kDV) {
      // the U slab of this thread's feature row: U_1, U_2 in registers, U_0 parked in the TMEM columns the dU variant uses for G_0
#pragma unroll
      for (int c8 = 0; c8 < HW8; ++c8) {
        uint32_t u0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = hsel * W + c8 * 8 + i;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f;
          if (n < p.N) {
            a0 = __ldg(p.U32 + u32g_index(n, srow, 0, p.N));
            a1 = __ldg(p.U32 + u32g_index(n, srow, 1, p.N));
            a2 = __ldg(p.U32 + u32g_index(n, srow, 2, p.N));
          }
          u0[i] = __float_as_uint(a0); g1[c8 * 8 + i] = a1; g2[c8 * 8 + i] = a2;
        }
        tc_st8(t_g0 + (uint32_t)(c8 * 8), u0);
      }
      tc_wait_st();
    }
    const float* sct = p.scaleT ? p.scaleT + (long long)srow * p.ldt : nullptr;
    float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f);
    for (int t = 0; t < p.T; ++t) {
      const int buf = t & 1;
      if (sct && (t & 3) == 0) sc4 = __ldg(reinterpret_cast<const float4*>(sct + t));      // ldt % 4 == 0, rows padded to ldt
      float sc = (t & 3) == 0 ? sc4.x : ((t & 3) == 1 ? sc4.y : ((t & 3) == 2 ? sc4.z : sc4.w));
      mbar_wait(bar_dfull0 + 8u * buf, (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const float v0 = vsm[t] * sc, v1 = vsm[p.T + t] * sc, v2 = vsm[2 * p.T + t] * sc;
      const uint32_t td = t_d + (uint32_t)buf * Np;
      // 16 columns of D and G_0 per TMEM round trip, software-pipelined: the loads of round r+1 are in flight while the
      // FMAs and the G_0 store of round r run
      constexpr int kN = 16, kRounds = (HW8 + 1) / 2;
      uint32_t d[2][kN], a0[2][kN];
      auto issue = [&](int r, int slot) {
        const int c = 2 * r;
        tc_ld8_issue(td + (uint32_t)(c * 8), d[slot]);
        if (c + 1 < HW8) tc_ld8_issue(td + (uint32_t)(c * 8 + 8), d[slot] + 8);
        if (t > 0) {
          tc_ld8_issue(t_g0 + (uint32_t)(c * 8), a0[slot]);
          if (c + 1 < HW8) tc_ld8_issue(t_g0 + (uint32_t)(c * 8 + 8), a0[slot] + 8);
        } else {
#pragma unroll
          for (int i = 0; i < kN; ++i) a0[slot][i] = 0u;
        }
      };
      auto release_d = [&]() {                 // every D column of this bin is in registers: the buffer may be overwritten
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(dempty_leader0 + 8u * (uint32_t)buf);
      };
      if constexpr (kDV) {
        auto issue_dv = [&](int r, int slot) {
          const int c = 2 * r;
          tc_ld8_issue(td + (uint32_t)(c * 8), d[slot]);
          tc_ld8_issue(t_g0 + (uint32_t)(c * 8), a0[slot]);
          if (c + 1 < HW8) { tc_ld8_issue(td + (uint32_t)(c * 8 + 8), d[slot] + 8); tc_ld8_issue(t_g0 + (uint32_t)(c * 8 + 8), a0[slot] + 8); }
        };
        // dV is a small difference of large sums (2.6 M products per entry): eight columns at a time in fp32, then float64
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        issue_dv(0, 0);
        tc_wait_ld();
        if (kRounds == 1) release_d();
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
          const int slot = r & 1, c = 2 * r;
          const bool two = (c + 1 < HW8);
          if (r + 1 < kRounds) issue_dv(r + 1, slot ^ 1);
#pragma unroll
          for (int i8 = 0; i8 < kN; i8 += 8) {
            if (i8 == 0 || two) {
              float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
              for (int i = i8; i < i8 + 8; ++i) {
                const int col = (c * 8 + i) < W ? (c * 8 + i) : 0;
                // columns past N hold whatever the pad rows of the residual operand contain (never written): 0 * NaN must not happen
                const float dv = (hsel * W + c * 8 + i) < p.N ? __uint_as_float(d[slot][i]) : 0.f;
                p0 = fmaf(__uint_as_float(a0[slot][i]), dv, p0);
                p1 = fmaf(g1[col], dv, p1);
                p2 = fmaf(g2[col], dv, p2);
              }
              q0 += (double)p0; q1 += (double)p1; q2 += (double)p2;
            }
          }
          if (r + 1 < kRounds) {
            tc_wait_ld();
            if (r + 2 == kRounds) release_d();
          }
        }
        // the 1/std of this row and bin (sc), then the sum over the warp's 32 feature rows.  Rows past C1 do not exist: a
        // phantom second tile of the last pair (odd tile count) re-reads the last real tile and must not be counted twice
        if (m_tile * BM + q * 32 + lane >= p.C1) sc = 0.f;
        q0 *= (double)sc; q1 *= (double)sc; q2 *= (double)sc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          q1 += __shfl_xor_sync(0xffffffffu, q1, o);
          q2 += __shfl_xor_sync(0xffffffffu, q2, o);
        }
        if (lane == 0) {
          double* o3 = p.dvpart + (((long long)blockIdx.x * 8 + (warp - 4)) * p.T + t) * 3;
          o3[0] = q0; o3[1] = q1; o3[2] = q2;
        }
        continue;
      }
      issue(0, 0);
      tc_wait_ld();
      if (kRounds == 1) release_d();
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int slot = r & 1, c = 2 * r;
        const bool two = (c + 1 < HW8);
        if (r + 1 < kRounds) issue(r + 1, slot ^ 1);
#pragma unroll
        for (int i = 0; i < kN; ++i) {
          if (i < 8 || two) {
            const int col = (c * 8 + i) < W ? (c * 8 + i) : 0;
            const float dv = __uint_as_float(d[slot][i]);
            g1[col] = fmaf(v1, dv, g1[col]);
            g2[col] = fmaf(v2, dv, g2[col]);
            a0[slot][i] = __float_as_uint(fmaf(v0, dv, __uint_as_float(a0[slot][i])));
          }
        }
        tc_st8(t_g0 + (uint32_t)(c * 8), a0[slot]);
        if (two) tc_st8(t_g0 + (uint32_t)(c * 8 + 8), a0[slot] + 8);
        if (r + 1 < kRounds) {
          tc_wait_ld();
          if (r + 2 == kRounds) release_d();
        }
      }
      tc_wait_st();
    }
    // write the tile: G[c, j*Npad + n]
    const int row = m_tile * BM + q * 32 + lane;
    float* grow = p.G + (long long)row * p.ldg + hsel * W;
#pragma unroll
    for (int c = 0; c < HW8; ++c) {
      if constexpr (kDV) break;
      uint32_t a0[8];
      tc_ld8_issue(t_g0 + (uint32_t)(c * 8), a0);
      tc_wait_ld();
      if (row < p.C1) {
        float4* o0 = reinterpret_cast<float4*>(grow + c * 8);
        float4* o1 = reinterpret_cast<float4*>(grow + p.Npad + c * 8);
        float4* o2 = reinterpret_cast<float4*>(grow + 2 * p.Npad + c * 8);
        o0[0] = make_float4(__uint_as_float(a0[0]), __uint_as_float(a0[1]), __uint_as_float(a0[2]), __uint_as_float(a0[3]));
        o0[1] = make_float4(__uint_as_float(a0[4]), __uint_as_float(a0[5]), __uint_as_float(a0[6]), __uint_as_float(a0[7]));
        o1[0] = make_float4(g1[c * 8 + 0], g1[c * 8 + 1], g1[c * 8 + 2], g1[c * 8 + 3]);
        o1[1] = make_float4(g1[c * 8 + 4], g1[c * 8 + 5], g1[c * 8 + 6], g1[c * 8 + 7]);
        o2[0] = make_float4(g2[c * 8 + 0], g2[c * 8 + 1], g2[c * 8 + 2], g2[c * 8 + 3]);
        o2[1] = make_float4(g2[c * 8 + 4], g2[c * 8 + 5], g2[c * 8 + 6], g2[c * 8 + 7]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ RRR forward, dense per time bin, CTA pairs
// yraw[(t,k), n] = sum_c Xc[(t,k), c] * beta'_t[n, c],   beta'_t[n,c] = (U[n,c,:] . V[:,t]) / std[t,c]      (src/model/rrr.py:105-116)
// The factorised forward (Z = X U, 3 N columns) executes r = 3x the flops of the reference's einsum and cannot use the
// exact integer operand (the z-score scale depends on (t, c)).  Here the coefficient tile of a time bin is GENERATED ON
// CHIP: nine warps per CTA read U (fp32, L2 resident), form beta'_t for THREE time bins at once -- so a U chunk is fetched
// once per three bins -- split it into hi + lo half planes and write the K-major, 128B-swizzled B tiles the tensor cores
// read; A = Xc (exact integers frame - round(mean), half) comes by TMA.  A CTA pair (cta_group::2, UMMA 256 x Npad) owns
// 256 trials of three bins: 3 accumulators of Npad columns in TMEM (3 * 144 = 432 of 512).  Each CTA generates the B rows
// of HALF of the neurons.  beta and Z never exist in HBM.
// warp 0 TMA (A tiles), warp 1 MMA issue, warps 2..10 B generators (item = (neuron, 16-byte chunk): two per thread and
// k-block at N = 144, the next item's U record requested while the current one is computed), warps 2..9 drain the
// accumulators at the end.
constexpr int FWD_THREADS = 352;
constexpr int FWD_BINS = 3;
constexpr int FWD_GEN_WARP0 = 2;
constexpr int FWD_GEN_WARPS = 9;
constexpr int FWD_GEN_THREADS = FWD_GEN_WARPS * 32;     // 288: 8 sixteen-byte chunks x 36 neuron slots

struct DenseFwdParams {
  int K, T, C1, N, Npad;
  int n_rt;                  // 256-trial row tiles per time bin
  int num_kb;                // 64-feature k-blocks
  int kb_per_split;          // k-blocks per accumulation run (blockIdx.y selects the run; its partial goes to Y + run * split_stride)
  long long split_stride;
  const float* U32;          // [N][ldu][3] fp32, zero past C1
  const float* isd;          // [T][ldu] fp32: 1/std[t,c], zero past C1
  long long ldu;
  const double* V;           // (3, T)
  const float* bscale;       // [T]: power-of-two scale of bin t's B tiles (keeps both half planes in the normal range)
  float* Y;                  // [T*K][ldy] fp32: sum_c Xc beta' (unscaled)
  long long ldy;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FWD_THREADS, 1)
rrr_fwd_dense_pair_kernel(const __grid_constant__ CUtensorMap tmA, const DenseFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int h = p.Npad >> 1;                                   // B rows (neurons) this CTA generates
  const uint32_t b_bytes = (uint32_t)h * KB_BYTES;             // one (bin, plane) B tile
  const uint32_t stage_bytes = FWD_BINS * (A_TILE_BYTES + 2u * b_bytes);
  constexpr int kStages = 2;
  const uint32_t bar_fullA0 = base, bar_fullB0 = base + 8u * 4, bar_empty0 = base + 8u * 8, bar_done = base + 8u * 12;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 8 * 13);
  const uint32_t tiles0 = base + CTRL_BYTES;

  const int item = (int)(blockIdx.x >> 1);
  const int grp = item / p.n_rt, rt = item % p.n_rt;
  const int t0 = grp * FWD_BINS;
  const int nbins = (p.T - t0) < FWD_BINS ? (p.T - t0) : FWD_BINS;
  const int row0 = rt * 256 + rank * BM;                       // first trial of this CTA's 128 rows
  // accumulation run of this CTA pair: the fp32 accumulator is truncated at every MMA, so a long contraction drifts by
  // ~(#MMA steps) * 2^-24 of its running magnitude; the host splits the 18,260-feature contraction into a few runs whose
  // partial tiles the epilogue kernel adds in float64
  const int kb0 = (int)blockIdx.y * p.kb_per_split;
  const int kb1 = (kb0 + p.kb_per_split) < p.num_kb ? (kb0 + p.kb_per_split) : p.num_kb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_fullA0 + 8u * s, 1);
      mbar_init(bar_fullB0 + 8u * s, 2 * FWD_GEN_WARPS);       // the generator warps of both CTAs (only the leader's copy is waited on)
      mbar_init(bar_empty0 + 8u * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 8u * 13), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 trials of each of the bins, one 64-feature box per bin and k-block =====
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(bar_empty0 + 8u * s, ph ^ 1u);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(bar_fullA0 + 8u * s, 2u * (uint32_t)nbins * A_TILE_BYTES);
        const uint32_t full = mapa_cluster(bar_fullA0 + 8u * s, 0);
        for (int b = 0; b < nbins; ++b)
          tma_load_3d_pair(tiles0 + s * stage_bytes + (uint32_t)b * A_TILE_BYTES, &tmA, full, kb * 64, (t0 + b) * p.K + row0, 0);
      }
      __syncwarp();
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): per k-block and bin, 4 K-steps x (lo plane, hi plane) into the bin's accumulator =====
    if (rank == 0) {
      const uint32_t idesc = make_idesc_pair(true, p.Npad);
      const uint64_t desc0 = make_smem_desc(tiles0);
      const uint32_t stage16 = stage_bytes >> 4, a16 = A_TILE_BYTES >> 4, b16 = b_bytes >> 4, boff16 = (FWD_BINS * A_TILE_BYTES) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_fullA0 + 8u * s, ph);
        mbar_wait(bar_fullB0 + 8u * s, ph);
        tc_fence_after();
        if (elect_one()) {
          for (int b = 0; b < nbins; ++b) {
            const uint64_t da = desc0 + (uint64_t)(s * stage16 + b * a16);
            const uint64_t dh = desc0 + (uint64_t)(s * stage16 + boff16 + (2 * b) * b16);
            const uint64_t dl = dh + (uint64_t)b16;
            const uint32_t dcol = tmem_base + (uint32_t)(b * p.Npad);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc_mma_pair(dcol, da + (uint64_t)(k * 2), dl + (uint64_t)(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              tc_mma_pair(dcol, da + (uint64_t)(k * 2), dh + (uint64_t)(k * 2), idesc, 1u);
            }
          }
          tc_commit_pair(bar_empty0 + 8u * s);
          if (kb == kb1 - 1) tc_commit_pair(bar_done);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  }
  if (warp >= FWD_GEN_WARP0) {
    // ===== B generators: thread = (16-byte chunk q of the 128-byte row, neuron slot ns); items (k-block, neuron ns + 36 it)
    // are walked in one flat sequence and the U record of the NEXT item is requested before the current one is computed:
    // with one item per thread and k-block nothing overlapped the L2 latency of the record (ncu: tensor pipe 33 %, issue 33 %)
    const int g = threadIdx.x - 32 * FWD_GEN_WARP0;
    const int q = g & 7, ns = g >> 3;
    constexpr int kSlots = FWD_GEN_THREADS / 8;
    const int NI = (h + kSlots - 1) / kSlots;                  // items per thread and k-block (2 at N = 144)
    float v[FWD_BINS][3];
#pragma unroll
    for (int b = 0; b < FWD_BINS; ++b) {
      const int t = (t0 + b) < p.T ? (t0 + b) : (p.T - 1);
      const float bsc = __ldg(p.bscale + t);
#pragma unroll
      for (int j = 0; j < 3; ++j) v[b][j] = (float)p.V[(long long)j * p.T + t] * bsc;       // scale folded into V (power of two: exact)
    }
    const uint32_t fullB_leader0 = mapa_cluster(bar_fullB0, 0);
    // u32g_index layout: record (kb, n) = 6 x 8 float4, lane q takes float4 (i, q): 128 contiguous bytes per 8 lanes
    auto load_u = [&](int kb, int it, float4 (&dst)[6]) {
      const int nl = ns + kSlots * it;
      const int n = rank * h + nl;
      if (nl < h && n < p.N) {
        const float4* up = reinterpret_cast<const float4*>(p.U32) + (((long long)kb * p.N + n) * 6) * 8 + q;
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[i] = __ldg(up + i * 8);
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 un[6];
    if (kb0 < kb1) load_u(kb0, 0, un);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      const long long c0 = (long long)kb * 64 + q * 8;
      float4 sa[FWD_BINS][2];                                  // 1/std of the 8 features of this chunk, per bin (L1-resident after the first warp)
      uint32_t bt0 = 0;
      for (int it = 0; it < NI; ++it) {
        float4 uc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) uc[i] = un[i];
        {   // request the next item's record
          int kbn = kb, itn = it + 1;
          if (itn == NI) { itn = 0; ++kbn; }
          if (kbn < kb1) load_u(kbn, itn, un);
        }
        if (it == 0) {
#pragma unroll
          for (int b = 0; b < FWD_BINS; ++b) {
            const int t = (t0 + b) < p.T ? (t0 + b) : (p.T - 1);
            const float4* ip = reinterpret_cast<const float4*>(p.isd + (long long)t * p.ldu + c0);
            sa[b][0] = __ldg(ip); sa[b][1] = __ldg(ip + 1);
          }
          mbar_wait(bar_empty0 + 8u * s, ph ^ 1u);
          bt0 = tiles0 + s * stage_bytes + FWD_BINS * A_TILE_BYTES;
        }
        const int nl = ns + kSlots * it;
        if (nl < h) {
          const float u[24] = {uc[0].x, uc[0].y, uc[0].z, uc[0].w, uc[1].x, uc[1].y, uc[1].z, uc[1].w, uc[2].x, uc[2].y, uc[2].z, uc[2].w,
                               uc[3].x, uc[3].y, uc[3].z, uc[3].w, uc[4].x, uc[4].y, uc[4].z, uc[4].w, uc[5].x, uc[5].y, uc[5].z, uc[5].w};
          const uint32_t roff = (uint32_t)(nl >> 3) * 1024u + (uint32_t)(nl & 7) * 128u + ((uint32_t)(q ^ (nl & 7)) << 4);   // SWIZZLE_128B
#pragma unroll
          for (int b = 0; b < FWD_BINS; ++b) {
            if (b < nbins) {
              const float sc[8] = {sa[b][0].x, sa[b][0].y, sa[b][0].z, sa[b][0].w, sa[b][1].x, sa[b][1].y, sa[b][1].z, sa[b][1].w};
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                f[i] = fmaf(u[3 * i + 2], v[b][2], fmaf(u[3 * i + 1], v[b][1], u[3 * i] * v[b][0])) * sc[i];
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                hi[i] = pack_h2(f[2 * i], f[2 * i + 1]);
                const float2 hf = unpack_h2(hi[i]);
                lo[i] = pack_h2(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
              }
              const uint32_t dst = bt0 + (uint32_t)(2 * b) * b_bytes + roff;
              sts128(dst, hi[0], hi[1], hi[2], hi[3]);
              sts128(dst + b_bytes, lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      }
      fence_proxy_async_smem();                 // generic-proxy stores -> visible to the tensor cores' async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(fullB_leader0 + 8u * (uint32_t)s);
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
  }
  if (warp >= 2 && warp < 10) {
    // ===== drain: Y[(t, k), n] = accumulator / scale; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int qd = warp & 3, hsel = (warp - 2) >> 2;
    const int W = h;                                           // columns per thread (Npad / 2, a multiple of 8)
    const int k = row0 + qd * 32 + lane;
    for (int b = 0; b < nbins; ++b) {
      const float inv = 1.0f / __ldg(p.bscale + t0 + b);
      const uint32_t ta = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(b * p.Npad + hsel * W);
      float* yrow = p.Y + (long long)blockIdx.y * p.split_stride + ((long long)(t0 + b) * p.K + k) * p.ldy + hsel * W;
      for (int c = 0; c < W; c += 8) {
        uint32_t d[8];
        tc_ld8_issue(ta + (uint32_t)c, d);
        tc_wait_ld();
        if (k < p.K) {
          float4* o = reinterpret_cast<float4*>(yrow + c);
          o[0] = make_float4(__uint_as_float(d[0]) * inv, __uint_as_float(d[1]) * inv, __uint_as_float(d[2]) * inv, __uint_as_float(d[3]) * inv);
          o[1] = make_float4(__uint_as_float(d[4]) * inv, __uint_as_float(d[5]) * inv, __uint_as_float(d[6]) * inv, __uint_as_float(d[7]) * inv);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// C tile = ordered sum of the tail-wave partial tiles
__global__ void __launch_bounds__(256) tail_reduce_kernel(const float* __restrict__ ws, int tail_cta0, int splits, int n_tiles, int BN,
                                                          int M, int N, float* __restrict__ C, long long ldc) {
  const int tt = blockIdx.y;
  const int tile = tail_cta0 + tt, m_tile = tile / n_tiles, n_tile = tile % n_tiles;
  const int e4 = blockIdx.x * 256 + threadIdx.x;           // float4 index inside the BM x BN tile
  if (e4 * 4 >= BM * BN) return;
  const int rl = (e4 * 4) / BN, c = (e4 * 4) % BN;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < splits; ++sp) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(ws + ((long long)(tt * splits + sp) * BM + rl) * BN + c));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const int row = m_tile * BM + rl, col = n_tile * BN + c;
  if (row >= M) return;
  float* dst = C + (long long)row * ldc + col;
  const float sv[4] = {s.x, s.y, s.z, s.w};
  if (col + 4 <= N && ((ldc & 3) == 0)) {
    *reinterpret_cast<float4*>(dst) = s;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (col + i < N) dst[i] = sv[i];
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D map over a (planes, rows, K) K-major operand; box = (128 B of K, box_rows, 1)
static int make_map(CUtensorMap* m, const Operand& op, bool tf32, bool f16, int box_rows) {
  EncodeTiledFn enc = get_encode();
  VS_REQUIRE(enc != nullptr, VS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int esz = tf32 ? 4 : 2;
  VS_REQUIRE(((uintptr_t)op.ptr & 15) == 0, VS_ERR_INVALID, "GEMM operand must be 16-byte aligned");
  VS_REQUIRE((op.ld * esz) % 16 == 0, VS_ERR_INVALID, "GEMM operand pitch must be a multiple of 16 bytes (ld=%lld)", (long long)op.ld);
  const int planes = op.planes > 0 ? op.planes : 1;
  VS_REQUIRE(planes == 1 || (op.plane_stride * esz) % 16 == 0, VS_ERR_INVALID, "plane stride must be a multiple of 16 bytes");
  cuuint64_t dims[3] = {(cuuint64_t)op.k, (cuuint64_t)op.rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)op.ld * esz, (cuuint64_t)(planes > 1 ? op.plane_stride : op.ld * op.rows) * esz};
  cuuint32_t box[3] = {(cuuint32_t)(tf32 ? 32 : 64), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 3, const_cast<void*>(op.ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VS_REQUIRE(r == CUDA_SUCCESS, VS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows=%lld k=%lld ld=%lld box_rows=%d",
             (int)r, (long long)op.rows, (long long)op.k, (long long)op.ld, box_rows);
  return VS_OK;
}

size_t balance_ws_bytes() { return (size_t)kNumSMs * BM * kMaxBN * sizeof(float); }

bool gemm_supported(const GemmDesc& g) {
  const int esz = g.tf32 ? 4 : 2;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if (((uintptr_t)g.A.ptr & 15) || ((uintptr_t)g.B.ptr & 15)) return false;
  if ((g.A.ld * esz) % 16 || (g.B.ld * esz) % 16) return false;
  if (g.n_pass < 1 || g.n_pass > kMaxPass) return false;
  return true;
}

int pick_bn(long long N) {
  long long n16 = round_up(N, 16);
  if (n16 <= kMaxBN) return (int)n16;
  long long tiles = ceil_div(n16, kMaxBN);
  return (int)round_up(ceil_div(n16, tiles), 16);
}

// CTA-pair route (16-bit operands, at least two m-tiles): see gemm_tn_pair_kernel
static bool pair_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VS_GEMM_PAIR");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

static int gemm_tn_pair(const GemmDesc& g, int BN, cudaStream_t stream) {
  KParams2 p;
  p.M = (int)g.M; p.N = (int)g.N; p.BN = BN;
  p.m_tiles = (int)ceil_div(g.M, BM); p.n_tiles = (int)ceil_div(g.N, BN);
  if (BN <= 256) {
    p.nchunk = 1; p.cn[0] = BN; p.cn[1] = 0; p.coff[0] = 0; p.coff[1] = 0;
  } else {
    p.nchunk = 2; p.cn[0] = (int)round_up(BN / 2, 16); p.cn[1] = BN - p.cn[0]; p.coff[0] = 0; p.coff[1] = p.cn[0];
  }
  const int h0 = p.cn[0] / 2, h1 = p.cn[1] / 2;
  p.num_kb = (int)ceil_div(g.K, 64);
  int splits = g.splits > 0 ? g.splits : 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (int)ceil_div(p.num_kb, splits);
  splits = (int)ceil_div(p.num_kb, p.kb_per_split);
  p.n_pass = g.n_pass;
  for (int i = 0; i < kMaxPass; ++i) { p.pa[i] = i < g.n_pass ? g.pa[i] : 0; p.pb[i] = i < g.n_pass ? g.pb[i] : 0; }
  const int stage_bytes = A_TILE_BYTES + (h0 + h1) * KB_BYTES;
  int stages = (227 * 1024 - CTRL_BYTES - 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  VS_REQUIRE(stages >= 2, VS_ERR_UNSUPPORTED, "tcgen05 GEMM: tile too large for shared memory");
  p.stages = stages;
  int tcols = 32;
  while (tcols < BN) tcols <<= 1;
  p.tmem_cols = tcols;
  p.f16 = g.f16 ? 1 : 0;
  p.C = g.C; p.ldc = g.ldc; p.split_stride = g.split_stride;
  VS_REQUIRE(splits == 1 || g.split_stride >= g.M * g.ldc, VS_ERR_INVALID, "split-K needs split_stride >= M*ldc");

  CUtensorMap tmA, tmB0, tmB1;
  int rc = make_map(&tmA, g.A, false, g.f16, BM);
  if (rc) return rc;
  rc = make_map(&tmB0, g.B, false, g.f16, h0);
  if (rc) return rc;
  rc = make_map(&tmB1, g.B, false, g.f16, h1 > 0 ? h1 : h0);
  if (rc) return rc;

  const size_t smem = (size_t)CTRL_BYTES + 1024 + (size_t)stages * stage_bytes;
  const int pair_slots = kNumSMs / 2;
  const int units = (int)ceil_div(p.m_tiles, 2) * p.n_tiles;
  dim3 grid(2 * units, splits, 1);
  p.tail_unit0 = 0; p.tail_splits = 1; p.tail_kb_per_split = p.num_kb; p.tail_ws = nullptr;
  const int full = (units / pair_slots) * pair_slots, tail = units - full;
  if (g.balance_ws && splits == 1 && p.n_tiles == 1 && full > 0 && tail > 0 && tail <= pair_slots / 2) {
    int ts = pair_slots / tail;
    if (ts > 16) ts = 16;
    if (ts > p.num_kb / 8) ts = p.num_kb / 8;
    if (ts >= 2) {
      p.tail_kb_per_split = (int)ceil_div(p.num_kb, ts);
      ts = (int)ceil_div(p.num_kb, p.tail_kb_per_split);
      p.tail_unit0 = full; p.tail_splits = ts; p.tail_ws = reinterpret_cast<float*>(g.balance_ws);
      grid.x = 2 * (full + tail * ts);
    }
  }
  prof_begin(PROF_GEMM_TC, stream);
  VS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH(gemm_tn_pair_kernel, grid, NUM_THREADS, smem, stream, tmA, tmB0, tmB1, p);
  if (p.tail_splits > 1) {
    dim3 rg((unsigned)ceil_div((long long)BM * BN / 4, 256), (unsigned)(2 * tail));
    VS_LAUNCH(tail_reduce_kernel, rg, 256, 0, stream, p.tail_ws, 2 * p.tail_unit0, p.tail_splits, 1, BN, p.M, p.N, p.C, p.ldc);
  }
  prof_end(PROF_GEMM_TC, stream);
  if (g.splits_out) *g.splits_out = splits;
  return VS_OK;
}

// ---- dense RRR forward (rrr_fwd_dense_pair_kernel) ----
bool rrr_fwd_dense_supported(const DenseFwdDesc& g) {
  if (g.Npad % 16 != 0 || g.Npad < 16 || g.Npad > 160 || g.N > g.Npad || g.N <= 0) return false;
  if (g.K <= 0 || g.T <= 0 || g.C1 <= 0 || g.K * g.T >= (1ll << 31)) return false;
  if (((uintptr_t)g.Xc & 15) || (g.ldc * 2) % 16 || g.ldc < g.C1) return false;
  if (g.ldu % 64 != 0 || g.ldu < round_up(g.C1, 64) || ((uintptr_t)g.U32 & 15) || ((uintptr_t)g.isd & 15)) return false;
  if (g.ldy % 4 != 0 || g.ldy < g.Npad || ((uintptr_t)g.Y & 15)) return false;
  return true;
}

int rrr_fwd_dense(const DenseFwdDesc& g, cudaStream_t stream) {
  VS_REQUIRE(rrr_fwd_dense_supported(g), VS_ERR_UNSUPPORTED, "dense RRR forward: unsupported shape or operand layout");
  VS_REQUIRE(g.U32 && g.isd && g.V && g.bscale && g.Y && g.Xc, VS_ERR_INVALID, "dense RRR forward: null pointer");
  DenseFwdParams p;
  p.K = (int)g.K; p.T = (int)g.T; p.C1 = (int)g.C1; p.N = (int)g.N; p.Npad = (int)g.Npad;
  p.n_rt = (int)ceil_div(g.K, 256);
  p.num_kb = (int)ceil_div(g.C1, 64);
  int splits = g.splits > 0 ? g.splits : 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (int)ceil_div(p.num_kb, splits);
  splits = (int)ceil_div(p.num_kb, p.kb_per_split);          // no empty run
  p.split_stride = g.split_stride;
  VS_REQUIRE(splits == 1 || g.split_stride >= g.K * g.T * g.ldy, VS_ERR_INVALID, "dense RRR forward: split_stride too small");
  if (g.splits_out) *g.splits_out = splits;
  p.U32 = g.U32; p.isd = g.isd; p.ldu = g.ldu; p.V = g.V; p.bscale = g.bscale; p.Y = g.Y; p.ldy = g.ldy;
  Operand a; a.ptr = g.Xc; a.rows = g.K * g.T; a.k = g.C1; a.ld = g.ldc;
  CUtensorMap tmA;
  int rc = make_map(&tmA, a, false, true, BM);
  if (rc) return rc;
  const int h = p.Npad / 2;
  const size_t stage_bytes = (size_t)FWD_BINS * (A_TILE_BYTES + 2 * (size_t)h * KB_BYTES);
  const size_t smem = (size_t)CTRL_BYTES + 1024 + 2 * stage_bytes;
  VS_REQUIRE(smem <= 227 * 1024, VS_ERR_UNSUPPORTED, "dense RRR forward: tile too large for shared memory");
  const int items = (int)ceil_div(g.T, FWD_BINS) * p.n_rt;
  dim3 grid(2 * (unsigned)items, (unsigned)splits, 1);
  prof_begin(PROF_RRR_FWD, stream);
  VS_CHECK_CUDA(cudaFuncSetAttribute(rrr_fwd_dense_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH(rrr_fwd_dense_pair_kernel, grid, FWD_THREADS, smem, stream, tmA, p);
  prof_end(PROF_RRR_FWD, stream);
  return VS_OK;
}

// ---- dense RRR backward (rrr_bwd_dense_kernel) ----
bool rrr_bwd_dense_supported(const DenseBwdDesc& g) {
  // Default route; VS_RRR_DENSE=0 selects the factorised GEMM-B (read per call: the tests compare both routes in one process).
  // The kernel executes a third of the factorised flops, but its 8 x 26 KB ring cannot keep more bytes in flight than the
  // CTA-pair GEMM does, so the kernel itself takes about the same 0.43 ms at the bench size; the fit still gains ~4 % because
  // the residual operand is a third of the size (epi_f) and the tensor pipe draws less power (DESIGN.md "dense backward").
  const char* e = getenv("VS_RRR_DENSE");
  if (e && e[0] == '0') return false;
  if (g.r != DENSE_R || g.Npad % 16 != 0 || g.Npad < 16 || g.Npad > 160) return false;
  if (g.C1 <= BM || g.T <= 0 || g.K <= 0) return false;
  if (((uintptr_t)g.Xb & 15) || ((uintptr_t)g.R & 15) || (g.ldr * 2) % 16 || g.Kp % 16 || g.Kp < g.K) return false;
  if (g.T * g.Kp + 64 >= (1ll << 31)) return false;
  return true;
}

template <int W4>
static int launch_dense(const CUtensorMap& tmA, const CUtensorMap& tmB, const DenseParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  VS_CHECK_CUDA(cudaFuncSetAttribute(rrr_bwd_dense_kernel<W4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH(rrr_bwd_dense_kernel<W4>, grid, DENSE_THREADS, smem, stream, tmA, tmB, p);
  return VS_OK;
}

template <int HW8>
static int launch_dense_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const DensePairParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  if (p.dvpart) {
    VS_CHECK_CUDA(cudaFuncSetAttribute((rrr_bwd_dense_pair_kernel<HW8, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH((rrr_bwd_dense_pair_kernel<HW8, true>), grid, DENSE_PAIR_THREADS, smem, stream, tmA, tmB, p);
    return VS_OK;
  }
  VS_CHECK_CUDA(cudaFuncSetAttribute((rrr_bwd_dense_pair_kernel<HW8, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VS_LAUNCH((rrr_bwd_dense_pair_kernel<HW8, false>), grid, DENSE_PAIR_THREADS, smem, stream, tmA, tmB, p);
  return VS_OK;
}

static int rrr_bwd_dense_pair(const DenseBwdDesc& g, cudaStream_t stream) {
  DensePairParams p;
  p.C1 = (int)g.C1; p.m_tiles = (int)ceil_div(g.C1, BM); p.Npad = (int)g.Npad; p.T = (int)g.T; p.K = (int)g.K; p.Kp = (int)g.Kp;
  p.nb = (int)ceil_div(g.K, 64);
  p.f16 = g.f16 ? 1 : 0;
  p.V = g.V; p.G = g.G; p.ldg = g.ldg;
  p.r_planes = g.r_planes > 1 ? 2 : 1;
  p.scaleT = g.scaleT; p.ldt = g.ldt;
  p.U32 = g.dv_U32; p.ldu = g.dv_ldu; p.N = (int)g.dv_N; p.dvpart = g.dvpart;
  VS_REQUIRE(!g.dvpart || (g.dv_U32 && g.dv_ldu >= g.C1 && g.dv_N > 0 && g.dv_N <= g.Npad), VS_ERR_INVALID, "dense RRR dV pass: bad U operand");
  VS_REQUIRE(!g.scaleT || (g.ldt % 4 == 0 && g.ldt >= g.T && ((uintptr_t)g.scaleT & 15) == 0), VS_ERR_INVALID,
             "dense RRR backward: the scale table needs a 16-byte aligned base and a pitch that is a multiple of 4 >= T");
  p.vbytes = (int)round_up((long long)DENSE_R * g.T * 4, 1024);
  const int h = p.Npad / 2;
  const int stage_bytes = A_TILE_BYTES + p.r_planes * h * KB_BYTES;
  int stages = (227 * 1024 - CTRL_BYTES - 1024 - p.vbytes) / stage_bytes;
  if (stages > 8) stages = 8;
  VS_REQUIRE(stages >= 2, VS_ERR_UNSUPPORTED, "dense RRR backward: too many time bins for shared memory");
  p.stages = stages;
  int tcols = 32;
  while (tcols < DENSE_R * p.Npad) tcols <<= 1;
  p.tmem_cols = tcols;
  Operand a; a.ptr = g.Xb; a.rows = g.C1; a.k = g.T * g.Kp; a.ld = g.ldr;
  Operand b; b.ptr = g.R; b.rows = g.Npad; b.k = g.T * g.Kp; b.ld = g.ldr;
  if (p.r_planes > 1) { b.planes = 2; b.plane_stride = g.r_plane_stride; }
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, a, false, g.f16, BM);
  if (rc) return rc;
  rc = make_map(&tmB, b, false, g.f16, h);
  if (rc) return rc;
  const size_t smem = (size_t)CTRL_BYTES + 1024 + p.vbytes + (size_t)stages * stage_bytes;
  dim3 grid(2 * (unsigned)ceil_div(p.m_tiles, 2), 1, 1);
  prof_begin(p.dvpart ? PROF_RRR_DV : PROF_RRR_BWD, stream);
  switch (p.Npad / 16) {
    case 1: rc = launch_dense_pair<1>(tmA, tmB, p, grid, smem, stream); break;
    case 2: rc = launch_dense_pair<2>(tmA, tmB, p, grid, smem, stream); break;
    case 3: rc = launch_dense_pair<3>(tmA, tmB, p, grid, smem, stream); break;
    case 4: rc = launch_dense_pair<4>(tmA, tmB, p, grid, smem, stream); break;
    case 5: rc = launch_dense_pair<5>(tmA, tmB, p, grid, smem, stream); break;
    case 6: rc = launch_dense_pair<6>(tmA, tmB, p, grid, smem, stream); break;
    case 7: rc = launch_dense_pair<7>(tmA, tmB, p, grid, smem, stream); break;
    case 8: rc = launch_dense_pair<8>(tmA, tmB, p, grid, smem, stream); break;
    case 9: rc = launch_dense_pair<9>(tmA, tmB, p, grid, smem, stream); break;
    default: rc = launch_dense_pair<10>(tmA, tmB, p, grid, smem, stream); break;
  }
  prof_end(p.dvpart ? PROF_RRR_DV : PROF_RRR_BWD, stream);
  return rc;
}

int rrr_bwd_dense(const DenseBwdDesc& g, cudaStream_t stream) {
  VS_REQUIRE(rrr_bwd_dense_supported(g), VS_ERR_UNSUPPORTED, "dense RRR backward: unsupported shape");
  {
    const char* e = getenv("VS_RRR_DENSE");   // 1 = half-split kernel (accumulators in registers), default = CTA-pair kernel
    if (!(e && e[0] == '1') || g.dvpart || g.r_planes > 1) return rrr_bwd_dense_pair(g, stream);
  }
  DenseParams p;
  p.C1 = (int)g.C1; p.Npad = (int)g.Npad; p.T = (int)g.T; p.K = (int)g.K; p.Kp = (int)g.Kp;
  p.nb = (int)ceil_div(g.K, 64);
  p.nc = p.Npad / 2;
  p.mma_n = (int)round_up(p.nc, 16);
  p.f16 = g.f16 ? 1 : 0;
  p.V = g.V; p.G = g.G; p.ldg = g.ldg;
  p.vbytes = (int)round_up((long long)DENSE_R * g.T * 4, 1024);
  {
    const char* e = getenv("VS_DENSE_FLAGS");
    p.flags = e ? atoi(e) : 0;
    const char* f = getenv("VS_DENSE_PREFETCH");
    p.prefetch = f ? atoi(f) : 24;
  }
  const int stage_bytes = A_TILE_BYTES + p.mma_n * KB_BYTES;
  int stages = (227 * 1024 - CTRL_BYTES - 1024 - p.vbytes) / stage_bytes;
  if (stages > 8) stages = 8;
  VS_REQUIRE(stages >= 2, VS_ERR_UNSUPPORTED, "dense RRR backward: too many time bins for shared memory");
  p.stages = stages;
  Operand a; a.ptr = g.Xb; a.rows = g.C1; a.k = g.T * g.Kp; a.ld = g.ldr;
  Operand b; b.ptr = g.R; b.rows = g.Npad; b.k = g.T * g.Kp; b.ld = g.ldr;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, a, false, g.f16, BM);
  if (rc) return rc;
  rc = make_map(&tmB, b, false, g.f16, p.mma_n);
  if (rc) return rc;
  const size_t smem = (size_t)CTRL_BYTES + 1024 + p.vbytes + (size_t)stages * stage_bytes;
  dim3 grid(2 * (unsigned)ceil_div(g.C1, BM), 1, 1);
  prof_begin(PROF_RRR_BWD, stream);
  switch (p.Npad / 16) {
    case 1: rc = launch_dense<1>(tmA, tmB, p, grid, smem, stream); break;
    case 2: rc = launch_dense<2>(tmA, tmB, p, grid, smem, stream); break;
    case 3: rc = launch_dense<3>(tmA, tmB, p, grid, smem, stream); break;
    case 4: rc = launch_dense<4>(tmA, tmB, p, grid, smem, stream); break;
    case 5: rc = launch_dense<5>(tmA, tmB, p, grid, smem, stream); break;
    case 6: rc = launch_dense<6>(tmA, tmB, p, grid, smem, stream); break;
    case 7: rc = launch_dense<7>(tmA, tmB, p, grid, smem, stream); break;
    case 8: rc = launch_dense<8>(tmA, tmB, p, grid, smem, stream); break;
    case 9: rc = launch_dense<9>(tmA, tmB, p, grid, smem, stream); break;
    default: rc = launch_dense<10>(tmA, tmB, p, grid, smem, stream); break;
  }
  prof_end(PROF_RRR_BWD, stream);
  return rc;
}

int gemm_tn(const GemmDesc& g, cudaStream_t stream) {
  VS_REQUIRE(gemm_supported(g), VS_ERR_UNSUPPORTED, "tcgen05 GEMM: unsupported operand layout/shape");
  const int BN = g.BN > 0 ? g.BN : pick_bn(g.N);
  VS_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= kMaxBN, VS_ERR_UNSUPPORTED, "tcgen05 GEMM: bad BN %d", BN);
  if (!g.tf32 && pair_enabled() && g.M > BM && BN >= 32) return gemm_tn_pair(g, BN, stream);
  // TMA boxes are limited to 256 rows: split the B tile into equal boxes of a multiple of 8 rows
  int n_tb = 1;
  while (BN / n_tb > 256 || BN % n_tb != 0 || (BN / n_tb) % 8 != 0) {
    ++n_tb;
    VS_REQUIRE(n_tb <= 8, VS_ERR_UNSUPPORTED, "tcgen05 GEMM: cannot box BN=%d", BN);
  }
  const int elems_per_kb = g.tf32 ? 32 : 64;
  KParams p;
  p.M = (int)g.M; p.N = (int)g.N; p.BN = BN; p.tb_rows = BN / n_tb; p.n_tb = n_tb;
  p.num_kb = (int)ceil_div(g.K, elems_per_kb);
  int splits = g.splits > 0 ? g.splits : 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (int)ceil_div(p.num_kb, splits);
  splits = (int)ceil_div(p.num_kb, p.kb_per_split);
  p.n_pass = g.n_pass;
  for (int i = 0; i < kMaxPass; ++i) { p.pa[i] = i < g.n_pass ? g.pa[i] : 0; p.pb[i] = i < g.n_pass ? g.pb[i] : 0; }
  // skinny tiles (first Linear layer: N = 16, pitch 7.8 MB): group 4 k-blocks per stage so that every operand row is read in
  // 512 contiguous bytes instead of 128 (DRAM page locality); VS_GEMM_KBGROUP overrides
  int kb_group = (BN <= 64 && p.kb_per_split >= 64) ? 4 : 1;
  { const char* e = getenv("VS_GEMM_KBGROUP"); if (e && atoi(e) >= 1 && atoi(e) <= 8) kb_group = atoi(e); }
  if (kb_group > 1 && (A_TILE_BYTES + BN * KB_BYTES) * kb_group * 2 > 220 * 1024) kb_group = 1;
  p.kb_group = kb_group;
  const int stage_bytes = (A_TILE_BYTES + BN * KB_BYTES) * kb_group;
  const int max_smem = 227 * 1024;
  int stages = (max_smem - CTRL_BYTES - 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  VS_REQUIRE(stages >= 2, VS_ERR_UNSUPPORTED, "tcgen05 GEMM: tile too large for shared memory");
  p.stages = stages;
  int tcols = 32;
  while (tcols < BN) tcols <<= 1;
  p.tmem_cols = tcols;
  p.f16 = g.f16 ? 1 : 0;
  p.C = g.C; p.ldc = g.ldc; p.split_stride = g.split_stride;
  VS_REQUIRE(splits == 1 || g.split_stride >= g.M * g.ldc, VS_ERR_INVALID, "split-K needs split_stride >= M*ldc");

  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, g.A, g.tf32, g.f16, BM);
  if (rc) return rc;
  rc = make_map(&tmB, g.B, g.tf32, g.f16, p.tb_rows);
  if (rc) return rc;

  const size_t smem = (size_t)CTRL_BYTES + 1024 + (size_t)stages * stage_bytes;
  const int m_tiles = (int)ceil_div(g.M, BM), n_tiles = (int)ceil_div(g.N, BN);
  dim3 grid(m_tiles * n_tiles, splits, 1);
  // tail-wave balancing: one CTA per SM is resident (the ring takes the whole shared memory), so a grid of
  // full*148 + tail tiles costs full+1 waves; split the tail tiles along K across the SMs the last wave leaves idle
  p.tail_cta0 = 0; p.tail_splits = 1; p.tail_kb_per_split = p.num_kb; p.tail_ws = nullptr;
  const int tiles = m_tiles * n_tiles;
  const int full = (tiles / kNumSMs) * kNumSMs, tail = tiles - full;
  if (g.balance_ws && splits == 1 && full > 0 && tail > 0 && tail <= kNumSMs / 2) {
    int ts = kNumSMs / tail;
    if (ts > 16) ts = 16;
    if (ts > p.num_kb / 8) ts = p.num_kb / 8;             // keep >= 8 k-blocks per CTA
    if (ts >= 2) {
      p.tail_kb_per_split = (int)ceil_div(p.num_kb, ts);
      ts = (int)ceil_div(p.num_kb, p.tail_kb_per_split);
      p.tail_cta0 = full; p.tail_splits = ts; p.tail_ws = reinterpret_cast<float*>(g.balance_ws);
      grid.x = full + tail * ts;
    }
  }
  prof_begin(PROF_GEMM_TC, stream);
  if (g.tf32) {
    VS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(gemm_tn_kernel<true>, grid, NUM_THREADS, smem, stream, tmA, tmB, p);
  } else {
    VS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_LAUNCH(gemm_tn_kernel<false>, grid, NUM_THREADS, smem, stream, tmA, tmB, p);
  }
  if (p.tail_splits > 1) {
    dim3 rg((unsigned)ceil_div((long long)BM * BN / 4, 256), (unsigned)tail);
    VS_LAUNCH(tail_reduce_kernel, rg, 256, 0, stream, p.tail_ws, p.tail_cta0, p.tail_splits, n_tiles, BN, p.M, p.N, p.C, p.ldc);
  }
  prof_end(PROF_GEMM_TC, stream);
  if (g.splits_out) *g.splits_out = splits;
  return VS_OK;
}

}  // namespace tc
}  // namespace vs
