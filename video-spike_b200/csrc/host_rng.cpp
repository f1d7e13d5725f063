// Host-side numpy-legacy random stream for the RRR initialisation (R1, src/model/rrr.py:35,42-43):
//   np.random.seed(0);  U = np.random.normal(size=(N, C-1, r)) / np.sqrt(T*r);  V = np.random.normal(size=(r, T)) / ...
// The reference draws ~8 M normals per session from numpy's global RandomState; in numpy that is a scalar loop
// (~0.2 s on the bench host -- longer than the whole GPU fit).  This file reproduces the stream BIT FOR BIT
// (MT19937 as seeded by `RandomState.seed(uint32)`, 53-bit doubles from two words, the polar "legacy_gauss" with its
// cached second value, libm log/sqrt) but splits the work so that only the MT19937 recurrence is sequential:
//   1. raw 32-bit words are produced block-wise (624 per twist),
//   2. candidate pairs are tested and transformed by all host threads; an ordered prefix count gives every accepted
//      pair its place in the output, so the result is identical to the scalar loop.
// Plain C++ (no CUDA): compiled by g++ with -ffp-contract=off (numpy's legacy code is built without FMA contraction).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/vs_b200.h"

namespace {

constexpr int kN = 624, kM = 397;

struct Rng {
  uint32_t mt[kN];
  int32_t pos;        // next word of the current block (kN = block exhausted), numpy's `pos`
  int32_t has_gauss;
  double gauss;
};
static_assert(sizeof(Rng) <= VS_HOST_RNG_STATE_BYTES, "state blob too small");

inline void twist(uint32_t* mt) {
  const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MA = 0x9908b0dfu;
  int kk = 0;
  for (; kk < kN - kM; ++kk) {
    const uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO);
    mt[kk] = mt[kk + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & MA);
  }
  for (; kk < kN - 1; ++kk) {
    const uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO);
    mt[kk] = mt[kk + (kM - kN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MA);
  }
  const uint32_t y = (mt[kN - 1] & UP) | (mt[0] & LO);
  mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MA);
}

inline uint32_t temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// numpy random_double / legacy_double: 53 bits from two consecutive words
inline double to_double(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0; }

struct Attempt {
  double x1, x2, r2;
  bool ok;
};
inline Attempt attempt(const uint32_t* w) {
  Attempt a;
  a.x1 = 2.0 * to_double(w[0], w[1]) - 1.0;
  a.x2 = 2.0 * to_double(w[2], w[3]) - 1.0;
  a.r2 = a.x1 * a.x1 + a.x2 * a.x2;
  a.ok = !(a.r2 >= 1.0 || a.r2 == 0.0);
  return a;
}

}  // namespace

extern "C" int vs_host_rng_seed(void* state, uint32_t seed) {
  if (!state) return VS_ERR_INVALID;
  Rng* g = reinterpret_cast<Rng*>(state);
  // numpy _legacy_seeding(int) -> mt19937_seed(state, seed): Knuth's initialiser, pos = 624
  g->mt[0] = seed;
  for (int i = 1; i < kN; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->pos = kN;
  g->has_gauss = 0;
  g->gauss = 0.0;
  return VS_OK;
}

extern "C" int vs_host_rng_get_state(const void* state, uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* cached) {
  if (!state || !key624 || !pos || !has_gauss || !cached) return VS_ERR_INVALID;
  const Rng* g = reinterpret_cast<const Rng*>(state);
  memcpy(key624, g->mt, sizeof(g->mt));
  *pos = g->pos;
  *has_gauss = g->has_gauss;
  *cached = g->gauss;
  return VS_OK;
}

extern "C" int vs_host_rng_set_state(void* state, const uint32_t* key624, int32_t pos, int32_t has_gauss, double cached) {
  if (!state || !key624 || pos < 0 || pos > kN) return VS_ERR_INVALID;
  Rng* g = reinterpret_cast<Rng*>(state);
  memcpy(g->mt, key624, sizeof(g->mt));
  g->pos = pos;
  g->has_gauss = has_gauss ? 1 : 0;
  g->gauss = cached;
  return VS_OK;
}

// The stream is handled in BLOCKS of 624 words: block 0 is the generator's current (already twisted) array, block
// j+1 = twist(block j).  Word W of the stream is tempered(block[W / 624][W % 624]); the words before the current
// position `pos` of block 0 are already consumed.  Attempt i of the polar method uses words W0+4i .. W0+4i+3.
// Only the twist recurrence is sequential (a few ns per block word); it is run once to drop a snapshot every kSnap
// blocks, after which every chunk of kSnap blocks can be regenerated, tested and transformed independently.
constexpr int kSnap = 64;

struct Chunker {
  const std::vector<std::vector<uint32_t>>* snaps;   // snaps[k] = untempered array of block k*kSnap
  size_t W0, attempts;
  // attempts whose first word lies in chunk k
  void range(size_t k, size_t* lo, size_t* hi) const {
    const size_t w_lo = k * kSnap * (size_t)kN, w_hi = (k + 1) * kSnap * (size_t)kN;
    *lo = w_lo <= W0 ? 0 : (w_lo - W0 + 3) / 4;
    *hi = w_hi <= W0 ? 0 : (w_hi - W0 + 3) / 4;
    if (*lo > attempts) *lo = attempts;
    if (*hi > attempts) *hi = attempts;
  }
  // tempered words of blocks [k*kSnap, (k+1)*kSnap] (one extra block: an attempt may straddle the chunk end)
  void words(size_t k, uint32_t* buf) const {
    uint32_t mt[kN];
    memcpy(mt, (*snaps)[k].data(), sizeof(mt));
    for (int j = 0; j <= kSnap; ++j) {
      if (j) twist(mt);
      for (int i = 0; i < kN; ++i) buf[(size_t)j * kN + i] = temper(mt[i]);
    }
  }
};

extern "C" int vs_host_rng_normal(void* state, int64_t n, double divisor, double* out_host, int threads) {
  if (!state || n < 0 || (n > 0 && !out_host)) return VS_ERR_INVALID;
  Rng* g = reinterpret_cast<Rng*>(state);
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  if (threads > 64) threads = 64;
  int64_t filled = 0;
  if (n > 0 && g->has_gauss) {   // legacy_gauss hands out the cached second value first
    out_host[filled++] = g->gauss / divisor;
    g->has_gauss = 0;
    g->gauss = 0.0;
  }
  while (filled < n) {
    const int64_t room = n - filled;
    const size_t need_pairs = (size_t)((room + 1) / 2);
    // acceptance probability of the polar method is pi/4; small head-room, the loop tops up if it falls short
    const size_t attempts = (size_t)((double)need_pairs / 0.7853981633974483 * 1.01) + 64;
    const size_t W0 = (size_t)g->pos;
    const size_t n_blocks = (W0 + 4 * attempts + kN - 1) / kN;          // blocks that hold the candidate words
    const size_t n_chunks = (n_blocks + kSnap - 1) / kSnap;
    // sequential part: twist through the blocks, snapshot every kSnap
    std::vector<std::vector<uint32_t>> snaps(n_chunks + 1);
    {
      uint32_t mt[kN];
      memcpy(mt, g->mt, sizeof(mt));
      for (size_t blk = 0; blk <= n_chunks * kSnap; ++blk) {
        if (blk) twist(mt);
        if (blk % kSnap == 0) snaps[blk / kSnap].assign(mt, mt + kN);
      }
    }
    Chunker ck{&snaps, W0, attempts};
    // pass 1: accepted attempts per chunk
    std::vector<size_t> counts(n_chunks + 1, 0);
    const int nthr = (int)std::min<size_t>((size_t)threads, n_chunks);
    // chunks are handed out dynamically (a shared counter): on a busy host a descheduled thread simply takes fewer
    auto for_chunks = [&](auto fn) {
      if (nthr <= 1) { std::vector<uint32_t> buf((size_t)(kSnap + 1) * kN); for (size_t k = 0; k < n_chunks; ++k) fn(k, buf.data()); return; }
      std::atomic<size_t> next{0};
      std::vector<std::thread> pool;
      for (int t = 0; t < nthr; ++t)
        pool.emplace_back([&] {
          std::vector<uint32_t> buf((size_t)(kSnap + 1) * kN);
          for (size_t k = next.fetch_add(1); k < n_chunks; k = next.fetch_add(1)) fn(k, buf.data());
        });
      for (auto& th : pool) th.join();
    };
    for_chunks([&](size_t k, uint32_t* buf) {
      size_t lo, hi;
      ck.range(k, &lo, &hi);
      if (lo >= hi) return;
      ck.words(k, buf);
      const size_t base = k * kSnap * (size_t)kN;
      size_t c = 0;
      for (size_t i = lo; i < hi; ++i) c += attempt(buf + (W0 + 4 * i - base)).ok ? 1 : 0;
      counts[k + 1] = c;
    });
    for (size_t k = 0; k < n_chunks; ++k) counts[k + 1] += counts[k];
    const size_t accepted = counts[n_chunks];
    const size_t use_pairs = std::min(accepted, need_pairs);
    // pass 2: every chunk writes its accepted pairs at their global rank
    double* dst = out_host + filled;
    std::vector<size_t> last_idx(n_chunks, (size_t)-1);
    std::vector<double> tails(n_chunks, 0.0);
    std::vector<char> tail_flags(n_chunks, 0);
    for_chunks([&](size_t k, uint32_t* buf) {
      size_t lo, hi, rank = counts[k];
      ck.range(k, &lo, &hi);
      if (lo >= hi || rank >= use_pairs) return;
      ck.words(k, buf);
      const size_t base = k * kSnap * (size_t)kN;
      for (size_t i = lo; i < hi && rank < use_pairs; ++i) {
        const Attempt at = attempt(buf + (W0 + 4 * i - base));
        if (!at.ok) continue;
        const double f = sqrt(-2.0 * log(at.r2) / at.r2);
        const int64_t o = 2 * (int64_t)rank;
        dst[o] = (f * at.x2) / divisor;                          // legacy_gauss returns f*x2 first ...
        if (o + 1 < room) dst[o + 1] = (f * at.x1) / divisor;    // ... and caches f*x1 for the next call
        else { tails[k] = f * at.x1; tail_flags[k] = 1; }
        if (rank + 1 == use_pairs) last_idx[k] = i;
        ++rank;
      }
    });
    filled += std::min<int64_t>(room, 2 * (int64_t)use_pairs);
    // position the generator right after the last consumed word
    size_t used_attempts = attempts;                             // whole batch consumed if it fell short
    if (accepted >= need_pairs) {
      for (size_t k = 0; k < n_chunks; ++k)
        if (last_idx[k] != (size_t)-1) used_attempts = last_idx[k] + 1;
      for (size_t k = 0; k < n_chunks; ++k)
        if (tail_flags[k]) { g->has_gauss = 1; g->gauss = tails[k]; }
    }
    const size_t W_end = W0 + 4 * used_attempts;
    // numpy leaves (block, pos) with pos in 1..624: a fully consumed block is NOT twisted until the next draw
    size_t blk = W_end / kN, pos = W_end % kN;
    if (pos == 0 && W_end > 0) { blk -= 1; pos = kN; }
    const size_t k = blk / kSnap;
    memcpy(g->mt, snaps[k].data(), sizeof(g->mt));
    for (size_t j = k * kSnap; j < blk; ++j) twist(g->mt);
    g->pos = (int32_t)pos;
  }
  return VS_OK;
}
