// C-ABI plumbing: version, error string, launch counter, GEMM test hook.
#include "common.cuh"
#include "gemm.h"

namespace vs {
std::atomic<long long> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// ---- profiling ring ------------------------------------------------------------------------
namespace {
constexpr int kProfMax = 8192;
struct ProfState {
  bool on = false;
  int n = 0;
  int open_idx = -1;
  cudaEvent_t ev[kProfMax][2];
  int tag[kProfMax];
  int created = 0;
} g_prof;
}  // namespace

void prof_begin(int tag, cudaStream_t st) {
  if (!g_prof.on || g_prof.n >= kProfMax) { g_prof.open_idx = -1; return; }
  const int i = g_prof.n;
  if (i >= g_prof.created) {
    if (cudaEventCreate(&g_prof.ev[i][0]) != cudaSuccess || cudaEventCreate(&g_prof.ev[i][1]) != cudaSuccess) { g_prof.open_idx = -1; return; }
    g_prof.created = i + 1;
  }
  g_prof.tag[i] = tag;
  cudaEventRecord(g_prof.ev[i][0], st);
  g_prof.open_idx = i;
}
void prof_end(int tag, cudaStream_t st) {
  if (g_prof.open_idx < 0) return;
  cudaEventRecord(g_prof.ev[g_prof.open_idx][1], st);
  g_prof.n = g_prof.open_idx + 1;
  g_prof.open_idx = -1;
}
}  // namespace vs

using namespace vs;

extern "C" void vs_profile_enable(int on) { g_prof.on = on != 0; g_prof.n = 0; g_prof.open_idx = -1; }
// Sums the recorded intervals of `tag` (synchronises on the recorded events).  Returns the count.
extern "C" int64_t vs_profile_read(int tag, double* total_ms, double* min_ms, double* max_ms) {
  double tot = 0.0, mn = 1e30, mx = 0.0;
  int64_t cnt = 0;
  for (int i = 0; i < g_prof.n; ++i) {
    if (g_prof.tag[i] != tag) continue;
    float ms = 0.f;
    if (cudaEventSynchronize(g_prof.ev[i][1]) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&ms, g_prof.ev[i][0], g_prof.ev[i][1]) != cudaSuccess) continue;
    tot += ms; mn = ms < mn ? ms : mn; mx = ms > mx ? ms : mx; ++cnt;
  }
  if (total_ms) *total_ms = tot;
  if (min_ms) *min_ms = cnt ? mn : 0.0;
  if (max_ms) *max_ms = mx;
  return cnt;
}

extern "C" int vs_version(void) { return VS_ABI_VERSION; }
extern "C" const char* vs_last_error(void) { return err_buf(); }
extern "C" int64_t vs_launch_count(void) { return g_launches.load(); }
extern "C" void vs_launch_count_reset(void) { g_launches.store(0); }

extern "C" int vs_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

extern "C" int vs_gemm_tn(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                          int64_t ldc, int dtype, int engine, void* stream) {
  VS_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, VS_ERR_INVALID, "vs_gemm_tn: null pointer or empty shape");
  VS_REQUIRE(dtype >= 0 && dtype <= 2, VS_ERR_INVALID, "vs_gemm_tn: dtype must be 0 (bf16), 1 (tf32) or 2 (f16)");
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == VS_ENGINE_SIMT) {
    simt::GemmDesc g;
    g.A.ptr = A; g.A.type = dtype == 0 ? simt::BF16 : (dtype == 2 ? simt::F16 : simt::F32); g.A.s_i = lda; g.A.s_k = 1;
    g.B.ptr = B; g.B.type = g.A.type; g.B.s_i = ldb; g.B.s_k = 1;
    g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc;
    return simt::gemm(g, st);
  }
  tc::GemmDesc g;
  g.A.ptr = A; g.A.rows = M; g.A.k = K; g.A.ld = lda;
  g.B.ptr = B; g.B.rows = N; g.B.k = K; g.B.ld = ldb;
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.tf32 = dtype == 1; g.f16 = dtype == 2;
  return tc::gemm_tn(g, st);
}
