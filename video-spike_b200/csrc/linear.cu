// Linear layers and the whole `Linear` MLP train step (src/model/linear.py, src/trainer/base.py:147-154).
#include <cuda_bf16.h>
#include "common.cuh"
#include "gemm.h"
#include "smallbatch.h"

namespace vs {
int launch_relu_mask(const float* dy, const float* y, float* out, long long n, cudaStream_t st);
int launch_colsum(const float* dy, float* db, long long batch, long long out_dim, cudaStream_t st);

// the tall contraction (K = pixels) goes to the tensor cores; everything else is launch-bound
constexpr long long kBigK = 4096;

static bool use_tc(int engine, long long batch, long long in_dim, long long out_dim, const float* W) {
  if (engine == VS_ENGINE_SIMT) return false;
  const bool ok = in_dim >= kBigK && in_dim % 4 == 0 && (((uintptr_t)W) & 15) == 0;
  return ok;
}

static int tc_splits(long long out_dim, long long batch) {
  const long long tiles = ceil_div(out_dim, 128) * ceil_div(batch, tc::pick_bn(batch));
  long long s = kNumSMs / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  return (int)s;
}

// ---- large-batch weight gradient of the tall layer on the tensor cores -----------------------------------------------
// dW[o, i] = sum_b g[b, o] x[b, i]  (M = out, N = pixels, contraction = batch).  Operands for the TN engine:
//   A = g^T   (out x Bpad) as TWO bf16 residual planes (16 significant bits: g is a gradient, not a pixel)
//   B = x^T   (pixels x Bpad) bf16 -- exact for uint8 frames (0..255 fit in 8 significant bits)
// Both are produced by small transpose kernels into the workspace; C = dW (out x pixels) fp32 is written once.
constexpr long long kTcBatchMin = 33;     // below: the register kernels (smallbatch.cu / dw_adamw_kernel) are HBM-bound already

__device__ __forceinline__ uint16_t to_bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

// xT[i][b] = bf16(x[b][i]); 32 x 32 tiles through shared memory, both sides coalesced
template <bool kU8>
__global__ void __launch_bounds__(256) xT_bf16_kernel(const uint8_t* __restrict__ xu, const float* __restrict__ xf, long long batch,
                                                      long long in_dim, long long bpad, uint16_t* __restrict__ xT) {
  __shared__ uint16_t tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32, b0 = (long long)blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int r = ly; r < 32; r += 8) {
    const long long b = b0 + r, i = i0 + lx;
    float v = 0.f;
    if (b < batch && i < in_dim) v = kU8 ? (float)xu[b * in_dim + i] : xf[b * in_dim + i];
    tile[r][lx] = to_bf16_bits(v);
  }
  __syncthreads();
  for (int r = ly; r < 32; r += 8) {
    const long long i = i0 + r, b = b0 + lx;
    if (i < in_dim && b < bpad) xT[i * bpad + b] = tile[lx][r];
  }
}
// gT planes: gT[p][o][b] = plane p of g[b][o]
__global__ void gT_planes_kernel(const float* __restrict__ g, long long batch, long long out_dim, long long bpad,
                                 uint16_t* __restrict__ gT) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out_dim * bpad) return;
  const long long o = idx / bpad, b = idx % bpad;
  const float v = b < batch ? g[b * out_dim + o] : 0.f;
  const uint16_t hi = to_bf16_bits(v);
  const float lo = v - __uint_as_float((uint32_t)hi << 16);
  gT[idx] = hi;
  gT[out_dim * bpad + idx] = to_bf16_bits(lo);
}

static size_t dw_tc_workspace(long long batch, long long in_dim, long long out_dim) {
  const long long bpad = round_up(batch, 64);
  return (size_t)round_up(in_dim * bpad * 2, 1024) + (size_t)round_up(2 * out_dim * bpad * 2, 1024) + 1024;
}
static bool dw_tc_ok(long long batch, long long in_dim, long long out_dim, const float* dW) {
  return batch >= kTcBatchMin && in_dim >= kBigK && in_dim % 4 == 0 && (((uintptr_t)dW) & 15) == 0;
}
static int dw_tc(const float* g, const float* x_f32, const uint8_t* x_u8, float* dW, long long batch, long long in_dim,
                 long long out_dim, void* workspace, cudaStream_t st) {
  const long long bpad = round_up(batch, 64);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  uint16_t* xT = reinterpret_cast<uint16_t*>(base);
  uint16_t* gT = reinterpret_cast<uint16_t*>(base + round_up(in_dim * bpad * 2, 1024));
  dim3 gx((unsigned)ceil_div(in_dim, 32), (unsigned)ceil_div(bpad, 32));
  if (x_f32) {
    VS_LAUNCH((xT_bf16_kernel<false>), gx, 256, 0, st, nullptr, x_f32, batch, in_dim, bpad, xT);
  } else {
    VS_LAUNCH((xT_bf16_kernel<true>), gx, 256, 0, st, x_u8, nullptr, batch, in_dim, bpad, xT);
  }
  VS_LAUNCH(gT_planes_kernel, (unsigned)ceil_div(out_dim * bpad, 256), 256, 0, st, g, batch, out_dim, bpad, gT);
  tc::GemmDesc d;
  d.A.ptr = gT; d.A.rows = out_dim; d.A.k = bpad; d.A.ld = bpad; d.A.planes = 2; d.A.plane_stride = out_dim * bpad;
  d.B.ptr = xT; d.B.rows = in_dim; d.B.k = bpad; d.B.ld = bpad;
  d.M = out_dim; d.N = in_dim; d.K = bpad; d.C = dW; d.ldc = in_dim;
  d.n_pass = 2; d.pa[0] = 1; d.pb[0] = 0; d.pa[1] = 0; d.pb[1] = 0;      // low plane first, dominant product last
  d.BN = 256;
  return tc::gemm_tn(d, st);
}
}  // namespace vs

using namespace vs;

extern "C" size_t vs_linear_fwd_workspace(int64_t batch, int64_t in_dim, int64_t out_dim) {
  size_t ws = 0;
  if (in_dim >= kBigK) {
    const size_t part = (size_t)kNumSMs * (size_t)round_up(out_dim, 128) * (size_t)round_up(batch, 16) * sizeof(float);
    const size_t xs = (size_t)round_up(batch * in_dim, 64) * sizeof(float);
    ws = round_up((long long)part, 256) + xs + 256;
  }
  return ws;
}

extern "C" int vs_linear_fwd(const float* x_f32, const uint8_t* x_u8, const float* W, const float* bias, float* y,
                             int64_t batch, int64_t in_dim, int64_t out_dim, int relu, int engine, void* workspace,
                             size_t workspace_bytes, void* stream) {
  VS_REQUIRE((x_f32 || x_u8) && W && y, VS_ERR_INVALID, "vs_linear_fwd: null pointer");
  VS_REQUIRE(batch > 0 && in_dim > 0 && out_dim > 0, VS_ERR_INVALID, "vs_linear_fwd: empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  const bool big = in_dim >= kBigK;
  if (big) {
    VS_REQUIRE(workspace && workspace_bytes >= vs_linear_fwd_workspace(batch, in_dim, out_dim), VS_ERR_WORKSPACE,
               "vs_linear_fwd: workspace too small (%zu < %zu)", workspace_bytes, vs_linear_fwd_workspace(batch, in_dim, out_dim));
  }
  if (use_tc(engine, batch, in_dim, out_dim, W)) {
    // out^T[o, b] = sum_i W[o,i] * x[b,i]  (W rows on the UMMA M axis, batch on N), split over pixels
    float* part = reinterpret_cast<float*>(workspace);
    const size_t part_bytes = round_up((long long)((size_t)kNumSMs * round_up(out_dim, 128) * round_up(batch, 16) * sizeof(float)), 256);
    const float* xs = x_f32;
    if (!xs) {
      float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + part_bytes);
      int rc = vs_u8_to_f32(x_u8, scratch, batch * in_dim, stream);
      if (rc) return rc;
      xs = scratch;
    }
    VS_REQUIRE((((uintptr_t)xs) & 15) == 0, VS_ERR_INVALID, "vs_linear_fwd: x must be 16-byte aligned");
    const long long ldc = round_up(batch, 16);
    tc::GemmDesc g;
    g.A.ptr = W; g.A.rows = out_dim; g.A.k = in_dim; g.A.ld = in_dim;
    g.B.ptr = xs; g.B.rows = batch; g.B.k = in_dim; g.B.ld = in_dim;
    g.M = out_dim; g.N = batch; g.K = in_dim;
    g.C = part; g.ldc = ldc; g.split_stride = round_up(out_dim, 128) * ldc;
    g.splits = tc_splits(out_dim, batch);
    int splits = 1; g.splits_out = &splits;
    g.tf32 = true;
    int rc = tc::gemm_tn(g, st);
    if (rc) return rc;
    return splitk_reduce_bias_act(part, splits, g.split_stride, ldc, 1, bias, y, batch, out_dim, relu, st);
  }
  VS_REQUIRE(engine != VS_ENGINE_TCGEN05, VS_ERR_UNSUPPORTED, "vs_linear_fwd: shape not supported by the tcgen05 engine");
  // the small layers at training batch sizes: weight-streaming kernel (smallbatch.cu)
  if (!big && x_f32 && sb::fwd_supported(batch, in_dim, out_dim) && ((((uintptr_t)x_f32) | ((uintptr_t)W)) & 15) == 0)
    return sb::fwd(x_f32, W, bias, y, batch, in_dim, out_dim, relu, st);
  simt::GemmDesc g;
  g.A.ptr = x_f32 ? (const void*)x_f32 : (const void*)x_u8; g.A.type = x_f32 ? simt::F32 : simt::U8; g.A.s_i = in_dim; g.A.s_k = 1;
  g.B.ptr = W; g.B.type = simt::F32; g.B.s_i = in_dim; g.B.s_k = 1;
  g.M = batch; g.N = out_dim; g.K = in_dim;
  if (big) {
    float* part = reinterpret_cast<float*>(workspace);
    long long tiles = ceil_div(batch, 64) * ceil_div(out_dim, 64);
    int splits = (int)(4 * kNumSMs / (tiles > 0 ? tiles : 1));
    if (splits < 1) splits = 1;
    if (splits > kNumSMs) splits = kNumSMs;
    g.C = part; g.ldc = out_dim; g.split_stride = batch * out_dim; g.splits = splits;
    int rc = simt::gemm(g, st);
    if (rc) return rc;
    const long long kps = round_up(ceil_div(in_dim, splits), 16);
    const int used = (int)ceil_div(in_dim, kps);
    return splitk_reduce_bias_act(part, used, g.split_stride, 1, out_dim, bias, y, batch, out_dim, relu, st);
  }
  g.C = y; g.ldc = out_dim; g.bias = bias; g.relu = relu;
  return simt::gemm(g, st);
}

extern "C" size_t vs_linear_bwd_workspace(int64_t batch, int64_t in_dim, int64_t out_dim) {
  size_t ws = sb::dx_workspace(batch, in_dim, out_dim);
  if (batch >= kTcBatchMin && in_dim >= kBigK) ws = ws > dw_tc_workspace(batch, in_dim, out_dim) ? ws : dw_tc_workspace(batch, in_dim, out_dim);
  return ws;
}

extern "C" int vs_linear_bwd(const float* dy, const float* y, const float* x_f32, const uint8_t* x_u8, const float* W,
                             float* dy_masked, float* dx, float* dW, float* dbias, int64_t batch, int64_t in_dim,
                             int64_t out_dim, int relu, void* workspace, size_t workspace_bytes, void* stream) {
  VS_REQUIRE(dy, VS_ERR_INVALID, "vs_linear_bwd: dy is null");
  VS_REQUIRE(batch > 0 && in_dim > 0 && out_dim > 0, VS_ERR_INVALID, "vs_linear_bwd: empty shape");
  cudaStream_t st = (cudaStream_t)stream;
  const float* g = dy;
  if (relu) {
    VS_REQUIRE(y && dy_masked, VS_ERR_INVALID, "vs_linear_bwd: relu backward needs y and dy_masked");
    int rc = launch_relu_mask(dy, y, dy_masked, batch * out_dim, st);
    if (rc) return rc;
    g = dy_masked;
  }
  const bool small = sb::supported(batch, in_dim, out_dim) && ((((uintptr_t)x_f32) | ((uintptr_t)W) | ((uintptr_t)dW) | ((uintptr_t)dx)) & 15) == 0 &&
                     (((uintptr_t)x_u8) & 3) == 0;
  if (small && dW && (x_f32 || x_u8)) {
    // dW (and dbias) in one pass over the output, the batch rows held in registers
    int rc = sb::dw_store(g, x_f32, x_u8, dW, dbias, batch, in_dim, out_dim, st);
    if (rc) return rc;
    dW = nullptr; dbias = nullptr;
  }
  if (dbias) {
    int rc = launch_colsum(g, dbias, batch, out_dim, st);
    if (rc) return rc;
  }
  if (dx && small && W && workspace && workspace_bytes >= sb::dx_workspace(batch, in_dim, out_dim) && (((uintptr_t)workspace) & 15) == 0) {
    int rc = sb::dx(g, W, nullptr, dx, batch, in_dim, out_dim, workspace, workspace_bytes, st);
    if (rc) return rc;
    dx = nullptr;
  }
  if (dW && (x_f32 || x_u8) && dw_tc_ok(batch, in_dim, out_dim, dW) && workspace &&
      workspace_bytes >= dw_tc_workspace(batch, in_dim, out_dim)) {
    // large batch on the tall layer: tensor cores (the SIMT engine would be FMA-bound at 2*B flops per weight)
    int rc = dw_tc(g, x_f32, x_u8, dW, batch, in_dim, out_dim, workspace, st);
    if (rc) return rc;
    dW = nullptr;
  }
  if (dW) {
    VS_REQUIRE(x_f32 || x_u8, VS_ERR_INVALID, "vs_linear_bwd: dW needs the layer input");
    simt::GemmDesc d;  // dW[o,i] = sum_b g[b,o] * x[b,i]
    d.A.ptr = g; d.A.type = simt::F32; d.A.s_i = 1; d.A.s_k = out_dim;
    d.B.ptr = x_f32 ? (const void*)x_f32 : (const void*)x_u8; d.B.type = x_f32 ? simt::F32 : simt::U8; d.B.s_i = 1; d.B.s_k = in_dim;
    d.M = out_dim; d.N = in_dim; d.K = batch; d.C = dW; d.ldc = in_dim;
    int rc = simt::gemm(d, st);
    if (rc) return rc;
  }
  if (dx) {
    VS_REQUIRE(W, VS_ERR_INVALID, "vs_linear_bwd: dx needs W");
    simt::GemmDesc d;  // dx[b,i] = sum_o g[b,o] * W[o,i]
    d.A.ptr = g; d.A.type = simt::F32; d.A.s_i = out_dim; d.A.s_k = 1;
    d.B.ptr = W; d.B.type = simt::F32; d.B.s_i = 1; d.B.s_k = in_dim;
    d.M = batch; d.N = in_dim; d.K = out_dim; d.C = dx; d.ldc = in_dim;
    int rc = simt::gemm(d, st);
    if (rc) return rc;
  }
  return VS_OK;
}

// ------------------------------------------------------------------ whole MLP
static int check_net(const vs_mlp* net, int64_t batch, bool train) {
  // gradients are only materialised (gW/gb) on the large-batch route; at batch <= 32 they live in registers
  const bool need_grads = train && batch > 32;
  VS_REQUIRE(net && net->n_layers >= 1 && net->n_layers <= VS_MAX_LAYERS, VS_ERR_INVALID, "vs_mlp: bad layer count");
  VS_REQUIRE(batch > 0, VS_ERR_INVALID, "vs_mlp: empty batch");
  for (int l = 0; l < net->n_layers; ++l) {
    VS_REQUIRE(net->W[l] && net->act[l] && net->dims[l] > 0 && net->dims[l + 1] > 0, VS_ERR_INVALID, "vs_mlp: layer %d incomplete", l);
    if (train) {
      VS_REQUIRE(net->mW[l] && net->vW[l] && net->gact[l], VS_ERR_INVALID, "vs_mlp: layer %d missing optimizer/grad buffers", l);
      VS_REQUIRE(!net->b[l] || (net->mb[l] && net->vb[l]), VS_ERR_INVALID, "vs_mlp: layer %d missing bias moment buffers", l);
      VS_REQUIRE(!need_grads || (net->gW[l] && (!net->b[l] || net->gb[l])), VS_ERR_INVALID, "vs_mlp: layer %d missing gW/gb scratch (batch > 32)", l);
    }
  }
  return VS_OK;
}

extern "C" size_t vs_mlp_workspace(const vs_mlp* net, int64_t batch) {
  if (!net) return 0;
  size_t ws = 0;
  for (int l = 0; l < net->n_layers && l < VS_MAX_LAYERS; ++l) {
    const size_t w = vs_linear_fwd_workspace(batch, net->dims[l], net->dims[l + 1]);
    if (w > ws) ws = w;
    const size_t wb = (l >= 1 || batch > 32) ? vs_linear_bwd_workspace(batch, net->dims[l], net->dims[l + 1]) : 0;
    if (wb > ws) ws = wb;
  }
  return ws;
}

extern "C" int vs_mlp_forward(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32, const float* target,
                              int64_t batch, double* loss_sum, int engine, void* workspace, size_t workspace_bytes,
                              void* stream) {
  int rc = check_net(net, batch, false);
  if (rc) return rc;
  VS_REQUIRE(frames_u8 || x_f32, VS_ERR_INVALID, "vs_mlp_forward: no input");
  const int L = net->n_layers;
  for (int l = 0; l < L; ++l) {
    const float* xin = l == 0 ? x_f32 : net->act[l - 1];
    const uint8_t* xu = l == 0 && !x_f32 ? frames_u8 : nullptr;
    // `engine` selects the engine of the tall (pixel) contraction; the small layers always take the SIMT kernels
    const int eng = net->dims[l] >= kBigK ? engine : VS_ENGINE_AUTO;
    rc = vs_linear_fwd(xin, xu, net->W[l], net->b[l], net->act[l], batch, net->dims[l], net->dims[l + 1], net->relu[l], eng,
                       workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  if (target) {
    VS_REQUIRE(loss_sum, VS_ERR_INVALID, "vs_mlp_forward: loss_sum is null");
    rc = vs_poisson_nll(net->act[L - 1], target, loss_sum, nullptr, batch * net->dims[L], stream);
  }
  return rc;
}

namespace vs {
int launch_bias_act(float* y, const float* bias, long long batch, long long out_dim, int relu, cudaStream_t st);
}

// forward of layers [l0, L): layer l reads act[l-1] (or the frames for l == 0)
static int mlp_forward_from(const vs_mlp* net, int l0, const uint8_t* frames_u8, const float* x_f32, int64_t batch, int engine,
                            void* workspace, size_t workspace_bytes, void* stream) {
  for (int l = l0; l < net->n_layers; ++l) {
    const float* xin = l == 0 ? x_f32 : net->act[l - 1];
    const uint8_t* xu = l == 0 && !x_f32 ? frames_u8 : nullptr;
    const int eng = net->dims[l] >= kBigK ? engine : VS_ENGINE_AUTO;
    int rc = vs_linear_fwd(xin, xu, net->W[l], net->b[l], net->act[l], batch, net->dims[l], net->dims[l + 1], net->relu[l], eng,
                           workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  return VS_OK;
}

static int mlp_backward_update(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32, const float* target, int64_t batch,
                               vs_adamw_hyper h, double* loss_sum, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int vs_mlp_train_step(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32, const float* target,
                                 int64_t batch, vs_adamw_hyper h, double* loss_sum, int engine, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc = check_net(net, batch, true);
  if (rc) return rc;
  VS_REQUIRE((frames_u8 || x_f32) && target && loss_sum, VS_ERR_INVALID, "vs_mlp_train_step: null pointer");
  // forward (M1-M3)
  rc = vs_mlp_forward(net, frames_u8, x_f32, nullptr, batch, nullptr, engine, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return mlp_backward_update(net, frames_u8, x_f32, target, batch, h, loss_sum, workspace, workspace_bytes, stream);
}

// Row-parallel first layer: this rank owns the pixel slice dims[0] of every frame and the matching columns
// of W0 (+ its Adam moments); everything after the first pre-activation is replicated.
//   phase 0: act[0] = frames_slice . W0_slice^T            (no bias, no ReLU)  -> caller all-reduces act[0] (batch x dims[1] fp32)
//   phase 1: act[0] = act(act[0] + b0); layers 1..L-1, loss, backward, AdamW on every local parameter
extern "C" int vs_mlp_train_step_rowpar(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32, const float* target,
                                        int64_t batch, vs_adamw_hyper h, double* loss_sum, int engine, void* workspace,
                                        size_t workspace_bytes, void* stream, int phase) {
  int rc = check_net(net, batch, true);
  if (rc) return rc;
  VS_REQUIRE(frames_u8 || x_f32, VS_ERR_INVALID, "vs_mlp_train_step_rowpar: no input");
  VS_REQUIRE(phase == 0 || phase == 1, VS_ERR_INVALID, "vs_mlp_train_step_rowpar: phase must be 0 or 1");
  if (phase == 0) {
    const int eng = net->dims[0] >= kBigK ? engine : VS_ENGINE_AUTO;
    return vs_linear_fwd(x_f32, x_f32 ? nullptr : frames_u8, net->W[0], nullptr, net->act[0], batch, net->dims[0], net->dims[1], 0, eng,
                         workspace, workspace_bytes, stream);
  }
  VS_REQUIRE(target && loss_sum, VS_ERR_INVALID, "vs_mlp_train_step_rowpar: null pointer");
  rc = launch_bias_act(net->act[0], net->b[0], batch, net->dims[1], net->relu[0], (cudaStream_t)stream);
  if (rc) return rc;
  rc = mlp_forward_from(net, 1, frames_u8, x_f32, batch, engine, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return mlp_backward_update(net, frames_u8, x_f32, target, batch, h, loss_sum, workspace, workspace_bytes, stream);
}

static int mlp_backward_update(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32, const float* target, int64_t batch,
                               vs_adamw_hyper h, double* loss_sum, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  const int L = net->n_layers;
  // loss + dlogits (C1)
  rc = vs_poisson_nll(net->act[L - 1], target, loss_sum, net->gact[L - 1], batch * net->dims[L], stream);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // small-batch route (the reference batch size): masked dx per layer, then dW+AdamW fused per layer
  bool small = batch <= 32;
  for (int l = 0; l < L && small; ++l) small = sb::supported(batch, net->dims[l], net->dims[l + 1]);
  if (small) {
    // backward data (G1) through layers L-1 .. 1 with the PRE-update weights; the ReLU mask of the layer
    // below is folded into the reduction, so gact[l-1] is the gradient w.r.t. that layer's pre-activation
    for (int l = L - 1; l >= 1; --l) {
      rc = sb::dx(net->gact[l], net->W[l], net->relu[l - 1] ? net->act[l - 1] : nullptr, net->gact[l - 1], batch, net->dims[l],
                  net->dims[l + 1], workspace, workspace_bytes, st);
      if (rc) return rc;
    }
    // weight gradient + AdamW (O1), gradient never written: layer 0 from the uint8 frames
    const uint8_t* xu0 = x_f32 ? nullptr : frames_u8;
    const bool fused0 = net->dims[1] * 32 * 4 <= 200 * 1024;
    if (fused0) {
      rc = vs_dw_adamw_fused(net->gact[0], x_f32, xu0, net->W[0], net->mW[0], net->vW[0], net->b[0], net->mb[0], net->vb[0], batch,
                             net->dims[0], net->dims[1], h, stream);
    } else {
      rc = sb::dw_adamw(net->gact[0], x_f32, xu0, net->W[0], net->mW[0], net->vW[0], net->b[0], net->mb[0], net->vb[0], batch,
                        net->dims[0], net->dims[1], h, st);
    }
    if (rc) return rc;
    for (int l = 1; l < L; ++l) {
      rc = sb::dw_adamw(net->gact[l], net->act[l - 1], nullptr, net->W[l], net->mW[l], net->vW[l], net->b[l], net->mb[l], net->vb[l],
                        batch, net->dims[l], net->dims[l + 1], h, st);
      if (rc) return rc;
    }
    return VS_OK;
  }
  // general route (large batches): per-layer backward with materialised gradients
  for (int l = L - 1; l >= 1; --l) {
    VS_REQUIRE(net->gW[l], VS_ERR_INVALID, "vs_mlp_train_step: layer %d has no gW scratch", l);
    rc = vs_linear_bwd(net->gact[l], net->act[l], net->act[l - 1], nullptr, net->W[l], net->gact[l], net->gact[l - 1], net->gW[l],
                       net->b[l] ? net->gb[l] : nullptr, batch, net->dims[l], net->dims[l + 1], net->relu[l], nullptr, 0, stream);
    if (rc) return rc;
  }
  const uint8_t* xu = x_f32 ? nullptr : frames_u8;
  VS_REQUIRE(net->gW[0], VS_ERR_INVALID, "vs_mlp_train_step: large-batch route needs gW[0]");
  rc = vs_linear_bwd(net->gact[0], net->act[0], x_f32, xu, net->W[0], net->gact[0], nullptr, net->gW[0],
                     net->b[0] ? net->gb[0] : nullptr, batch, net->dims[0], net->dims[1], net->relu[0], workspace, workspace_bytes,
                     stream);
  if (rc) return rc;
  for (int l = 0; l < L; ++l) {
    rc = vs_adamw(net->W[l], net->gW[l], net->mW[l], net->vW[l], net->dims[l] * net->dims[l + 1], h, stream);
    if (rc) return rc;
    if (net->b[l]) {
      rc = vs_adamw(net->b[l], net->gb[l], net->mb[l], net->vb[l], net->dims[l + 1], h, stream);
      if (rc) return rc;
    }
  }
  return VS_OK;
}
