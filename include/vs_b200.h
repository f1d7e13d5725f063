/*
 * vs_b200.h -- C-ABI of libvs_b200.so, the B200 (sm_100a) engine behind the
 * video->spike encoder hot path of PPWangyc/video-spike.
 *
 * The reference is pure Python/PyTorch and has NO FFI of its own; each
 * entry point below therefore names the reference Python code whose arithmetic it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer owned by the caller (PyTorch caching
 *     allocator) unless the name ends in _host; the library allocates nothing persistent
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*), no
 *     internal synchronisation unless stated
 *   - return 0 on success, non-zero VS_ERR_* otherwise; vs_last_error() gives the
 *     thread-local message.  No exceptions cross the ABI.
 *   - one process per GPU; distinct streams may be used from distinct threads
 */
#ifndef VS_B200_H
#define VS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_ABI_VERSION 7

enum {
  VS_OK = 0,
  VS_ERR_INVALID = 1,     /* bad argument (null pointer, shape, alignment) */
  VS_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
  VS_ERR_UNSUPPORTED = 3, /* shape outside what the kernels cover */
  VS_ERR_WORKSPACE = 4    /* caller workspace too small */
};

/* GEMM engine selector for the entry points that take `engine` */
enum {
  VS_ENGINE_AUTO = 0,    /* tcgen05 where the shape allows, else SIMT */
  VS_ENGINE_SIMT = 1,    /* CUDA-core fp32 FMA path (bring-up / small layers / cross-check) */
  VS_ENGINE_TCGEN05 = 2  /* force the tensor-core path; VS_ERR_UNSUPPORTED if it cannot run */
};

int vs_version(void);
const char* vs_last_error(void);
/* 1 if the current device is compute capability 10.x (the only supported target) */
int vs_device_ok(void);

/* ------------------------------------------------------------------ loader (L1/L2)
 * src/loader/base.py:39,54 (`.float()` on uint8 frames, values stay 0..255) followed by
 * src/trainer/base.py:64-67 (`flatten(1)`): frames (B, T, 1, H, W) uint8, contiguous, is
 * already the flattened (B, D) matrix; the cast is the only arithmetic.  Bit-exact. */
int vs_u8_to_f32(const uint8_t* frames, float* out, int64_t n, void* stream);
int vs_u8_to_bf16(const uint8_t* frames, uint16_t* out_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------ trial windows (L0)
 * src/utils/ibl_data_utils.py:958-967 defines each trial as frames [start, start + 120) of the session video with
 * start = searchsorted(timestamps, t0) (host, utils/dataset_utils.py::load_video_index).  This cuts the windows on
 * the device: frames (n_frames, row_bytes) uint8, start_idx (n_trials) int64 -> out (n_trials, frames_per_trial,
 * row_bytes).  Byte copy, bit-exact; frames past the end of the video read as 0.                              */
int vs_gather_windows(const uint8_t* frames, int64_t n_frames, int64_t row_bytes, const int64_t* start_idx,
                      int64_t n_trials, int64_t frames_per_trial, uint8_t* out, void* stream);

/* Upload of the frames a model actually reads.  src/train_rrr.py:48-49,171 keeps T = 100 of the 120 frames of every trial
 * (`X[:, sorted_idx]`, selected AFTER the per-frame z-score, so unselected frames never influence the fit): this copies
 * host_frames[k, idx_host[t], :] -> dev_out[k, t, :] for all K trials, one strided DMA (cudaMemcpy2DAsync) per run of
 * consecutive indices, on `stream`.  host_frames (K, Tf, row_bytes) uint8 in HOST memory (pinned for an asynchronous copy),
 * idx_host (T) strictly increasing int32 in HOST memory, dev_out (K, T, row_bytes) on the device.  Byte copy, bit-exact.   */
int vs_h2d_select_frames(const uint8_t* host_frames, int64_t K, int64_t Tf, int64_t row_bytes, const int32_t* idx_host,
                         int64_t T, uint8_t* dev_out, void* stream);

/* ------------------------------------------------------------------ Linear layers (M1-M3, G1)
 * torch.nn.Linear / ReLU as used by src/model/linear.py:24-32,45-53.
 *   y[b,o] = act( sum_i x[b,i] * W[o,i] + bias[o] ),  W is (out,in) row-major like nn.Linear.
 * x may be given as fp32 (x_f32) or, for the first layer, as raw uint8 frames (x_u8; then
 * x_f32 may be NULL).  workspace: vs_linear_fwd_workspace() bytes (split-K partials).   */
size_t vs_linear_fwd_workspace(int64_t batch, int64_t in_dim, int64_t out_dim);
int vs_linear_fwd(const float* x_f32, const uint8_t* x_u8, const float* W, const float* bias,
                  float* y, int64_t batch, int64_t in_dim, int64_t out_dim, int relu,
                  int engine, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the same layer (autograd of src/trainer/base.py:150).  dy is the gradient
 * w.r.t. the layer OUTPUT; if relu != 0, y (the forward output) masks it first.
 * Any of dx / dW / dbias may be NULL to skip.  dy_masked (batch,out) receives the masked
 * gradient when relu != 0 (may alias dy).  workspace: vs_linear_bwd_workspace() bytes
 * (row-range partials of dx for batch <= 32); NULL selects the generic kernels.          */
size_t vs_linear_bwd_workspace(int64_t batch, int64_t in_dim, int64_t out_dim);
int vs_linear_bwd(const float* dy, const float* y, const float* x_f32, const uint8_t* x_u8,
                  const float* W, float* dy_masked, float* dx, float* dW, float* dbias,
                  int64_t batch, int64_t in_dim, int64_t out_dim, int relu, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ Poisson NLL (C1)
 * torch.nn.PoissonNLLLoss(reduction="none", log_input=True) + .mean()
 * (src/train.py:59, src/trainer/base.py:141-143): loss = mean(exp(x) - t*x),
 * dlogits = (exp(x) - t)/n.  loss_sum (1 double, zeroed by the call) receives the SUM;
 * dlogits may be NULL (eval).                                                            */
int vs_poisson_nll(const float* logits, const float* target, double* loss_sum, float* dlogits,
                   int64_t n, void* stream);

/* ------------------------------------------------------------------ evaluation metrics (E2)
 * src/utils/metric_utils.py:36-102 + the loops of src/utils/utils.py:122-181, on the device.  rates / spikes / gt / pred
 * are (K, T, N) fp32 (what src/trainer/base.py:180-189 concatenates; rates = exp(logits)).
 * vs_bits_per_spike: bps_n[n] for every neuron (zero rates -> 1e-9, NaN spike bins masked, like the reference).
 * vs_r2_rows: r2_kt[k*T + t] = sklearn r2_score over the N neurons of (trial k, bin t); averaging over t gives the
 * `r2_score(gt[:, :, k], pred[:, :, k])` of utils.py:158.                                                        */
int vs_bits_per_spike(const float* rates, const float* spikes, int64_t K, int64_t T, int64_t N, double* bps_n, void* stream);
int vs_r2_rows(const float* gt, const float* pred, int64_t K, int64_t T, int64_t N, double* r2_kt, void* stream);

/* ------------------------------------------------------------------ AdamW (O1)
 * torch.optim.AdamW single-tensor update (src/train.py:44-49, src/trainer/base.py:151):
 *   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * lr and beta1 are per-step arguments because OneCycleLR cycles both.       */
typedef struct {
  double lr, beta1, beta2, eps, weight_decay; /* doubles: torch derives 1-beta, lr*wd, ... from Python floats */
  int64_t step; /* 1-based step count AFTER this update (torch's state['step']) */
} vs_adamw_hyper;

int vs_adamw(float* p, const float* g, float* m, float* v, int64_t n, vs_adamw_hyper h, void* stream);

/* G1 + O1 fused for the first layer: dW[o,i] = sum_b dy[b,o] x[b,i] is formed in registers
 * and consumed by the AdamW update without ever being written to HBM.  x as in
 * vs_linear_fwd.  batch <= 32.  bias (out) with its moments mb, vb is updated in the same
 * launch from dbias[o] = sum_b dy[b,o]; pass NULL to leave the bias alone.                */
int vs_dw_adamw_fused(const float* dy, const float* x_f32, const uint8_t* x_u8, float* W,
                      float* m, float* v, float* bias, float* mb, float* vb, int64_t batch,
                      int64_t in_dim, int64_t out_dim, vs_adamw_hyper h, void* stream);

/* ------------------------------------------------------------------ whole train step
 * The step body of src/trainer/base.py:147-154 for the `Linear` model
 * (src/model/linear.py:10-15): cast+flatten, 6 Linear layers with ReLUs, Poisson NLL,
 * backward, AdamW on every parameter -- enqueued back-to-back on `stream`.
 * Layers are listed first to last; relu[l] != 0 applies ReLU after layer l.
 * act[l]  : (batch, dims[l+1]) fp32 scratch for the layer outputs (act[L-1] = logits)
 * gact[l] : (batch, dims[l+1]) fp32 scratch for gradients
 * Layer 0 takes uint8 frames directly and uses the fused dW+AdamW kernel.               */
#define VS_MAX_LAYERS 16
typedef struct {
  int32_t n_layers;
  int64_t dims[VS_MAX_LAYERS + 1]; /* dims[0] = D, dims[L] = 100*N */
  int32_t relu[VS_MAX_LAYERS];
  float* W[VS_MAX_LAYERS];
  float* b[VS_MAX_LAYERS];
  float* mW[VS_MAX_LAYERS];
  float* vW[VS_MAX_LAYERS];
  float* mb[VS_MAX_LAYERS];
  float* vb[VS_MAX_LAYERS];
  float* act[VS_MAX_LAYERS];
  float* gact[VS_MAX_LAYERS];
  float* gW[VS_MAX_LAYERS]; /* gradient scratch, only read when batch > 32 (materialised-gradient route) */
  float* gb[VS_MAX_LAYERS];
} vs_mlp;

size_t vs_mlp_workspace(const vs_mlp* net, int64_t batch);
/* training step; loss_sum as in vs_poisson_nll (divide by batch*dims[L] for the mean) */
int vs_mlp_train_step(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32,
                      const float* target, int64_t batch, vs_adamw_hyper h, double* loss_sum,
                      int engine, void* workspace, size_t workspace_bytes, void* stream);
/* The same step with the FIRST layer row-parallel over ranks: this rank holds the pixel slice dims[0] of
 * every frame and the matching columns of W0 with their Adam moments; all later layers are replicated.
 *   phase 0: act[0] <- frames_slice . W0_slice^T (no bias, no ReLU).  The caller then all-reduces (sum) act[0], a
 *            (batch, dims[1]) fp32 matrix -- 16 KB at the reference sizes -- over NCCL.
 *   phase 1: act[0] <- act(act[0] + b0), layers 1.., loss, backward, AdamW on every local parameter (dW0 needs no
 *            collective: dH1 is replicated and the pixel slice is local).                                          */
int vs_mlp_train_step_rowpar(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32, const float* target,
                             int64_t batch, vs_adamw_hyper h, double* loss_sum, int engine, void* workspace,
                             size_t workspace_bytes, void* stream, int phase);
/* forward only (E1: src/trainer/base.py:161-206); logits land in net->act[L-1]; if target
 * is non-NULL the Poisson loss sum is also produced.                                    */
int vs_mlp_forward(const vs_mlp* net, const uint8_t* frames_u8, const float* x_f32,
                   const float* target, int64_t batch, double* loss_sum, int engine,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ RRR (R0-R6)
 * Reduced-rank regression of src/model/rrr.py.  Per session: X (K,T,C) with C = C1+1
 * (last column = the ones/bias column of src/train_rrr.py:155-162), y (K,T,N),
 * U (N,C1,r), V (r,T), b (N,1,T), all fp64 in the reference.
 *
 * Device layout built once per split by vs_rrr_pack (DESIGN.md "RRR data layout"); operand
 * rows are TIME-MAJOR, d = t*K + k, so the trials of one time bin are adjacent:
 *   Xa : planes x (K*T) x ldc   bf16, row d, c contiguous           (forward  A operand)
 *   Xb : planes x C1 x ldr      bf16, row c, column t*Kp + k        (backward A operand; every time bin is padded with
 *                               zeros to Kp = K rounded up to 16 trials, so a bin starts 32-byte aligned and ends on a
 *                               tensor-core K-step: the backward contracts one bin at a time)
 *   xl : (K*T) fp32, the last column of X, indexed by d
 * planes = 1 stores the 16-bit rounding of X; planes = 2 or 3 store the exact residual expansion
 * X ~= X0 + X1 (+ X2), each 16-bit, giving ~16 / ~24 significant bits with bf16 planes.  d.fmt selects bf16 or
 * IEEE half for all planes; with half, *overflow_flag (device int32, may be NULL; OR-ed, never cleared) becomes
 * non-zero when a value exceeds the half range (e.g. the exploding test features of SURVEY A18).          */
enum {
  VS_OPERAND_BF16 = 0, /* 8 significant bits, fp32 range */
  VS_OPERAND_F16 = 1   /* 11 significant bits at the same tensor-core rate; |x| <= 65504 (the pack calls report overflow) */
};
enum {
  VS_RRR_MODE_CLASSIC = 0, /* Xa and Xb both hold the z-scored matrix in `planes` residual planes */
  VS_RRR_MODE_EXACT = 1,   /* uint8 frames only: Xa = z-score as hi + lo half planes (factorised forward), the backward operand
                              holds the EXACT integers frame - round(mean); see vs_rrr_pack_u8_exact / vs_rrr_closure_exact */
  VS_RRR_MODE_DENSE = 2    /* uint8 frames only: BOTH contractions use the exact integers, one time bin at a time, like the
                              reference's einsum (no factorised product Z = X U): forward coefficient tiles generated on chip */
};
typedef struct {
  int64_t K, T, C1, N, r; /* trials, time bins, features without the bias column, neurons, rank */
  int32_t planes;         /* 1, 2 or 3 */
  int64_t ldc;            /* row pitch of Xa in elements, multiple of 64, >= C1 */
  int64_t ldr;            /* row pitch of Xb in elements, multiple of 64, >= T * roundup(K, 16) */
  int32_t fmt;            /* VS_OPERAND_*: 16-bit format of every tensor-core operand plane (X, U, residuals) */
  int32_t mode;           /* VS_RRR_MODE_* */
} vs_rrr_dims;

/* pitches the library wants for given sizes */
int64_t vs_rrr_ldc(int64_t C1);
int64_t vs_rrr_ldr(int64_t K, int64_t T);

/* X_trials: trials [k0, k0+nk) of the (K, T, C1+1) fp64 array the reference hands to RRRGD
 * (src/model/rrr.py:37-39), on the device.  Writes the matching rows of Xa / columns of Xb / xl;
 * call once with (0, K) or chunk by chunk to bound the fp64 staging buffer.               */
int vs_rrr_pack(const double* X_trials, int64_t k0, int64_t nk, vs_rrr_dims d, uint16_t* Xa, uint16_t* Xb,
                float* xl, int32_t* overflow_flag, void* stream);

/* R0 on device from raw frames (src/train_rrr.py:143-165 for the video modalities):
 * frames (K, Tf, F) uint8; sorted_idx (T) int32 frame indices; mean/std (Tf*F) fp64 as
 * produced by vs_rrr_colstats on the TRAIN split.  Writes the same Xa/Xb/xl as vs_rrr_pack
 * would for X = ((frames - mean)/std)[:, sorted_idx] with a ones column appended.        */
int vs_rrr_colstats(const uint8_t* frames, int64_t K, int64_t cols, double* mean, double* std_clipped, void* stream);
int vs_rrr_pack_u8(const uint8_t* frames, int64_t Tf, const int32_t* sorted_idx, const double* mean,
                   const double* std_clipped, vs_rrr_dims d, uint16_t* Xa, uint16_t* Xb, float* xl,
                   int32_t* overflow_flag, void* stream);
/* y (K,T,N) = (gaussian_filter1d(counts, sigma, axis=1, mode='reflect') - mean)/std as
 * src/train_rrr.py:118,145,165 (scipy.ndimage truncate=4).  counts (K,T,N) fp32; mean/std (T*N) fp64
 * may be NULL to get the smoothed counts only (used to derive the train statistics).      */
int vs_rrr_smooth_y(const float* counts, int64_t K, int64_t T, int64_t N, double sigma, const double* mean,
                    const double* std_clipped, float* y_out, void* stream);
/* same, with the float32 rounding residual as a second output (may be NULL): y_out + y_lo_out is the float64 value */
int vs_rrr_smooth_y2(const float* counts, int64_t K, int64_t T, int64_t N, double sigma, const double* mean,
                     const double* std_clipped, float* y_out, float* y_lo_out, void* stream);
/* mean / clipped population std over the K trials of a (K, cols) fp32 matrix (src/utils/utils.py:107-112) */
int vs_colstats_f32(const float* x, int64_t K, int64_t cols, double* mean, double* std_clipped, void* stream);

size_t vs_rrr_workspace(vs_rrr_dims d);

/* One closure evaluation of src/model/rrr.py:165-175 for one session:
 *   loss = sum (yhat - y)^2 + l2 * sum beta^2,   beta = cat(U@V, b)
 * and its gradient.  y is (K,T,N) fp32 (z-scored, smoothed targets).  Outputs fp64:
 * loss (1), sse_n (N) = per-neuron sum of squared residuals (rrr.py:151), dU (N,C1,r),
 * dV (r,T) [ACCUMULATED into: caller zeroes it once per closure so that several sessions
 * sharing V add up, rrr.py:46-49], db (N,1,T).  Any gradient pointer may be NULL
 * (evaluation only: rrr.py:179-181).                                                    */
int vs_rrr_closure(vs_rrr_dims d, const uint16_t* Xa, const uint16_t* Xb, const float* xl,
                   const float* y, const double* U, const double* V, const double* b, double l2,
                   double* loss, double* sse_n, double* dU, double* dV, double* db, int engine,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- exact-operand modes (VS_RRR_MODE_EXACT / VS_RRR_MODE_DENSE; d.planes = 2, d.fmt = VS_OPERAND_F16): the modes whose WHOLE
 * FIT lands within 1e-3 of the float64 reference (the reference trains with one un-line-searched LBFGS.step,
 * src/model/rrr.py:177,199, which amplifies operand rounding by 3-4 orders of magnitude: DESIGN.md "RRR precision").  For
 * uint8 frames (src/train_rrr.py:143-165) frame - round(mean) is an INTEGER of magnitude <= 255, exact in IEEE half; the
 * z-score scale 1/std[t,c] and the fractional part of the mean are applied around the tensor-core products, and the small
 * operand of each product (coefficients, residuals) is split into hi + lo half planes.
 *   EXACT: forward = factorised GEMM on Xa = z as hi + lo planes (3 plane products); backward on the integers Xi.
 *   DENSE: forward AND backward on the integers, per time bin (the reference's own contraction: a third of the flops per
 *          plane product); Xa is unused (NULL); dV comes from a second pass of the backward kernel over the same tiles.
 * Operand table of one split (device pointers; `const` because the closure only reads them, vs_rrr_pack_u8_exact writes):  */
typedef struct {
  const uint16_t* Xi;   /* (C1, ldr) half: exact integers, layout of Xb (backward A operand); NULL for evaluation-only splits */
  const float* isdT;    /* (C1, ldt) fp32: 1/std[t,c]                      (weights of the dense backward), ldt = vs_rrr_ldt(T) */
  const float* qT;      /* (C1, ldt) fp32: (mean - round(mean))[t,c]/std   (rank-T correction of the backward's result)       */
  int64_t ldt;
  const float* y_lo;    /* (K,T,N) fp32 or NULL: y + y_lo is the target at float64 precision */
  /* VS_RRR_MODE_DENSE only */
  const uint16_t* Xc;   /* (K*T, ldc) half: exact integers, layout of Xa, row t*K + k (forward A operand) */
  const float* isd;     /* (T, ldc) fp32: 1/std[t,c], 0 for features that are constant in the train split */
  const uint16_t* qh;   /* (2, roundup(T,16), ldc) half: hi + lo planes of q[t,c] = (mean - round(mean))/std */
  const float* isdmax;  /* (T) fp32: max_c isd[t,c] */
} vs_rrr_exact_ops;
int64_t vs_rrr_ldt(int64_t T);
/* 1 if the exact-operand closures cover the shape (rank 3, N <= 160 after padding to 16, more than 128 features) */
int vs_rrr_exact_supported(int64_t K, int64_t T, int64_t C1, int64_t N, int64_t r);
/* fills Xa (EXACT) / out->Xc (DENSE), out->Xi and whichever tables of `out` are non-NULL */
int vs_rrr_pack_u8_exact(const uint8_t* frames, int64_t Tf, const int32_t* sorted_idx, const double* mean,
                         const double* std_clipped, vs_rrr_dims d, uint16_t* Xa, const vs_rrr_exact_ops* out, float* xl,
                         int32_t* overflow_flag, void* stream);
/* The same operands from ONE read of the frames (statistics and pack fused per time bin: the K trials of a 128-feature
 * chunk of one frame are staged in shared memory).  compute_stats != 0 (train split): mean / std_clipped of the selected
 * frames sorted_idx[t] are computed (src/utils/utils.py:107-112) and written into the (Tf, C1) tables; == 0: read.
 * Needs C1 % 4 == 0 and K <= 1600; otherwise VS_ERR_UNSUPPORTED (use vs_rrr_colstats + vs_rrr_pack_u8_exact).          */
int vs_rrr_pack_u8_fused(const uint8_t* frames, int64_t Tf, const int32_t* sorted_idx, double* mean, double* std_clipped,
                         int compute_stats, vs_rrr_dims d, uint16_t* Xa, const vs_rrr_exact_ops* out, float* xl,
                         int32_t* overflow_flag, void* stream);
/* vs_rrr_closure for an exact-operand split; all epilogue sums are float64 */
int vs_rrr_closure_exact(vs_rrr_dims d, const uint16_t* Xa, const vs_rrr_exact_ops* ops, const float* xl, const float* y,
                         const double* U, const double* V, const double* b, double l2, double* loss, double* sse_n,
                         double* dU, double* dV, double* db, void* workspace, size_t workspace_bytes, void* stream);
/* vs_rrr_predict for a VS_RRR_MODE_DENSE split (EXACT splits use vs_rrr_predict with their Xa) */
int vs_rrr_predict_exact(vs_rrr_dims d, const vs_rrr_exact_ops* ops, const float* xl, const double* U, const double* V,
                         const double* b, double* yhat, void* workspace, size_t workspace_bytes, void* stream);

/* src/model/rrr.py:105-130 (predict_y): yhat (K,T,N) fp64 for one split. */
int vs_rrr_predict(vs_rrr_dims d, const uint16_t* Xa, const float* xl, const double* U, const double* V,
                   const double* b, double* yhat, int engine, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ------------------------------------------------------------------ L-BFGS vector passes (R5)
 * torch.optim.LBFGS(...).step(closure) without line search, as src/model/rrr.py:177,199 uses it, on flat
 * fp64 vectors of n elements.  The two-loop recursion itself runs on the host in coefficient space
 * (optim.py, FusedLBFGS); the device does two streaming passes per iteration.  History vectors live in one
 * buffer `hist` of slots of `hist_stride` elements; s_slots_host / y_slots_host (HOST arrays, m entries,
 * oldest pair first) name the slots holding s_i and y_i.
 *
 * vs_lbfgs_dots: y_out = g - g_prev (skipped if y_out is NULL; g_prev / s_new may be NULL on the very first
 * evaluation) and out[0..8+6m) =
 *   [0] g.g  [1] sum|g|  [2] max|g|  [3] y.y  [4] y.s_new  [5] s_new.g  [6] y.g  [7] 0
 *   [8+3i+{0,1,2}]       s_i.g, s_i.y, s_i.s_new          i < m
 *   [8+3(m+i)+{0,1,2}]   y_i.g, y_i.y, y_i.s_new          i < m
 * vs_lbfgs_direction: d = coef[0]*g + sum_i coef[1+i]*s_i + coef[1+m+i]*y_i (coef_host: 2m+1 HOST doubles);
 * s_out = t*d; x += t*d (skipped when x is NULL); *dmax_out = max|t*d|.                               */
#define VS_LBFGS_MAX_HIST 100
size_t vs_lbfgs_workspace(int64_t n, int m);
/* hist_f32 != 0: the history buffer (and therefore s_new / y_out / s_out, which are history slots) holds float32
 * instead of float64 -- half the traffic of both passes; arithmetic stays fp64 and every inner product is taken with
 * the stored (rounded) vectors.  hist_stride in elements; a multiple of 4 enables the 128-bit path.            */
int vs_lbfgs_dots(int64_t n, const double* g, const double* g_prev, const void* s_new, void* y_out,
                  const void* hist, int64_t hist_stride, int hist_f32, const int32_t* s_slots_host,
                  const int32_t* y_slots_host, int m, double* out, void* workspace, size_t workspace_bytes,
                  void* stream);
int vs_lbfgs_direction(int64_t n, const double* g, const void* hist, int64_t hist_stride, int hist_f32,
                       const int32_t* s_slots_host, const int32_t* y_slots_host, int m,
                       const double* coef_host, double t, double* x, void* s_out, double* dmax_out,
                       void* stream);

/* ---- device-driven variant: the optimiser state lives on the GPU and a one-thread kernel takes every decision of
 * torch.optim.LBFGS.step (memory update, two-loop recursion, step length, termination), so a whole step(closure) is
 * enqueued without a single host<->device synchronisation: per iteration  closure kernels -> vs_lbfgs_dev_dots ->
 * vs_lbfgs_dev_update -> vs_lbfgs_dev_direction.  Once the state says `done` the remaining launches are no-ops.
 * The caller owns the state (a device buffer of sizeof(vs_lbfgs_dev), initialised on the host with
 * vs_lbfgs_dev_init_host and copied over) and reads it back once at the end.                                   */
typedef struct {
  int32_t m;            /* curvature pairs in memory */
  int32_t n_iter;       /* iterations of the current step() */
  int32_t total_iter;   /* torch's state["n_iter"] (all step() calls) */
  int32_t func_evals;   /* torch's state["func_evals"] */
  int32_t cur_evals;    /* closure evaluations of the current step() */
  int32_t done;         /* 0 running | 1 converged at the first evaluation | 2 directional derivative | 3 max_eval |
                           4 gradient tolerance | 5 step tolerance | 6 loss tolerance */
  int32_t have_prev, have_s;
  int32_t s_cur, y_next;      /* slot of the latest step s = t*d; slot the next dots pass writes y into */
  int32_t n_free, pad0;
  int32_t free_slots[2 * VS_LBFGS_MAX_HIST + 8];
  int32_t s_slots[VS_LBFGS_MAX_HIST], y_slots[VS_LBFGS_MAX_HIST];
  double H_diag, prev_loss, loss, t, gtd, dmax;
  double coef[2 * VS_LBFGS_MAX_HIST + 2];
  double out[8 + 6 * VS_LBFGS_MAX_HIST];
  double SY[VS_LBFGS_MAX_HIST * VS_LBFGS_MAX_HIST], YY[VS_LBFGS_MAX_HIST * VS_LBFGS_MAX_HIST];
} vs_lbfgs_dev;

/* fills *state_host for an empty memory over `n_slots` history slots (>= 2*min(history_size, iterations) + 4) */
int vs_lbfgs_dev_init_host(vs_lbfgs_dev* state_host, int n_slots);
size_t vs_lbfgs_dev_state_bytes(void);   /* sizeof(vs_lbfgs_dev), for bindings that mirror the struct */
size_t vs_lbfgs_dev_workspace(int64_t n);
/* one pass over g (and g_prev, the latest step, the history): y, all inner products -> state->out */
int vs_lbfgs_dev_dots(vs_lbfgs_dev* state, int64_t n, const double* g, const double* g_prev, void* hist,
                      int64_t hist_stride, int hist_f32, void* workspace, size_t workspace_bytes, void* stream);
/* decisions of one iteration.  loss: device scalar of the latest closure evaluation; first_eval != 0 for the evaluation
 * that opens a step() call (resets the per-step counters and applies torch's "already converged" early return)      */
int vs_lbfgs_dev_update(vs_lbfgs_dev* state, const double* loss, double lr, double tolerance_grad, double tolerance_change,
                        int max_eval, int history_size, int first_eval, void* stream);
/* d from the state's coefficients; hist[s_cur] = t*d; x += t*d */
int vs_lbfgs_dev_direction(vs_lbfgs_dev* state, int64_t n, const double* g, void* hist, int64_t hist_stride, int hist_f32,
                           double* x, void* stream);

/* ---- compact history (device-driven, float64): ONE stored vector per closure evaluation -- the basis b_0 = g_0,
 * b_l = g_l - g_(l-1) -- instead of the (s_i, y_i) pair; the steps s_i live as coefficient rows over the basis and every
 * inner product of the two-loop recursion is a contraction of the basis Gram matrix P.  Same decisions, same order as
 * vs_lbfgs_dev_update (= torch.optim.LBFGS.step); both passes stream nb vectors instead of 2m and no step vector is
 * written: about half the HBM traffic.  At most VS_LBFGS_CMAX closure evaluations over the life of the state (done = 7
 * when the basis is full); history slot l of `hist` holds b_l.                                                       */
#define VS_LBFGS_CMAX 64
typedef struct {
  int32_t m, nb;         /* curvature pairs in memory; basis vectors stored (= closure evaluations seen) */
  int32_t n_iter, total_iter, func_evals, cur_evals, done, have_prev, have_s, pad0;
  int32_t iy[VS_LBFGS_CMAX];                 /* pair i's y is basis vector iy[i] */
  double H_diag, prev_loss, loss, t, gtd, dmax;
  double coef[VS_LBFGS_CMAX + 2];            /* [0] = cg (weight of g), [1+l] = weight of b_l in the direction */
  double Acur[VS_LBFGS_CMAX];                /* coefficients of the latest step s = t*d */
  double out[8 + 3 * VS_LBFGS_CMAX];         /* scalars of the dots pass */
  double P[VS_LBFGS_CMAX * VS_LBFGS_CMAX];   /* basis Gram matrix */
  double A[VS_LBFGS_CMAX * VS_LBFGS_CMAX];   /* row i: coefficients of s_i */
  double SY[VS_LBFGS_CMAX * VS_LBFGS_CMAX], YY[VS_LBFGS_CMAX * VS_LBFGS_CMAX];
} vs_lbfgs_cdev;
int vs_lbfgs_cdev_init_host(vs_lbfgs_cdev* state_host);
size_t vs_lbfgs_cdev_state_bytes(void);
size_t vs_lbfgs_cdev_workspace(int64_t n);
int vs_lbfgs_cdev_dots(vs_lbfgs_cdev* state, int64_t n, const double* g, const double* g_prev, double* hist,
                       int64_t hist_stride, void* workspace, size_t workspace_bytes, void* stream);
int vs_lbfgs_cdev_update(vs_lbfgs_cdev* state, const double* loss, double lr, double tolerance_grad, double tolerance_change,
                         int max_eval, int first_eval, void* stream);
int vs_lbfgs_cdev_direction(vs_lbfgs_cdev* state, int64_t n, const double* g, double* hist, int64_t hist_stride, double* x,
                            void* stream);

/* HOST: the two-loop recursion of torch.optim.LBFGS in coefficient space, between the two device passes.  Inputs are
 * the inner products vs_lbfgs_dots gathered (SY[i*ld+j] = s_i.y_j, YY[i*ld+j] = y_i.y_j, sg[i] = s_i.g,
 * yg[i] = y_i.g, gg = g.g) and torch's H_diag; coef_out (2m+1) receives [cg, cs_0.., cy_0..] for
 * vs_lbfgs_direction and *gtd_out the directional derivative g.d.                                            */
int vs_host_lbfgs_two_loop(int m, double gg, const double* sg, const double* yg, const double* SY, const double* YY,
                           int ld, double H_diag, double* coef_out, double* gtd_out);

/* ------------------------------------------------------------------ RRR initialisation stream (R1, HOST)
 * src/model/rrr.py:35,42-43 draws U and V from numpy's global legacy RandomState after np.random.seed(0).
 * These HOST functions (the only entry points that take host pointers and launch nothing) reproduce that
 * stream bit for bit -- MT19937 seeded like RandomState.seed(uint32), 53-bit doubles from two words, the polar
 * legacy_gauss with its cached second value, libm log/sqrt -- with the transform spread over `threads` host
 * threads (0 = all cores); only the MT19937 recurrence is sequential.
 * state: caller-allocated blob of VS_HOST_RNG_STATE_BYTES.  vs_host_rng_normal writes the next n normals,
 * each divided by `divisor` (rrr.py divides by sqrt(T*ncomp); pass 1.0 for plain draws).
 * get/set_state exchange numpy's ('MT19937', key[624], pos, has_gauss, cached_gaussian) tuple so the global
 * numpy stream can be left exactly where the reference would leave it.                              */
#define VS_HOST_RNG_STATE_BYTES 2560
int vs_host_rng_seed(void* state, uint32_t seed);
int vs_host_rng_normal(void* state, int64_t n, double divisor, double* out_host, int threads);
int vs_host_rng_get_state(const void* state, uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* cached);
int vs_host_rng_set_state(void* state, const uint32_t* key624, int32_t pos, int32_t has_gauss, double cached);

/* ------------------------------------------------------------------ plain TN GEMM (test hook)
 * C[M,N] (fp32, row-major, ldc) = A[M,K] * B[N,K]^T with bf16 or tf32(fp32) operands, both
 * K-major with pitches lda/ldb (elements).  Exposed so tests can exercise the tcgen05 core
 * against the SIMT engine on arbitrary shapes.  dtype: 0 = bf16, 1 = tf32, 2 = f16.     */
int vs_gemm_tn(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
               int64_t ldb, int64_t ldc, int dtype, int engine, void* stream);

/* ------------------------------------------------------------------ instrumentation
 * Number of kernel launches this library has enqueued from the calling process since the
 * last reset (bench.py's gpu_launches).                                                 */
int64_t vs_launch_count(void);
void vs_launch_count_reset(void);
/* Per-kernel device timing for the roofline report: when enabled, the dominant kernels (tag 0 = the
 * tcgen05 GEMM, tag 1 = fused dW+AdamW, tag 2 = the dense RRR backward, tag 3 = the dense RRR forward, tag 4 = its dV pass) are bracketed by CUDA events on the launching stream.
 * vs_profile_read sums the recorded intervals of one tag and returns how many there were. */
void vs_profile_enable(int on);
int64_t vs_profile_read(int tag, double* total_ms, double* min_ms, double* max_ms);

#ifdef __cplusplus
}
#endif
#endif /* VS_B200_H */
