// Small-batch (B <= 32) layer kernels: internal interface (smallbatch.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vs_b200.h"

namespace vs {
namespace sb {
bool supported(long long batch, long long in_dim, long long out_dim);      // dx / dw
bool fwd_supported(long long batch, long long in_dim, long long out_dim);  // + x fits in shared memory
// y = act(x W^T + bias)
int fwd(const float* x, const float* W, const float* bias, float* y, long long batch, long long in_dim, long long out_dim,
        int relu, cudaStream_t st);
// dx = g W, optionally masked by (act_prev > 0) (ReLU of the layer below); workspace holds the row-range partials
size_t dx_workspace(long long batch, long long in_dim, long long out_dim);
int dx(const float* g, const float* W, const float* act_prev, float* dx_out, long long batch, long long in_dim, long long out_dim,
       void* workspace, size_t workspace_bytes, cudaStream_t st);
// dW = g^T x consumed by AdamW in registers (bias = sum_b g likewise; bias may be NULL)
int dw_adamw(const float* dy, const float* xf, const uint8_t* xu, float* W, float* M, float* V, float* bias, float* mb, float* vb,
             long long batch, long long in_dim, long long out_dim, const vs_adamw_hyper& h, cudaStream_t st);
// dW / dbias stored (dbias may be NULL)
int dw_store(const float* dy, const float* xf, const uint8_t* xu, float* dW, float* dbias, long long batch, long long in_dim,
             long long out_dim, cudaStream_t st);
}  // namespace sb
}  // namespace vs
