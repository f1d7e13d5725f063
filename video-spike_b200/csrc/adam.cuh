// AdamW arithmetic shared by the update kernels (torch.optim.AdamW single-tensor semantics,
// src/train.py:44-49): scalars derived in double on the host like the Python floats torch uses.
#pragma once
#include <math.h>
#include "common.cuh"

namespace vs {

struct AdamConsts {
  float decay;      // 1 - lr*wd
  float beta1, w1;  // w1 = 1 - beta1
  float beta2, w2;  // w2 = 1 - beta2
  float step_size;  // lr / (1 - beta1^t)
  float inv_bc2s;   // 1 / sqrt(1 - beta2^t)
  float eps;
};

static inline AdamConsts make_consts(const vs_adamw_hyper& h) {
  // scalars in double like the Python floats torch.optim.AdamW computes them with
  const double bc1 = 1.0 - pow(h.beta1, (double)h.step);
  const double bc2 = 1.0 - pow(h.beta2, (double)h.step);
  AdamConsts c;
  c.decay = (float)(1.0 - h.lr * h.weight_decay);
  c.beta1 = (float)h.beta1; c.w1 = (float)(1.0 - h.beta1);
  c.beta2 = (float)h.beta2; c.w2 = (float)(1.0 - h.beta2);
  c.step_size = (float)(h.lr / bc1);
  c.inv_bc2s = (float)(1.0 / sqrt(bc2));
  c.eps = (float)h.eps;
  return c;
}

__device__ __forceinline__ void adamw_elem(float& p, float& m, float& v, float g, const AdamConsts& c) {
  p *= c.decay;
  m = fmaf(g - m, c.w1, m);                    // lerp_(grad, 1-beta1)
  v = fmaf(v, c.beta2, c.w2 * g * g);          // mul_(beta2).addcmul_(g, g, 1-beta2)
  const float denom = sqrtf(v) * c.inv_bc2s + c.eps;
  p -= c.step_size * (m / denom);
}

}  // namespace vs
