// Adam arithmetic shared by the optimiser kernels (optim.cu) and by tile kernels that replay pending
// zero-gradient steps while they gather (tile_small.cu).  torch.optim.Adam single-tensor math (torch 2.11
// optim/adam.py), per step t = 1, 2, ...:
//   m <- m + (1-b1)(g - m);  v <- b2 v + (1-b2) g^2
//   p <- p - (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// A row whose gradient is zero at step s still moves: m <- b1 m, v <- b2 v, same p update.
#pragma once
#include <cuda_runtime.h>

// Zero-gradient steps replayed exactly per row.  The per-step term decays like
// (b1/sqrt(b2))^j ~ 0.9005^j, so what is dropped beyond 160 steps is < 1e-7 of the first term.
constexpr int kMaxReplay = 160;

struct AdamConst {
  float lr, b1, b2, eps, ln_b1, ln_b2, sqrt_b2;
};

__device__ __forceinline__ float bias_c1(const AdamConst& c, float s) {  // lr / (1 - b1^s)
  return c.lr / (-expm1f(s * c.ln_b1));
}
__device__ __forceinline__ float bias_c2(const AdamConst& c, float s) {  // 1 / sqrt(1 - b2^s)
  return 1.f / sqrtf(-expm1f(s * c.ln_b2));
}

// Brings one element from "state after step `last`" to "state after step last+gap" under zero
// gradient.  c1s / c2s hold the bias terms of steps last+1 .. last+min(gap, kMaxReplay).
__device__ __forceinline__ void replay_zero_steps(float& p, float& m, float& v, int gap,
                                                  const float* c1s, const float* c2s,
                                                  const AdamConst& c) {
  if (gap <= 0) return;
  const int n = min(gap, kMaxReplay);
  const float m0 = m, v0 = v;
  float mm = m0, r = sqrtf(v0);
  if (m0 != 0.f) {
    // the terms are independent but for mm, r and the running p: unrolled, their table loads and reciprocals
    // overlap (a row 160 steps behind is the critical path of a small-batch step)
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      mm *= c.b1;
      r *= c.sqrt_b2;
      p -= c1s[j] * __fdividef(mm, fmaf(r, c2s[j], c.eps));
    }
  }
  m = m0 * expf((float)gap * c.ln_b1);
  v = v0 * expf((float)gap * c.ln_b2);
}

// The same replay with the bias terms evaluated on the fly (rows whose steps are not in a caller's table).
__device__ __forceinline__ void replay_inline(float& p, float& m, float& v, int last, int gap, const AdamConst& c) {
  const int n = min(gap, kMaxReplay);
  const float m0 = m, v0 = v;
  float mm = m0, r = sqrtf(v0);
  if (m0 != 0.f) {
    for (int j = 0; j < n; ++j) {
      const float s = (float)(last + 1 + j);
      mm *= c.b1;
      r *= c.sqrt_b2;
      p -= bias_c1(c, s) * __fdividef(mm, fmaf(r, bias_c2(c, s), c.eps));
    }
  }
  m = m0 * expf((float)gap * c.ln_b1);
  v = v0 * expf((float)gap * c.ln_b2);
}

__device__ __forceinline__ void adam_real_step(float& p, float& m, float& v, float g, float c1t,
                                               float c2t, const AdamConst& c) {
  m = fmaf(1.f - c.b1, g - m, m);
  v = fmaf(c.b2, v, (1.f - c.b2) * g * g);
  p -= c1t * (m / fmaf(sqrtf(v), c2t, c.eps));
}

inline AdamConst make_adam_const(float lr, float beta1, float beta2, float eps) {
  AdamConst c;
  c.lr = lr; c.b1 = beta1; c.b2 = beta2; c.eps = eps;
  c.ln_b1 = (float)log((double)beta1);
  c.ln_b2 = (float)log((double)beta2);
  c.sqrt_b2 = (float)sqrt((double)beta2);
  return c;
}
