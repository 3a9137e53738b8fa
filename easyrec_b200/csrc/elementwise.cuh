// Elementwise formulas shared by the kernels and by the host-side check tests/native/elementwise_host.cpp: the SAME
// source is compiled for the device (dense.cu, metrics.cu) and, with a plain C++ compiler, for the CPU tests - so the
// formulas are checked on the CPU as well (the launches themselves are GPU tests).
#pragma once
#include <math.h>
#include <stdint.h>

#include "er_b200.h"

#ifdef __CUDACC__
#define ER_HD __host__ __device__ __forceinline__
#else
#define ER_HD inline
#endif

namespace er {

// ---- the stateless non-relu activations of get_activation (utils/activation.py:66-118) ---------------------------
// value and derivative from the pre-activation x, in the branch conventions of TF's gradient kernels (EluGrad / SeluGrad
// take the negative branch for x < 0, LeakyReluGrad for x <= 0).
template <int KIND>
ER_HD float act_value(float x) {
  if (KIND == ER_ACT_GELU) {   // x * 0.5 * (1 + tanh(sqrt(2/pi) * (x + 0.044715 x^3)))  (activation.py:46-60)
    const float u = 0.7978845608028654f * (x + 0.044715f * (x * x * x));
    return x * (0.5f * (1.0f + tanhf(u)));
  }
  if (KIND == ER_ACT_LEAKY_RELU) return fmaxf(0.2f * x, x);              // tf.nn.leaky_relu, alpha 0.2
  if (KIND == ER_ACT_ELU) return x < 0.f ? expm1f(x) : x;
  if (KIND == ER_ACT_SELU) return x < 0.f ? 1.7580993408473768f * expm1f(x) : 1.0507009873554805f * x;
  if (KIND == ER_ACT_TANH) return tanhf(x);
  if (KIND == ER_ACT_SWISH) return x / (1.0f + expf(-x));                 // x * sigmoid(x)
  return 1.0f / (1.0f + expf(-x));                                        // ER_ACT_SIGMOID
}

template <int KIND>
ER_HD float act_slope(float x) {
  if (KIND == ER_ACT_GELU) {
    const float c = 0.7978845608028654f;
    const float t = tanhf(c * (x + 0.044715f * (x * x * x)));
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x * x);
  }
  if (KIND == ER_ACT_LEAKY_RELU) return x > 0.f ? 1.0f : 0.2f;
  if (KIND == ER_ACT_ELU) return x < 0.f ? expf(x) : 1.0f;
  if (KIND == ER_ACT_SELU) return x < 0.f ? 1.7580993408473768f * expf(x) : 1.0507009873554805f;
  if (KIND == ER_ACT_TANH) {
    const float t = tanhf(x);
    return 1.0f - t * t;
  }
  const float s = 1.0f / (1.0f + expf(-x));
  if (KIND == ER_ACT_SWISH) return s * (1.0f + x * (1.0f - s));
  return s * (1.0f - s);
}


// ---- tf.metrics.auc (model/rank_model.py:360-373) -------------------------------------------------------------------
// bin of a prediction = number of thresholds strictly below it (math_ops.greater(pred, thr)); thr ascending; a NaN
// prediction exceeds none.
ER_HD int auc_bin(const float* thr, int n_thr, float p) {
  int lo = 0, hi = n_thr;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (thr[mid] < p)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
// tf.to_int64(label) then cast to bool (rank_model.py:361, metrics_impl.auc): positive = truncated label != 0
ER_HD bool auc_positive(float label) { return (long long)label != 0; }

}  // namespace er
