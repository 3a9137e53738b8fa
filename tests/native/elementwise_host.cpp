// TEST INFRASTRUCTURE: compiles easyrec_b200/csrc/elementwise.cuh - the very source the kernels are built from - with
// a plain C++ compiler, so the CPU suite can check the activation formulas and the threshold binning against the
// oracle where no GPU is present.  Not linked into liber_b200.so; built into a temporary .so by the test.
#include "elementwise.cuh"

namespace {
template <int KIND>
void run(const float* x, long n, float* y, float* slope) {
  for (long i = 0; i < n; ++i) {
    y[i] = er::act_value<KIND>(x[i]);
    slope[i] = er::act_slope<KIND>(x[i]);
  }
}
}  // namespace

extern "C" int host_act(int kind, const float* x, long n, float* y, float* slope) {
  switch (kind) {
    case ER_ACT_GELU: run<ER_ACT_GELU>(x, n, y, slope); return 0;
    case ER_ACT_LEAKY_RELU: run<ER_ACT_LEAKY_RELU>(x, n, y, slope); return 0;
    case ER_ACT_ELU: run<ER_ACT_ELU>(x, n, y, slope); return 0;
    case ER_ACT_SELU: run<ER_ACT_SELU>(x, n, y, slope); return 0;
    case ER_ACT_TANH: run<ER_ACT_TANH>(x, n, y, slope); return 0;
    case ER_ACT_SWISH: run<ER_ACT_SWISH>(x, n, y, slope); return 0;
    case ER_ACT_SIGMOID: run<ER_ACT_SIGMOID>(x, n, y, slope); return 0;
  }
  return 1;
}

extern "C" void host_auc_hist(const float* probs, const float* labels, long n, const float* thr, int n_thr,
                              unsigned long long* hist) {
  for (long i = 0; i < n; ++i)
    hist[(er::auc_positive(labels[i]) ? n_thr + 1 : 0) + er::auc_bin(thr, n_thr, probs[i])] += 1;
}
