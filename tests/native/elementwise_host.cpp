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

// ---- small_gemm.cuh: the thread -> (output, k-slice) map of er_gemm_small replayed serially ----------------------------
#include <vector>

#include "small_gemm.cuh"

extern "C" long host_gemm_small(const float* A, long sa_m, long sa_k, const float* B, long sb_k, long sb_n,
                                const float* bias, float* C, long ldc, long M, long N, long K) {
  er::SmallGemm g;
  g.a = A; g.b = B; g.sa_i = sa_m; g.sa_k = sa_k; g.sb_k = sb_k; g.sb_j = sb_n; g.M = M; g.N = N; g.K = K;
  const long n_slice = er::small_gemm_slices(M, N, K);
  g.k_per_slice = (K + n_slice - 1) / n_slice;
  const long n_out = M * N;
  if (n_slice == 1) {
    for (long o = 0; o < n_out; ++o) {
      const float v = er::small_gemm_cell(g, o, 0);
      C[(o / N) * ldc + o % N] = bias ? v + bias[o % N] : v;
    }
    return 1;
  }
  std::vector<float> part((size_t)(n_slice * n_out));
  for (long s = 0; s < n_slice; ++s)
    for (long o = 0; o < n_out; ++o) part[(size_t)(s * n_out + o)] = er::small_gemm_cell(g, o, s);
  for (long o = 0; o < n_out; ++o) C[(o / N) * ldc + o % N] = er::small_gemm_reduce(part.data(), n_out, n_slice, o, bias, N);
  return n_slice;
}

// dice gate: value and the three gradient terms, out = [4][n]
extern "C" void host_dice(const float* x, const float* xn, const float* alpha, const float* gy, long n, float* out) {
  for (long i = 0; i < n; ++i) {
    out[i] = er::dice_value(x[i], xn[i], alpha[i]);
    er::dice_grads(x[i], xn[i], alpha[i], gy[i], out + n + i, out + 2 * n + i, out + 3 * n + i);
  }
}
