// Streaming evaluation metrics on the device: the confusion-matrix accumulators behind
// tf.metrics.auc (model/rank_model.py:360-373, eval.proto AUC.num_thresholds) and max_f1 (core/metrics.py:25-56).
//
// TF keeps tp / fn / tn / fp per threshold and adds `label & (pred > thr[i])` counts batch by batch.  All four follow
// from two histograms over k(pred) = #{i : thr[i] < pred}:  tp[i] = sum_{k > i} pos[k],  fp[i] = sum_{k > i} neg[k],
// fn = P - tp, tn = N - fp.  The kernel adds one batch into those histograms (uint64 counters in device memory: exact,
// order-independent, so an evaluate() pass never leaves the device until the final read of 2 * (T + 1) counters).
#include "common.cuh"
#include "elementwise.cuh"

namespace er {

// thresholds (ascending) and the CTA's private counters live in shared memory; integer atomics only.
__global__ void __launch_bounds__(256)
    auc_hist_kernel(const float* __restrict__ probs, const float* __restrict__ labels, int64_t n,
                    const float* __restrict__ thr, int n_thr, unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned char smem_raw[];
  float* s_thr = reinterpret_cast<float*>(smem_raw);
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(smem_raw + (size_t)n_thr * sizeof(float));
  const int n_bin = n_thr + 1;
  for (int i = threadIdx.x; i < n_thr; i += blockDim.x) s_thr[i] = thr[i];
  for (int i = threadIdx.x; i < 2 * n_bin; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = probs[i];
    const int lo = auc_bin(s_thr, n_thr, p);
    const bool pos = auc_positive(labels[i]);
    atomicAdd(&s_cnt[(pos ? n_bin : 0) + lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * n_bin; i += blockDim.x) {
    const unsigned int c = s_cnt[i];
    if (c) atomicAdd(&hist[i], (unsigned long long)c);
  }
}

}  // namespace er

extern "C" int er_auc_hist(const float* probs, const float* labels, int64_t n, const float* thresholds,
                           int32_t n_thresholds, uint64_t* hist, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(probs && labels && thresholds && hist, "null argument");
  ER_REQUIRE(n > 0, "n must be positive");
  ER_REQUIRE(n_thresholds >= 2 && n_thresholds <= 4095, "num_thresholds must be in [2, 4095]");
  const size_t smem = (size_t)n_thresholds * sizeof(float) + 2 * (size_t)(n_thresholds + 1) * sizeof(unsigned int);
  static bool attr = false;   // (up to 48 KB at 4095 thresholds: opt in explicitly, the default limit sits right there)
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(auc_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return fail(ER_ERR_CUDA, std::string("er_auc_hist: ") + cudaGetErrorString(e));
    attr = true;
  }
  const int grid = grid_for(n, 256 * 8, 2);
  auc_hist_kernel<<<grid, 256, smem, as_stream(stream)>>>(probs, labels, n, thresholds, (int)n_thresholds,
                                                          reinterpret_cast<unsigned long long*>(hist));
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}
