// er_gemm_small: see small_gemm.cuh.  One thread per output element (and per k-slice when K is cut); the k-slice
// partials are summed in slice order by a second launch - no atomics.
#include <math.h>

#include "common.cuh"
#include "small_gemm.cuh"

namespace er {

__global__ void __launch_bounds__(256)
    small_gemm_kernel(SmallGemm g, const float* __restrict__ bias, float* __restrict__ c, int64_t ldc,
                      float* __restrict__ part, int64_t n_slice) {
  const int64_t n_out = g.M * g.N;
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int64_t s = blockIdx.y;
  const float v = small_gemm_cell(g, o, s);
  if (n_slice == 1) {
    const int64_t i = o / g.N, j = o - i * g.N;
    c[i * ldc + j] = bias ? v + bias[j] : v;
  } else {
    part[s * n_out + o] = v;
  }
}

__global__ void __launch_bounds__(256)
    small_gemm_reduce_kernel(const float* __restrict__ part, int64_t M, int64_t N, int64_t n_slice,
                             const float* __restrict__ bias, float* __restrict__ c, int64_t ldc) {
  const int64_t n_out = M * N;
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int64_t i = o / N, j = o - i * N;
  c[i * ldc + j] = small_gemm_reduce(part, n_out, n_slice, o, bias, N);
}

}  // namespace er

extern "C" size_t er_gemm_small_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int64_t s = er::small_gemm_slices(M, N, K);
  return s > 1 ? (size_t)(s * M * N) * sizeof(float) : 0;
}

extern "C" int er_gemm_small(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n,
                             const float* bias, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* ws,
                             size_t ws_bytes, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(A && B && C, "null argument");
  ER_REQUIRE(M > 0 && N > 0 && K > 0 && ldc >= N, "bad shape");
  ER_REQUIRE(M * N <= ((int64_t)1 << 31), "output too large for the vector-sized path");
  const int64_t n_slice = small_gemm_slices(M, N, K);
  if (n_slice > 1 && (!ws || ws_bytes < er_gemm_small_workspace_bytes(M, N, K)))
    return fail(ER_ERR_WORKSPACE, "er_gemm_small: workspace too small");
  SmallGemm g;
  g.a = A;
  g.b = B;
  g.sa_i = sa_m;
  g.sa_k = sa_k;
  g.sb_k = sb_k;
  g.sb_j = sb_n;
  g.M = M;
  g.N = N;
  g.K = K;
  g.k_per_slice = ceil_div(K, n_slice);
  cudaStream_t st = as_stream(stream);
  const dim3 grid((unsigned)ceil_div(M * N, 256), (unsigned)n_slice);
  small_gemm_kernel<<<grid, 256, 0, st>>>(g, bias, C, ldc, (float*)ws, n_slice);
  if (n_slice > 1)
    small_gemm_reduce_kernel<<<(unsigned)ceil_div(M * N, 256), 256, 0, st>>>((const float*)ws, M, N, n_slice, bias, C, ldc);
  count_launches(n_slice > 1 ? 2 : 1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}
