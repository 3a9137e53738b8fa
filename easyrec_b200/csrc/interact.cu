// K3 FM second order, sigmoid cross entropy, row marking.
//   FM:   layers/fm.py:20-26            0.5 * ((sum_f v)^2 - sum_f v^2)
//   loss: builders/loss_builder.py:36-39 tf.losses.sigmoid_cross_entropy
// Elementwise / small-reduction kernels: HBM- (in practice L2-) bound streaming
// reads of the [B, F*D] group matrix that K2 just wrote.
#include "common.cuh"

namespace er {

// one thread per (sample, 4-column chunk); D % 4 == 0, x 16 B aligned
__global__ void __launch_bounds__(256)
    fm_fwd_vec_kernel(const float* __restrict__ x, int64_t batch, int n_field, int dim,
                      int x_stride, float* __restrict__ y) {
  const int per = dim / 4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < batch * per;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per;
    const int c = (int)(t - b * per);
    const float4* row = reinterpret_cast<const float4*>(x + b * x_stride) + c;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
#pragma unroll 4
    for (int f = 0; f < n_field; ++f) {
      float4 v = row[f * per];
      s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y);
      s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w);
      q.x = __fadd_rn(q.x, __fmul_rn(v.x, v.x)); q.y = __fadd_rn(q.y, __fmul_rn(v.y, v.y));
      q.z = __fadd_rn(q.z, __fmul_rn(v.z, v.z)); q.w = __fadd_rn(q.w, __fmul_rn(v.w, v.w));
    }
    float4 o;
    o.x = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.x, s.x), q.x));
    o.y = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.y, s.y), q.y));
    o.z = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.z, s.z), q.z));
    o.w = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.w, s.w), q.w));
    reinterpret_cast<float4*>(y + b * dim)[c] = o;
  }
}

__global__ void __launch_bounds__(256)
    fm_fwd_scalar_kernel(const float* __restrict__ x, int64_t batch, int n_field, int dim,
                         int x_stride, float* __restrict__ y) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < batch * dim;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / dim;
    const int c = (int)(t - b * dim);
    float s = 0.f, q = 0.f;
    for (int f = 0; f < n_field; ++f) {
      float v = x[b * x_stride + f * dim + c];
      s = __fadd_rn(s, v);
      q = __fadd_rn(q, __fmul_rn(v, v));
    }
    y[b * dim + c] = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s, s), q));
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
    fm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, int64_t batch,
                  int n_field, int dim, int x_stride, float* __restrict__ gx, int gx_stride,
                  int accumulate) {
  const int per = dim / VEC;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < batch * per;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per;
    const int c = (int)(t - b * per) * VEC;
    float s[VEC], g[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      s[k] = 0.f;
      g[k] = gy[b * dim + c + k];
    }
    for (int f = 0; f < n_field; ++f) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) s[k] = __fadd_rn(s[k], x[b * x_stride + f * dim + c + k]);
    }
    for (int f = 0; f < n_field; ++f) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const int64_t xi = b * x_stride + f * dim + c + k;
        const int64_t gi = b * gx_stride + f * dim + c + k;
        float v = __fmul_rn(g[k], __fsub_rn(s[k], x[xi]));
        gx[gi] = accumulate ? __fadd_rn(gx[gi], v) : v;
      }
    }
  }
}

// one CTA-wide reduction of the loss; per-sample probs and dL/dlogit
__global__ void __launch_bounds__(256)
    sigmoid_ce_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                      const float* __restrict__ weights, int64_t batch, float inv_count,
                      float* __restrict__ loss_out, float* __restrict__ probs,
                      float* __restrict__ g_logits) {
  __shared__ float s_red[256];
  float acc = 0.f;
  for (int64_t b = threadIdx.x; b < batch; b += blockDim.x) {
    const float x = logits[b], z = labels[b];
    const float w = weights ? weights[b] : 1.0f;
    // max(x,0) - x*z + log1p(exp(-|x|))   (tf.nn.sigmoid_cross_entropy_with_logits)
    const float l = fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
    acc += w * l;
    const float p = 1.0f / (1.0f + expf(-x));
    if (probs) probs[b] = p;
    if (g_logits) g_logits[b] = w * (p - z) * inv_count;
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss_out) *loss_out = s_red[0] * inv_count;
}

__global__ void __launch_bounds__(256)
    mark_rows_kernel(const int64_t* __restrict__ rows, int64_t cap, const int32_t* __restrict__ n_dev,
                     int64_t n_rows, uint8_t* __restrict__ touched, uint8_t value) {
  const int64_t n = n_dev ? (int64_t)(*n_dev < cap ? *n_dev : cap) : cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[i];
    if (r >= 0 && r < n_rows) touched[r] = value;
  }
}

}  // namespace er

extern "C" int er_fm_fwd(const float* x, int64_t batch, int32_t n_field, int32_t dim,
                         int32_t x_stride, float* y, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && y, "null argument");
  ER_REQUIRE(batch >= 0 && n_field > 0 && dim > 0 && x_stride >= n_field * dim, "bad shape");
  if (batch == 0) return ER_OK;
  cudaStream_t st = as_stream(stream);
  const bool vec = dim % 4 == 0 && x_stride % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(y) % 16 == 0;
  if (vec)
    fm_fwd_vec_kernel<<<grid_for(batch * (dim / 4), 256, 8), 256, 0, st>>>(x, batch, n_field, dim,
                                                                          x_stride, y);
  else
    fm_fwd_scalar_kernel<<<grid_for(batch * dim, 256, 8), 256, 0, st>>>(x, batch, n_field, dim,
                                                                       x_stride, y);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_fm_bwd(const float* x, const float* gy, int64_t batch, int32_t n_field,
                         int32_t dim, int32_t x_stride, float* gx, int32_t gx_stride,
                         int32_t accumulate, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && gy && gx, "null argument");
  ER_REQUIRE(batch >= 0 && n_field > 0 && dim > 0 && x_stride >= n_field * dim &&
                 gx_stride >= n_field * dim,
             "bad shape");
  if (batch == 0) return ER_OK;
  cudaStream_t st = as_stream(stream);
  if (dim % 4 == 0)
    fm_bwd_kernel<4><<<grid_for(batch * (dim / 4), 256, 8), 256, 0, st>>>(
        x, gy, batch, n_field, dim, x_stride, gx, gx_stride, accumulate);
  else
    fm_bwd_kernel<1><<<grid_for(batch * dim, 256, 8), 256, 0, st>>>(
        x, gy, batch, n_field, dim, x_stride, gx, gx_stride, accumulate);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_sigmoid_ce_fwd_bwd(const float* logits, const float* labels,
                                     const float* weights, int64_t batch, float inv_count,
                                     float* loss_out, float* probs, float* g_logits,
                                     er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(logits && labels, "null argument");
  ER_REQUIRE(batch > 0, "batch must be positive");
  sigmoid_ce_kernel<<<1, 256, 0, as_stream(stream)>>>(logits, labels, weights, batch, inv_count,
                                                      loss_out, probs, g_logits);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_mark_rows(const int64_t* rows, int64_t n_lookups_cap, const int32_t* n_dev,
                            int64_t n_rows, uint8_t* touched, int32_t value,
                            er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(rows && touched, "null argument");
  if (n_lookups_cap <= 0) return ER_OK;
  mark_rows_kernel<<<grid_for(n_lookups_cap, 256, 8), 256, 0, as_stream(stream)>>>(
      rows, n_lookups_cap, n_dev, n_rows, touched, (uint8_t)value);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}
