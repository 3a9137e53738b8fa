// K3 FM second order, sigmoid cross entropy, row marking.
//   FM:   layers/fm.py:20-26            0.5 * ((sum_f v)^2 - sum_f v^2)
//   loss: builders/loss_builder.py:36-39 tf.losses.sigmoid_cross_entropy
// Elementwise / small-reduction kernels: HBM- (in practice L2-) bound streaming
// reads of the [B, F*D] group matrix that K2 just wrote.
#include <algorithm>

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

// one CTA-wide reduction of the loss (1024 threads; fixed shuffle + shared-memory tree: deterministic);
// per-sample probs and dL/dlogit
__global__ void __launch_bounds__(1024)
    sigmoid_ce_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                      const float* __restrict__ weights, int64_t batch, float inv_count,
                      float* __restrict__ loss_out, float* __restrict__ probs,
                      float* __restrict__ g_logits) {
  er_pdl_wait();
  __shared__ float s_red[32];
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
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = s_red[threadIdx.x];
    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0 && loss_out) *loss_out = t * inv_count;
  }
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

// ---- FM block, one warp per sample row --------------------------------------------------------
// The [F*D] row is read once, coalesced (lane + 32 j float4s), and stays in registers.  32 % (D/4) == 0,
// so every float4 a lane holds belongs to the same 4-column chunk (lane % (D/4)) and the per-chunk
// field sums are an xor-shuffle tree over the lanes of equal residue (fixed order: deterministic).
//   fwd: y = 0.5((sum_f x)^2 - sum_f x^2);  sumsq = sum_{b,f,d} x^2 (the embedding-regulariser term of
//        layers/input_layer.py:369-375, free here because FM needs sum_f x^2 anyway)
//   bwd: gx = g_pass + gy*(S - x) + coef*x   -- the three gradients that reach the group matrix (deep
//        tower input, FM, regulariser) in ONE pass instead of three kernels + two autograd adds.
template <int J>
__device__ __forceinline__ void fm_row_load(const float4* __restrict__ row, int n4, int lane, float4 (&v)[J]) {
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int i = lane + 32 * j;
    v[j] = i < n4 ? row[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ float4 f4_xor_add(float4 v, int o) {
  v.x = __fadd_rn(v.x, __shfl_xor_sync(0xffffffffu, v.x, o));
  v.y = __fadd_rn(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
  v.z = __fadd_rn(v.z, __shfl_xor_sync(0xffffffffu, v.z, o));
  v.w = __fadd_rn(v.w, __shfl_xor_sync(0xffffffffu, v.w, o));
  return v;
}

template <int J>
__global__ void __launch_bounds__(256)
    fm_block_fwd_kernel(const float* __restrict__ x, int64_t batch, int n_field, int dim, int x_stride,
                        float* __restrict__ y, float* __restrict__ partials, unsigned int* counter,
                        float* __restrict__ sumsq_out) {
  er_pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = dim >> 2, n4 = n_field * d4;
  __shared__ float s_w[8];
  __shared__ int s_last;
  float warp_sq = 0.f;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < batch; b += (int64_t)gridDim.x * 8) {
    float4 v[J];
    fm_row_load<J>(reinterpret_cast<const float4*>(x + b * x_stride), n4, lane, v);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      s.x = __fadd_rn(s.x, v[j].x); s.y = __fadd_rn(s.y, v[j].y);
      s.z = __fadd_rn(s.z, v[j].z); s.w = __fadd_rn(s.w, v[j].w);
      q.x = __fadd_rn(q.x, __fmul_rn(v[j].x, v[j].x)); q.y = __fadd_rn(q.y, __fmul_rn(v[j].y, v[j].y));
      q.z = __fadd_rn(q.z, __fmul_rn(v[j].z, v[j].z)); q.w = __fadd_rn(q.w, __fmul_rn(v[j].w, v[j].w));
    }
    for (int o = 16; o >= d4; o >>= 1) {
      s = f4_xor_add(s, o);
      q = f4_xor_add(q, o);
    }
    if (lane < d4) {
      float4 o4;
      o4.x = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.x, s.x), q.x));
      o4.y = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.y, s.y), q.y));
      o4.z = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.z, s.z), q.z));
      o4.w = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(s.w, s.w), q.w));
      reinterpret_cast<float4*>(y + b * dim)[lane] = o4;
    }
    if (partials) {
      float t = (lane < d4) ? __fadd_rn(__fadd_rn(q.x, q.y), __fadd_rn(q.z, q.w)) : 0.f;
      for (int o = 16; o >= 1; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
      warp_sq = __fadd_rn(warp_sq, t);
    }
  }
  if (!partials) return;
  if (lane == 0) s_w[warp] = warp_sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t = __fadd_rn(t, s_w[w]);
    __stcg(partials + blockIdx.x, t);
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {   // last CTA: fixed-order sum of the CTA partials; leaves the counter at 0 for the next call
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t = __fadd_rn(t, __ldcg(partials + i));
    for (int o = 16; o >= 1; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
    __syncthreads();
    if (lane == 0) s_w[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot = __fadd_rn(tot, s_w[w]);
      *sumsq_out = tot;
      *counter = 0u;
    }
  }
}

template <int J>
__global__ void __launch_bounds__(256)
    fm_block_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                        const float* __restrict__ g_pass, const float* __restrict__ coef_dev, float coef_mul,
                        int64_t batch, int n_field, int dim, int x_stride, int gp_stride,
                        float* __restrict__ gx, int gx_stride) {
  er_pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = dim >> 2, n4 = n_field * d4;
  const float coef = coef_dev ? __fmul_rn(*coef_dev, coef_mul) : 0.f;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < batch; b += (int64_t)gridDim.x * 8) {
    float4 v[J], gp[J];
    fm_row_load<J>(reinterpret_cast<const float4*>(x + b * x_stride), n4, lane, v);
    if (g_pass) fm_row_load<J>(reinterpret_cast<const float4*>(g_pass + b * gp_stride), n4, lane, gp);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy) g = reinterpret_cast<const float4*>(gy + b * dim)[lane % d4];
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      s.x = __fadd_rn(s.x, v[j].x); s.y = __fadd_rn(s.y, v[j].y);
      s.z = __fadd_rn(s.z, v[j].z); s.w = __fadd_rn(s.w, v[j].w);
    }
    for (int o = 16; o >= d4; o >>= 1) s = f4_xor_add(s, o);
    float4* orow = reinterpret_cast<float4*>(gx + b * gx_stride);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int i = lane + 32 * j;
      if (i < n4) {
        float4 r;
        r.x = __fadd_rn(__fmul_rn(g.x, __fsub_rn(s.x, v[j].x)), __fmul_rn(coef, v[j].x));
        r.y = __fadd_rn(__fmul_rn(g.y, __fsub_rn(s.y, v[j].y)), __fmul_rn(coef, v[j].y));
        r.z = __fadd_rn(__fmul_rn(g.z, __fsub_rn(s.z, v[j].z)), __fmul_rn(coef, v[j].z));
        r.w = __fadd_rn(__fmul_rn(g.w, __fsub_rn(s.w, v[j].w)), __fmul_rn(coef, v[j].w));
        if (g_pass) {
          r.x = __fadd_rn(r.x, gp[j].x); r.y = __fadd_rn(r.y, gp[j].y);
          r.z = __fadd_rn(r.z, gp[j].z); r.w = __fadd_rn(r.w, gp[j].w);
        }
        orow[i] = r;
      }
    }
  }
}

// ---- wide block: y[b] = sum_f x[b,f] (model/deepfm.py:62-63 reduce_sum over the wide group) and
// sumsq = sum x^2 (embedding regulariser), one warp per sample row; bwd: gx[b,f] = gy[b] + coef*x[b,f].
__global__ void __launch_bounds__(256)
    rowsum_block_fwd_kernel(const float* __restrict__ x, int64_t batch, int width, int x_stride,
                            float* __restrict__ y, float* __restrict__ partials, unsigned int* counter,
                            float* __restrict__ sumsq_out) {
  er_pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float s_w[8];
  __shared__ int s_last;
  float warp_sq = 0.f;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < batch; b += (int64_t)gridDim.x * 8) {
    float sv = 0.f, q = 0.f;
    for (int f = lane; f < width; f += 32) {
      const float v = x[b * x_stride + f];
      sv = __fadd_rn(sv, v);
      q = __fadd_rn(q, __fmul_rn(v, v));
    }
    for (int o = 16; o >= 1; o >>= 1) {
      sv = __fadd_rn(sv, __shfl_xor_sync(0xffffffffu, sv, o));
      q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, o));
    }
    if (lane == 0) y[b] = sv;
    warp_sq = __fadd_rn(warp_sq, q);
  }
  if (!partials) return;
  if (lane == 0) s_w[warp] = warp_sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t = __fadd_rn(t, s_w[w]);
    __stcg(partials + blockIdx.x, t);
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t = __fadd_rn(t, __ldcg(partials + i));
    for (int o = 16; o >= 1; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
    __syncthreads();
    if (lane == 0) s_w[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot = __fadd_rn(tot, s_w[w]);
      *sumsq_out = tot;
      *counter = 0u;
    }
  }
}

__global__ void __launch_bounds__(256)
    rowsum_block_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                            const float* __restrict__ coef_dev, float coef_mul, int64_t batch, int width,
                            int x_stride, float* __restrict__ gx, int gx_stride) {
  er_pdl_wait();
  const float coef = coef_dev ? __fmul_rn(*coef_dev, coef_mul) : 0.f;
  const int64_t total = batch * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / width;
    const int f = (int)(t - b * width);
    gx[b * gx_stride + f] = __fadd_rn(gy ? gy[b] : 0.f, __fmul_rn(coef, x[b * x_stride + f]));
  }
}

// ---- single-unit dense head (the logit layer, tf.layers.dense(units=1)): a GEMV, one warp per row ----
__global__ void __launch_bounds__(256)
    dense1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      int64_t batch, int width, int x_stride, float* __restrict__ y) {
  er_pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float b0 = bias ? bias[0] : 0.f;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < batch; b += (int64_t)gridDim.x * 8) {
    float acc = 0.f;
    for (int f = lane; f < width; f += 32) acc = fmaf(x[b * x_stride + f], w[f], acc);
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[b] = acc + b0;
  }
}
// gx[b,f] = g[b]*w[f];  gw[f] = sum_b g[b]*x[b,f], gb = sum_b g[b].  Warps stride over the rows with the
// column sums in registers (lane owns columns lane, lane+32, ...), the 8 warps of a CTA and then the CTAs are
// combined in index order (per-CTA partials, last CTA finishes): deterministic, no float atomics.
template <int JW>
__global__ void __launch_bounds__(256)
    dense1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ g,
                      int64_t batch, int width, int x_stride, float* __restrict__ gx, int gx_stride,
                      float* __restrict__ partials /* [grid][width+1] */, unsigned int* counter,
                      float* __restrict__ gw, float* __restrict__ gb) {
  er_pdl_wait();
  extern __shared__ float s_col[];   // [8][width+1], later [256] for the final combine
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wp = width + 1;
  float wv[JW], acc[JW];
#pragma unroll
  for (int j = 0; j < JW; ++j) {
    const int f = lane + 32 * j;
    wv[j] = f < width ? w[f] : 0.f;
    acc[j] = 0.f;
  }
  float accg = 0.f;
  const int64_t total_warps = (int64_t)gridDim.x * 8;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < batch; b += total_warps) {
    const float gb_ = g[b];
    accg += gb_;
#pragma unroll
    for (int j = 0; j < JW; ++j) {
      const int f = lane + 32 * j;
      if (f < width) {
        acc[j] = fmaf(gb_, x[b * x_stride + f], acc[j]);
        if (gx) gx[b * gx_stride + f] = gb_ * wv[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < JW; ++j) {
    const int f = lane + 32 * j;
    if (f < width) s_col[warp * wp + f] = acc[j];
  }
  if (lane == 0) s_col[warp * wp + width] = accg;
  __syncthreads();
  for (int f = threadIdx.x; f < wp; f += 256) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s_col[k * wp + f];
    __stcg(partials + (int64_t)blockIdx.x * wp + f, t);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // P threads per column split the CTA partials (interleaved), then combine in order
  const int P = max(1, 256 / wp);
  const int f = threadIdx.x % wp, part = threadIdx.x / wp;
  float t = 0.f;
  if (part < P) {
#pragma unroll 8
    for (int k = part; k < (int)gridDim.x; k += P) t += __ldcg(partials + (int64_t)k * wp + f);
  }
  __syncthreads();
  if (part < P) s_col[part * wp + f] = t;
  __syncthreads();
  if (part == 0) {
    for (int k = 1; k < P; ++k) t += s_col[k * wp + f];
    if (f < width) gw[f] = t;
    else if (gb) gb[0] = t;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// ---- column concat into a pitched buffer / split back (tf.concat(axis=1) of model/deepfm.py:76 and its
// gradient): one launch each way; the destination pitch is a multiple of 4 floats so the next dense layer's
// GEMM reads it in place, padding columns are written as zeros.
struct CatArgs {
  float* p[ER_MAX_CAT];        // source (concat) / destination (split) matrices [B, width_i]
  int width[ER_MAX_CAT];
  int stride[ER_MAX_CAT];
  int first[ER_MAX_CAT + 1];   // first column of each piece inside the wide matrix
  int n;
};
__global__ void __launch_bounds__(256)
    concat_cols_kernel(CatArgs a, int64_t batch, float* __restrict__ dst, int dst_stride) {
  er_pdl_wait();
  const int64_t total = batch * dst_stride;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / dst_stride;
    const int c = (int)(t - b * dst_stride);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < ER_MAX_CAT; ++i)
      if (i < a.n && c >= a.first[i] && c < a.first[i + 1]) v = a.p[i][b * a.stride[i] + (c - a.first[i])];
    dst[t] = v;
  }
}
__global__ void __launch_bounds__(256)
    split_cols_kernel(CatArgs a, int64_t batch, const float* __restrict__ src, int src_stride, int total_w) {
  er_pdl_wait();
  const int64_t total = batch * total_w;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / total_w;
    const int c = (int)(t - b * total_w);
    const float v = src[b * src_stride + c];
#pragma unroll
    for (int i = 0; i < ER_MAX_CAT; ++i)
      if (i < a.n && c >= a.first[i] && c < a.first[i + 1]) a.p[i][b * a.stride[i] + (c - a.first[i])] = v;
  }
}

static bool fm_block_shape_ok(int n_field, int dim) {
  const int d4 = dim / 4;
  return dim % 4 == 0 && d4 >= 1 && d4 <= 32 && (d4 & (d4 - 1)) == 0 && (int64_t)n_field * d4 <= 32 * 8;
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
  launch_pdl(sigmoid_ce_kernel, dim3(1), dim3(1024), 0, as_stream(stream), logits, labels, weights, batch, inv_count,
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

extern "C" size_t er_fm_block_workspace_bytes(int64_t batch) {
  (void)batch;
  return 16 + sizeof(float) * 4 * er::kSmCount * 2;
}

// ws must be zero-filled once by the caller (the kernel leaves its counter at zero).
extern "C" int er_fm_block_fwd(const float* x, int64_t batch, int32_t n_field, int32_t dim,
                               int32_t x_stride, float* y, float* sumsq_out, void* ws, size_t ws_bytes,
                               er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && y, "null argument");
  ER_REQUIRE(batch > 0 && n_field > 0 && x_stride >= n_field * dim, "bad shape");
  ER_REQUIRE(fm_block_shape_ok(n_field, dim), "dim must be 4*2^k (<= 128) and n_field*dim <= 1024");
  ER_REQUIRE(x_stride % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(y) % 16 == 0, "x / y must be 16-byte aligned rows");
  const int grid = (int)std::min<int64_t>(ceil_div(batch, 8), 4 * kSmCount);
  float* partials = nullptr;
  unsigned int* counter = nullptr;
  if (sumsq_out) {
    ER_REQUIRE(ws && ws_bytes >= er_fm_block_workspace_bytes(batch), "workspace too small");
    counter = reinterpret_cast<unsigned int*>(ws);
    partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 16);
  }
  const int J = (int)ceil_div((int64_t)n_field * (dim / 4), 32);
  cudaStream_t st = as_stream(stream);
#define ER_FM_FWD(JJ) launch_pdl(fm_block_fwd_kernel<JJ>, dim3(grid), dim3(256), 0, st, x, batch, (int)n_field, (int)dim, (int)x_stride, y, partials, counter, sumsq_out)
  switch (J) {
    case 1: ER_FM_FWD(1); break; case 2: ER_FM_FWD(2); break; case 3: ER_FM_FWD(3); break;
    case 4: ER_FM_FWD(4); break; case 5: ER_FM_FWD(5); break; case 6: ER_FM_FWD(6); break;
    case 7: ER_FM_FWD(7); break; default: ER_FM_FWD(8); break;
  }
#undef ER_FM_FWD
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_fm_block_bwd(const float* x, const float* gy, const float* g_pass,
                               const float* coef_dev, float coef_mul, int64_t batch, int32_t n_field,
                               int32_t dim, int32_t x_stride, int32_t g_pass_stride, float* gx,
                               int32_t gx_stride, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && gx, "null argument");
  ER_REQUIRE(batch > 0 && n_field > 0 && x_stride >= n_field * dim && gx_stride >= n_field * dim, "bad shape");
  ER_REQUIRE(!g_pass || g_pass_stride >= n_field * dim, "bad g_pass stride");
  ER_REQUIRE(fm_block_shape_ok(n_field, dim), "dim must be 4*2^k (<= 128) and n_field*dim <= 1024");
  ER_REQUIRE(x_stride % 4 == 0 && gx_stride % 4 == 0 && (!g_pass || g_pass_stride % 4 == 0) &&
                 reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(gx) % 16 == 0 &&
                 (!g_pass || reinterpret_cast<uintptr_t>(g_pass) % 16 == 0) &&
                 (!gy || reinterpret_cast<uintptr_t>(gy) % 16 == 0), "operands must be 16-byte aligned rows");
  const int grid = (int)std::min<int64_t>(ceil_div(batch, 8), 8 * kSmCount);
  const int J = (int)ceil_div((int64_t)n_field * (dim / 4), 32);
  cudaStream_t st = as_stream(stream);
#define ER_FM_BWD(JJ) launch_pdl(fm_block_bwd_kernel<JJ>, dim3(grid), dim3(256), 0, st, x, gy, g_pass, coef_dev, coef_mul, batch, (int)n_field, (int)dim, (int)x_stride, (int)g_pass_stride, gx, (int)gx_stride)
  switch (J) {
    case 1: ER_FM_BWD(1); break; case 2: ER_FM_BWD(2); break; case 3: ER_FM_BWD(3); break;
    case 4: ER_FM_BWD(4); break; case 5: ER_FM_BWD(5); break; case 6: ER_FM_BWD(6); break;
    case 7: ER_FM_BWD(7); break; default: ER_FM_BWD(8); break;
  }
#undef ER_FM_BWD
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_rowsum_block_fwd(const float* x, int64_t batch, int32_t width, int32_t x_stride, float* y,
                                   float* sumsq_out, void* ws, size_t ws_bytes, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && y, "null argument");
  ER_REQUIRE(batch > 0 && width > 0 && x_stride >= width, "bad shape");
  const int grid = (int)std::min<int64_t>(ceil_div(batch, 8), 4 * kSmCount);
  float* partials = nullptr;
  unsigned int* counter = nullptr;
  if (sumsq_out) {
    ER_REQUIRE(ws && ws_bytes >= er_fm_block_workspace_bytes(batch), "workspace too small");
    counter = reinterpret_cast<unsigned int*>(ws);
    partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 16);
  }
  launch_pdl(rowsum_block_fwd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), x, batch, (int)width, (int)x_stride, y,
             partials, counter, sumsq_out);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_rowsum_block_bwd(const float* x, const float* gy, const float* coef_dev, float coef_mul,
                                   int64_t batch, int32_t width, int32_t x_stride, float* gx,
                                   int32_t gx_stride, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && gx, "null argument");
  ER_REQUIRE(batch > 0 && width > 0 && x_stride >= width && gx_stride >= width, "bad shape");
  launch_pdl(rowsum_block_bwd_kernel, dim3(grid_for(batch * width, 256, 8)), dim3(256), 0, as_stream(stream), x, gy,
             coef_dev, coef_mul, batch, (int)width, (int)x_stride, gx, (int)gx_stride);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" size_t er_dense1_workspace_bytes(int32_t width) {
  return 16 + (size_t)er::kSmCount * ((size_t)width + 1) * sizeof(float);
}

extern "C" int er_dense1_fwd(const float* x, const float* w, const float* bias, int64_t batch, int32_t width,
                             int32_t x_stride, float* y, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && w && y, "null argument");
  ER_REQUIRE(batch > 0 && width > 0 && x_stride >= width, "bad shape");
  const int grid = (int)std::min<int64_t>(ceil_div(batch, 8), 4 * kSmCount);
  launch_pdl(dense1_fwd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), x, w, bias, batch, (int)width, (int)x_stride, y);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

// ws: er_dense1_workspace_bytes(width), first 16 bytes zero on first use (left zero).
extern "C" int er_dense1_bwd(const float* x, const float* w, const float* g, int64_t batch, int32_t width,
                             int32_t x_stride, float* gx, int32_t gx_stride, float* gw, float* gb, void* ws,
                             size_t ws_bytes, er_stream_t stream) {
  using namespace er;
  ER_REQUIRE(x && w && g && gw, "null argument");
  ER_REQUIRE(batch > 0 && width > 0 && width <= 255 && x_stride >= width && (!gx || gx_stride >= width), "bad shape (width <= 255)");
  ER_REQUIRE(ws && ws_bytes >= er_dense1_workspace_bytes(width), "workspace too small");
  const int grid = (int)std::min<int64_t>(ceil_div(batch, 8), kSmCount);
  const size_t smem = (size_t)std::max(8 * (width + 1), 256) * sizeof(float);
  float* part = reinterpret_cast<float*>(static_cast<char*>(ws) + 16);
  unsigned int* cnt = reinterpret_cast<unsigned int*>(ws);
  cudaStream_t st = as_stream(stream);
#define ER_D1(JJ) launch_pdl(dense1_bwd_kernel<JJ>, dim3(grid), dim3(256), smem, st, x, w, g, batch, (int)width, (int)x_stride, gx, (int)gx_stride, part, cnt, gw, gb)
  switch ((width + 31) / 32) {
    case 1: ER_D1(1); break; case 2: ER_D1(2); break; case 3: ER_D1(3); break; case 4: ER_D1(4); break;
    case 5: ER_D1(5); break; case 6: ER_D1(6); break; case 7: ER_D1(7); break; default: ER_D1(8); break;
  }
#undef ER_D1
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

static int cat_args(er::CatArgs* a, float* const* mats, const int32_t* widths, const int32_t* strides,
                    int32_t n) {
  using namespace er;
  ER_REQUIRE(mats && widths && strides, "null argument");
  ER_REQUIRE(n > 0 && n <= ER_MAX_CAT, "1..ER_MAX_CAT pieces");
  a->n = n;
  int col = 0;
  for (int i = 0; i < ER_MAX_CAT; ++i) {
    a->first[i] = col;
    if (i < n) {
      ER_REQUIRE(mats[i] && widths[i] > 0 && strides[i] >= widths[i], "bad piece");
      a->p[i] = mats[i];
      a->width[i] = widths[i];
      a->stride[i] = strides[i];
      col += widths[i];
    } else {
      a->p[i] = nullptr;
      a->width[i] = a->stride[i] = 0;
    }
  }
  a->first[ER_MAX_CAT] = col;
  for (int i = n; i <= ER_MAX_CAT; ++i) a->first[i] = col;
  return ER_OK;
}

extern "C" int er_concat_cols(const float* const* srcs, const int32_t* widths, const int32_t* strides,
                              int32_t n, int64_t batch, float* dst, int32_t dst_stride, er_stream_t stream) {
  using namespace er;
  CatArgs a;
  int rc = cat_args(&a, const_cast<float* const*>(srcs), widths, strides, n);
  if (rc != ER_OK) return rc;
  ER_REQUIRE(dst && batch > 0 && dst_stride >= a.first[ER_MAX_CAT], "bad destination");
  launch_pdl(concat_cols_kernel, dim3(grid_for(batch * dst_stride, 256, 8)), dim3(256), 0, as_stream(stream), a, batch, dst,
             (int)dst_stride);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}

extern "C" int er_split_cols(const float* src, int32_t src_stride, int64_t batch, float* const* dsts,
                             const int32_t* widths, const int32_t* strides, int32_t n, er_stream_t stream) {
  using namespace er;
  CatArgs a;
  int rc = cat_args(&a, dsts, widths, strides, n);
  if (rc != ER_OK) return rc;
  ER_REQUIRE(src && batch > 0 && src_stride >= a.first[ER_MAX_CAT], "bad source");
  launch_pdl(split_cols_kernel, dim3(grid_for(batch * a.first[ER_MAX_CAT], 256, 8)), dim3(256), 0, as_stream(stream), a,
             batch, src, (int)src_stride, a.first[ER_MAX_CAT]);
  count_launches(1);
  ER_CUDA_LAUNCH_CHECK();
  return ER_OK;
}
