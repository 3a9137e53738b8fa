/*
 * er_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the reference's sparse-embedding hot path
 * (alibaba/EasyRec @ bd230cb), used only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs as the checker.  Nothing under
 * easyrec_b200/ may import or link it.
 *
 * PINNING STATUS (the table in DESIGN.md section 4 is authoritative):
 *   - Fingerprint64: pinned by TensorFlow's own known-answer tests for inputs up to
 *     16 bytes (string_to_hash_bucket_op_test.py 'a','b','c','d'; fingerprint_test.cc
 *     "Hello","World"; the hashed feature column test 'omar','stringer','marlo' and the
 *     int64 path 101,201,301; lookup_ops OOV buckets; docstring examples) --
 *     tests/golden/reference_kats.json.  Longer inputs follow the published farmhashna
 *     code, cross-checked only against an independently written second implementation:
 *     parity unpinned.
 *   - lookup + pooling: the reference's numeric tests (test/embed_test.py:23-86, :88-151),
 *     TensorFlow's SafeEmbeddingLookupSparseTest case table, and the reference's
 *     embedding_lookup_ragged executed on a numpy shim.
 *   - shard rule (owner = id mod N, local row = id div N): embedding_parallel_lookup
 *     executed on two ranks (tests/golden/make_lookup_golden.py).
 *   - sparse Adagrad: the constants of TensorFlow's adagrad_test.py; lazy Adam:
 *     compat/adam_s.py _apply_sparse_shared executed (tests/golden/make_formula_golden.py).
 *   - sigmoid cross entropy and the summation order of duplicated rows: parity unpinned (TensorFlow-owned, no
 *     known-answer vector available offline).
 *
 * Third-party arithmetic restated here (absent from /root/reference):
 *   TensorFlow 1.15.5 / 2.12.0 (docker/Dockerfile:1, docker/Dockerfile_tf212:1):
 *   StringToHashBucketFast (farmhash::Fingerprint64), AsString("%lld"),
 *   safe_embedding_lookup_sparse, SparseApplyAdagrad, sigmoid_cross_entropy.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* FarmHash farmhashna::Hash64 (== tensorflow::Fingerprint64).          */
/* Call sites in the reference: feature_column_v2.py:3915-3921,         */
/* layers/input_layer.py:235,240.                                       */
/* ------------------------------------------------------------------ */
#define K0 0xc3a5c85c97cb3127ULL
#define K1 0xb492b66fbe98f273ULL
#define K2 0x9ae16a3b2f90404fULL

static uint64_t ld64(const uint8_t* p) {
  uint64_t r = 0;
  for (int i = 7; i >= 0; --i) r = (r << 8) | p[i];
  return r;
}
static uint64_t ld32(const uint8_t* p) {
  return (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24);
}
static uint64_t ror(uint64_t v, int s) { return s == 0 ? v : (v >> s) | (v << (64 - s)); }
static uint64_t smix(uint64_t v) { return v ^ (v >> 47); }
static uint64_t h16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= (a >> 47);
  uint64_t b = (v ^ a) * mul;
  b ^= (b >> 47);
  b *= mul;
  return b;
}
static void weak_seeds(const uint8_t* s, uint64_t a, uint64_t b, uint64_t* o1, uint64_t* o2) {
  uint64_t w = ld64(s), x = ld64(s + 8), y = ld64(s + 16), z = ld64(s + 24);
  a += w;
  b = ror(b + a + z, 21);
  uint64_t c = a;
  a += x;
  a += y;
  b += ror(a, 44);
  *o1 = a + z;
  *o2 = b + c;
}

uint64_t oracle_fingerprint64(const uint8_t* s, size_t len) {
  if (len <= 32) {
    if (len <= 16) {
      if (len >= 8) {
        uint64_t mul = K2 + len * 2;
        uint64_t a = ld64(s) + K2;
        uint64_t b = ld64(s + len - 8);
        uint64_t c = ror(b, 37) * mul + a;
        uint64_t d = (ror(a, 25) + b) * mul;
        return h16(c, d, mul);
      }
      if (len >= 4) {
        uint64_t mul = K2 + len * 2;
        uint64_t a = ld32(s);
        return h16(len + (a << 3), ld32(s + len - 4), mul);
      }
      if (len > 0) {
        uint8_t a = s[0], b = s[len >> 1], c = s[len - 1];
        uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
        uint32_t z = (uint32_t)len + ((uint32_t)c << 2);
        return smix(y * K2 ^ z * K0) * K2;
      }
      return K2;
    }
    {
      uint64_t mul = K2 + len * 2;
      uint64_t a = ld64(s) * K1;
      uint64_t b = ld64(s + 8);
      uint64_t c = ld64(s + len - 8) * mul;
      uint64_t d = ld64(s + len - 16) * K2;
      return h16(ror(a + b, 43) + ror(c, 30) + d, a + ror(b + K2, 18) + c, mul);
    }
  }
  if (len <= 64) {
    uint64_t mul = K2 + len * 2;
    uint64_t a = ld64(s) * K2;
    uint64_t b = ld64(s + 8);
    uint64_t c = ld64(s + len - 8) * mul;
    uint64_t d = ld64(s + len - 16) * K2;
    uint64_t y = ror(a + b, 43) + ror(c, 30) + d;
    uint64_t z = h16(y, a + ror(b + K2, 18) + c, mul);
    uint64_t e = ld64(s + 16) * mul;
    uint64_t f = ld64(s + 24);
    uint64_t g = (y + ld64(s + len - 32)) * mul;
    uint64_t h = (z + ld64(s + len - 24)) * mul;
    return h16(ror(e + f, 43) + ror(g, 30) + h, e + ror(f + a, 18) + g, mul);
  }
  {
    const uint64_t seed = 81;
    uint64_t x = seed, y = seed * K1 + 113, z = smix(y * K2 + 113) * K2;
    uint64_t v1 = 0, v2 = 0, w1 = 0, w2 = 0, t;
    const uint8_t* end = s + ((len - 1) / 64) * 64;
    const uint8_t* last64 = end + ((len - 1) & 63) - 63;
    x = x * K2 + ld64(s);
    do {
      x = ror(x + y + v1 + ld64(s + 8), 37) * K1;
      y = ror(y + v2 + ld64(s + 48), 42) * K1;
      x ^= w2;
      y += v1 + ld64(s + 40);
      z = ror(z + w1, 33) * K1;
      weak_seeds(s, v2 * K1, x + w1, &v1, &v2);
      weak_seeds(s + 32, z + w2, y + ld64(s + 16), &w1, &w2);
      t = z; z = x; x = t;
      s += 64;
    } while (s != end);
    uint64_t mul = K1 + ((z & 0xff) << 1);
    s = last64;
    w1 += ((len - 1) & 63);
    v1 += w1;
    w1 += v1;
    x = ror(x + y + v1 + ld64(s + 8), 37) * mul;
    y = ror(y + v2 + ld64(s + 48), 42) * mul;
    x ^= w2 * 9;
    y += v1 * 9 + ld64(s + 40);
    z = ror(z + w1, 33) * mul;
    weak_seeds(s, v2 * mul, x + w1, &v1, &v2);
    weak_seeds(s + 32, z + w2, y + ld64(s + 16), &w1, &w2);
    t = z; z = x; x = t;
    return h16(h16(v1, w1, mul) + smix(y) * K0 + z, h16(v2, w2, mul) + x, mul);
  }
}

/* ------------------------------------------------------------------ */
/* raw id -> table row.  modes mirror er_bucket_mode but */
/* are restated from the reference, not from the product header:        */
/*  0: Fingerprint64(as_string(v)) % hash_bucket_size                    */
/*     (input/input.py:356-376,541-543; feature_column_v2.py:3915-3921)  */
/*  1: v floormod num_buckets (input/parquet_input.py:221)               */
/*  2: identity; -1 dropped (feature_column_v2.py:2566-2585); out of     */
/*     range -> 0 (feature_column_v2.py:4268-4292)                       */
/*  3: passthrough, negative dropped                                     */
/*  4: one-row table: every value >= 0 is row 0 (input/input.py:648-673) */
/* sharding: owner = r mod N, local = int64(r / N)                       */
/*     (compat/feature_column/feature_column.py:296,317)                 */
/* ------------------------------------------------------------------ */
void oracle_bucketize(const int64_t* ids, int64_t n, const int32_t* mode, const int64_t* nb,
                      const int64_t* offset, const int32_t* shard_n, int64_t* rows,
                      int32_t* owner) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int64_t v = ids[i], r = 0;
    int drop = 0;
    if (mode[i] == 0) {
      char buf[32];
      int len = snprintf(buf, sizeof buf, "%lld", (long long)v); /* tf.as_string(int64) */
      uint64_t h = oracle_fingerprint64((const uint8_t*)buf, (size_t)len);
      r = (int64_t)(h % (uint64_t)nb[i]);
    } else if (mode[i] == 1) {
      int64_t m = v % nb[i];
      r = m < 0 ? m + nb[i] : m; /* python / tf floormod */
    } else if (mode[i] == 2) {
      drop = (v == -1);
      r = (v < 0 || v >= nb[i]) ? 0 : v;
    } else if (mode[i] == 4) { /* one-row table (RawFeature projection, raw_input_dim 1: id 0) */
      drop = v < 0;
      r = 0;
    } else {
      drop = v < 0;
      r = v;
    }
    int32_t own = 0;
    if (shard_n && shard_n[i] > 1) {
      own = (int32_t)(r % shard_n[i]);
      r = (int64_t)((double)r / (double)shard_n[i]); /* tf: cast(recv_ids / N, int64) */
    }
    rows[i] = drop ? -1 : offset[i] + r;
    if (owner) owner[i] = drop ? -1 : own;
  }
}

/* cumsum(lens) (feature_column.py:264) */
void oracle_csr_from_lens(const int32_t* lens, int64_t n_seg, int32_t* row_ptr, int32_t* seg_ids) {
  row_ptr[0] = 0;
  for (int64_t s = 0; s < n_seg; ++s) {
    row_ptr[s + 1] = row_ptr[s] + lens[s];
    if (seg_ids)
      for (int32_t j = row_ptr[s]; j < row_ptr[s + 1]; ++j) seg_ids[j] = (int32_t)s;
  }
}

/* ------------------------------------------------------------------ */
/* safe_embedding_lookup_sparse (compat/embedding_ops.py:37-162) +      */
/* combiner math (compat/feature_column/feature_column.py:202-244).     */
/* combiner per segment: 0 sum, 1 mean, 2 sqrtn.  out is [n_seg, dim].  */
/* ------------------------------------------------------------------ */
void oracle_embedding_fwd(const float* table, int32_t dim, int32_t row_stride, const int64_t* rows,
                          const float* weights, const int32_t* row_ptr, int64_t n_seg,
                          const int32_t* combiner, float* out, float* seg_scale) {
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < n_seg; ++s) {
    float* o = out + s * dim;
    for (int c = 0; c < dim; ++c) o[c] = 0.f;
    float wsum = 0.f, w2sum = 0.f;
    int comb = combiner[s];
    for (int32_t j = row_ptr[s]; j < row_ptr[s + 1]; ++j) {
      int64_t r = rows[j];
      float w = weights ? weights[j] : 1.0f;
      if (r < 0) continue;                      /* _prune_invalid_ids */
      if (comb != 0 && !(w > 0.f)) continue;    /* _prune_invalid_weights */
      const float* e = table + r * row_stride;
      for (int c = 0; c < dim; ++c) {
        float t = weights ? e[c] * w : e[c]; /* embeddings *= weights, then segment_sum */
        o[c] = o[c] + t;
      }
      wsum += w;
      { float t = w * w; w2sum += t; }
    }
    float scale = 1.f;
    if (comb == 1) {
      if (wsum != 0.f) { for (int c = 0; c < dim; ++c) o[c] = o[c] / wsum; scale = 1.f / wsum; }
      else { for (int c = 0; c < dim; ++c) o[c] = 0.f; scale = 0.f; }   /* div_no_nan / empty row */
    } else if (comb == 2) {
      float d = sqrtf(w2sum);
      if (d != 0.f) { for (int c = 0; c < dim; ++c) o[c] = o[c] / d; scale = 1.f / d; }
      else { for (int c = 0; c < dim; ++c) o[c] = 0.f; scale = 0.f; }
    }
    if (seg_scale) seg_scale[s] = scale;
  }
}

/* ------------------------------------------------------------------ */
/* backward: IndexedSlices -> dedup (sum duplicates in lookup order,    */
/* TF Optimizer._deduplicate_indexed_slices) -> sparse apply.           */
/* kind: 0 sgd, 1 adagrad (acc+=g^2; w-=lr*g*rsqrt(acc)), 2 lazy adam   */
/* (compat/adam_s.py:185-213).  gseg: dL/d(pooled) per segment.         */
/* uniq_rows/uniq_grads (optional): dedup result sorted by row.          */
/* returns the number of distinct rows.                                  */
/* ------------------------------------------------------------------ */
typedef struct { int64_t row; int64_t pos; } rp_t;
static int rp_cmp(const void* a, const void* b) {
  const rp_t* x = (const rp_t*)a; const rp_t* y = (const rp_t*)b;
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

int64_t oracle_embedding_bwd(float* table, float* s0, float* s1, int32_t dim, int32_t row_stride,
                             const int64_t* rows, const float* weights, const int32_t* seg_of,
                             int64_t n_lookups, const float* gseg, const float* seg_scale,
                             int32_t kind, float lr, float beta1, float beta2, float eps,
                             float beta1_power, float beta2_power, float grad_scale,
                             int64_t* uniq_rows, float* uniq_grads) {
  rp_t* a = (rp_t*)malloc(sizeof(rp_t) * (size_t)(n_lookups > 0 ? n_lookups : 1));
  int64_t m = 0;
  for (int64_t l = 0; l < n_lookups; ++l)
    if (rows[l] >= 0) { a[m].row = rows[l]; a[m].pos = l; ++m; }
  qsort(a, (size_t)m, sizeof(rp_t), rp_cmp);
  float* g = (float*)malloc(sizeof(float) * (size_t)dim);
  float lr_t = lr;
  if (kind == 2) lr_t = lr * sqrtf(1.0f - beta2_power) / (1.0f - beta1_power);
  int64_t u = 0;
  for (int64_t i = 0; i < m;) {
    int64_t j = i;
    for (int c = 0; c < dim; ++c) g[c] = 0.f;
    while (j < m && a[j].row == a[i].row) {
      int64_t l = a[j].pos;
      int32_t s = seg_of ? seg_of[l] : (int32_t)l;
      float coef = weights ? weights[l] : 1.0f;
      if (seg_scale) coef = coef * seg_scale[s];
      for (int c = 0; c < dim; ++c) {
        float t = gseg[(int64_t)s * dim + c] * coef;
        g[c] = g[c] + t;
      }
      ++j;
    }
    for (int c = 0; c < dim; ++c) g[c] = g[c] * grad_scale;
    if (uniq_rows) {
      uniq_rows[u] = a[i].row;
      memcpy(uniq_grads + u * dim, g, sizeof(float) * (size_t)dim);
    }
    if (table) {
      int64_t off = a[i].row * row_stride;
      for (int c = 0; c < dim; ++c) {
        float gg = g[c];
        if (kind == 1) {
          float g2 = gg * gg;
          s0[off + c] = s0[off + c] + g2;
          float step = (lr * gg) * (1.0f / sqrtf(s0[off + c]));
          table[off + c] = table[off + c] - step;
        } else if (kind == 2) {
          float m1 = gg * (1.0f - beta1), m2 = s0[off + c] * beta1;
          float mp = m1 + m2;
          float v1 = (gg * gg) * (1.0f - beta2), v2 = s1[off + c] * beta2;
          float vp = v1 + v2;
          s0[off + c] = mp;
          s1[off + c] = vp;
          float step = (-lr_t * mp) / (sqrtf(vp) + eps);
          table[off + c] = table[off + c] + step;
        } else if (kind == 4) {   /* tf.train.MomentumOptimizer, use_nesterov false (momentum in beta1):
                                     accum = accum * momentum + g ; var -= lr * accum */
          float acc = s0[off + c] * beta1;
          acc = acc + gg;
          s0[off + c] = acc;
          float step = lr * acc;
          table[off + c] = table[off + c] - step;
        } else {
          float step = lr * gg;
          table[off + c] = table[off + c] - step;
        }
      }
    }
    ++u;
    i = j;
  }
  free(g);
  free(a);
  return u;
}

/* FM second order, layers/fm.py:20-26.  x [B, F*D] -> y [B, D] */
void oracle_fm_fwd(const float* x, int64_t batch, int32_t n_field, int32_t dim, float* y) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b)
    for (int c = 0; c < dim; ++c) {
      float s = 0.f, q = 0.f;
      for (int f = 0; f < n_field; ++f) {
        float v = x[(b * n_field + f) * dim + c];
        s += v;
        { float t = v * v; q += t; }
      }
      float ss = s * s;
      y[b * dim + c] = 0.5f * (ss - q);
    }
}
void oracle_fm_bwd(const float* x, const float* gy, int64_t batch, int32_t n_field, int32_t dim,
                   float* gx) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b)
    for (int c = 0; c < dim; ++c) {
      float s = 0.f;
      for (int f = 0; f < n_field; ++f) s += x[(b * n_field + f) * dim + c];
      for (int f = 0; f < n_field; ++f) {
        int64_t i = (b * n_field + f) * dim + c;
        gx[i] = gy[b * dim + c] * (s - x[i]);
      }
    }
}

/* tf.losses.sigmoid_cross_entropy (builders/loss_builder.py:36-39):
 * sum_b w*(max(x,0)-x*z+log1p(exp(-|x|))) / count_nonzero(w) */
double oracle_sigmoid_ce(const float* logits, const float* labels, const float* weights,
                         int64_t batch, float* probs, float* g_logits) {
  double acc = 0.0;
  int64_t nz = 0;
  for (int64_t b = 0; b < batch; ++b) nz += (!weights || weights[b] != 0.f);
  double inv = nz > 0 ? 1.0 / (double)nz : 0.0;
  for (int64_t b = 0; b < batch; ++b) {
    double x = logits[b], z = labels[b], w = weights ? weights[b] : 1.0;
    acc += w * (fmax(x, 0.0) - x * z + log1p(exp(-fabs(x))));
    double p = 1.0 / (1.0 + exp(-x));
    if (probs) probs[b] = (float)p;
    if (g_logits) g_logits[b] = (float)(w * (p - z) * inv);
  }
  return acc * inv;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* thread count of the OpenMP loops above (bench.py picks the fastest setting on the box) */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
