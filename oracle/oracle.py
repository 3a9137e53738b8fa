"""TEST INFRASTRUCTURE -- numpy/ctypes front end of the CPU oracle (oracle/er_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; the product package (easyrec_b200/) never does.

The sparse path is the C restatement; the dense model pieces (DNN + batch-norm,
DeepFM head, DIN attention, DCN cross, MMoE, DSSM, losses) are restated in numpy
fp32 below, each citing the reference lines it follows.  Parity status:
TensorFlow cannot be imported in this container, so the oracle is pinned piece by piece -- to
TensorFlow's own frozen test vectors (Fingerprint64 up to 16 bytes, the hashed feature column, the
sparse Adagrad constants, the safe-lookup case table), to the reference's embed_test.py known
answers, and to golden vectors obtained by EXECUTING the reference's own function bodies on a numpy
shim (tests/golden/make_*_golden.py: FM, DCN cross, keras Cross, dot interaction, DIN attention,
MMoE, list-wise match loss, DeepFM head, lazy Adam, LR schedule, pooling, shard rule, Parquet
batches, checkpoint layout, GAUC).  Still unpinned: Fingerprint64 beyond 16 bytes, batch-norm /
dense arithmetic and sigmoid cross entropy (TensorFlow-owned), the summation order inside duplicated rows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liber_oracle.so')
_lib = None

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


def build():
  subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(_SO):
      build()
    _lib = ctypes.CDLL(_SO)
    _lib.oracle_fingerprint64.restype = ctypes.c_uint64
    _lib.oracle_fingerprint64.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    _lib.oracle_embedding_bwd.restype = c_i64
    _lib.oracle_sigmoid_ce.restype = ctypes.c_double
    _lib.oracle_num_threads.restype = c_i32
  return _lib


def _ptr(a):
  return None if a is None else a.ctypes.data_as(c_vp)


def _c(a, dtype):
  return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def num_threads():
  return int(lib().oracle_num_threads())


def set_num_threads(n):
  lib().oracle_set_num_threads(int(n))


def fingerprint64(data):
  if isinstance(data, str):
    data = data.encode('utf-8')
  return int(lib().oracle_fingerprint64(data, len(data)))


def bucketize(ids, mode, num_buckets, offset, shard_n=None):
  """Per-lookup arrays (broadcast scalars first). Returns (rows int64, owner int32)."""
  ids = _c(ids, np.int64)
  n = ids.size
  mode = _c(np.broadcast_to(mode, n), np.int32)
  nb = _c(np.broadcast_to(num_buckets, n), np.int64)
  off = _c(np.broadcast_to(offset, n), np.int64)
  sh = None if shard_n is None else _c(np.broadcast_to(shard_n, n), np.int32)
  rows = np.empty(n, np.int64)
  owner = np.empty(n, np.int32)
  lib().oracle_bucketize(_ptr(ids), c_i64(n), _ptr(mode), _ptr(nb), _ptr(off), _ptr(sh),
                         _ptr(rows), _ptr(owner))
  return rows, owner


def csr_from_lens(lens):
  lens = _c(lens, np.int32)
  n_seg = lens.size
  row_ptr = np.empty(n_seg + 1, np.int32)
  seg_ids = np.empty(max(int(lens.sum()), 1), np.int32)
  lib().oracle_csr_from_lens(_ptr(lens), c_i64(n_seg), _ptr(row_ptr), _ptr(seg_ids))
  return row_ptr, seg_ids[:int(lens.sum())]


def embedding_fwd(table, rows, row_ptr, combiner, weights=None):
  """table [V, D]; returns (pooled [n_seg, D], seg_scale [n_seg])."""
  table = _c(table, np.float32)
  rows = _c(rows, np.int64)
  row_ptr = _c(row_ptr, np.int32)
  n_seg = row_ptr.size - 1
  comb = _c(np.broadcast_to(combiner, n_seg), np.int32)
  weights = _c(weights, np.float32)
  dim = table.shape[1]
  out = np.empty((n_seg, dim), np.float32)
  scale = np.empty(n_seg, np.float32)
  lib().oracle_embedding_fwd(_ptr(table), c_i32(dim), c_i32(dim), _ptr(rows), _ptr(weights),
                             _ptr(row_ptr), c_i64(n_seg), _ptr(comb), _ptr(out), _ptr(scale))
  return out, scale


OPT_SGD, OPT_ADAGRAD, OPT_LAZY_ADAM = 0, 1, 2
OPT_MOMENTUM = 4   # tf.train.MomentumOptimizer: accum = accum * momentum + g, var -= lr * accum (momentum passed as beta1)


def embedding_bwd(table, s0, s1, rows, seg_of, gseg, kind, lr, weights=None, seg_scale=None,
                  beta1=0.9, beta2=0.999, eps=1e-8, beta1_power=0.9, beta2_power=0.999,
                  grad_scale=1.0, want_uniq=False):
  """In-place update of table/s0/s1 (float32 C-contiguous [V, D]); gseg [n_seg, D].

  Returns (n_uniq, uniq_rows, uniq_grads)."""
  assert table is None or (table.dtype == np.float32 and table.flags.c_contiguous)
  rows = _c(rows, np.int64)
  gseg = _c(gseg, np.float32)
  seg_of = _c(seg_of, np.int32)
  weights = _c(weights, np.float32)
  seg_scale = _c(seg_scale, np.float32)
  dim = gseg.shape[1]
  n = rows.size
  ur = np.empty(max(n, 1), np.int64) if want_uniq else None
  ug = np.empty((max(n, 1), dim), np.float32) if want_uniq else None
  u = lib().oracle_embedding_bwd(_ptr(table), _ptr(s0), _ptr(s1), c_i32(dim), c_i32(dim),
                                 _ptr(rows), _ptr(weights), _ptr(seg_of), c_i64(n), _ptr(gseg),
                                 _ptr(seg_scale), c_i32(kind), c_f32(lr), c_f32(beta1),
                                 c_f32(beta2), c_f32(eps), c_f32(beta1_power), c_f32(beta2_power),
                                 c_f32(grad_scale), _ptr(ur), _ptr(ug))
  if want_uniq:
    return int(u), ur[:u], ug[:u]
  return int(u), None, None


def adam_lr_t(lr, beta1_power, beta2_power):
  """lr * sqrt(1 - beta2_power) / (1 - beta1_power) in fp32, the order of the TF graph (compat/adam_s.py:193;
  TensorFlow's adam.py builds the same expression)."""
  f = np.float32
  return f(f(f(lr) * np.sqrt(f(1.0) - f(beta2_power))) / (f(1.0) - f(beta1_power)))


def embedding_bwd_adam_dense(table, m, v, rows, seg_of, gseg, lr, weights=None, seg_scale=None, beta1=0.9,
                             beta2=0.999, eps=1e-8, beta1_power=0.9, beta2_power=0.999, grad_scale=1.0):
  """tf.train.AdamOptimizer._apply_sparse_shared (TensorFlow 1.15 python/training/adam.py; the optimizer
  builders/optimizer_builder.py:61-66 constructs; its dense behaviour is described at compat/adam_s.py:74-81):

      m <- m*beta1 (ALL rows);  m[indices] += (1-beta1)*g        v likewise with g*g
      var <- var - lr_t * m / (sqrt(v) + eps)   (ALL rows)

  so a row with a gradient gets exactly the lazy row rule (same operations on the same operands) and every
  other row decays: m*beta1, v*beta2, var - lr_t*m/(sqrt(v)+eps).  In-place on float32 [V, D] arrays.
  Parity unpinned against TensorFlow itself (TensorFlow is absent here): restated from the TF source."""
  f = np.float32
  n_uniq, ur, _ = embedding_bwd(None, None, None, rows, seg_of, gseg, OPT_SGD, 0.0, weights=weights,
                                seg_scale=seg_scale, grad_scale=grad_scale, want_uniq=True)
  embedding_bwd(table, m, v, rows, seg_of, gseg, OPT_LAZY_ADAM, lr, weights=weights, seg_scale=seg_scale, beta1=beta1,
                beta2=beta2, eps=eps, beta1_power=beta1_power, beta2_power=beta2_power, grad_scale=grad_scale)
  cold = np.ones(table.shape[0], bool)
  cold[ur] = False
  lr_t = adam_lr_t(lr, beta1_power, beta2_power)
  mc = (m[cold] * f(beta1)).astype(np.float32)
  vc = (v[cold] * f(beta2)).astype(np.float32)
  m[cold] = mc
  v[cold] = vc
  table[cold] = table[cold] - (lr_t * mc) / (np.sqrt(vc) + f(eps))
  return n_uniq


def fm_fwd(x, n_field, dim):
  x = _c(x, np.float32)
  y = np.empty((x.shape[0], dim), np.float32)
  lib().oracle_fm_fwd(_ptr(x), c_i64(x.shape[0]), c_i32(n_field), c_i32(dim), _ptr(y))
  return y


def fm_bwd(x, gy, n_field, dim):
  x = _c(x, np.float32)
  gy = _c(gy, np.float32)
  gx = np.empty_like(x)
  lib().oracle_fm_bwd(_ptr(x), _ptr(gy), c_i64(x.shape[0]), c_i32(n_field), c_i32(dim), _ptr(gx))
  return gx


def sigmoid_ce(logits, labels, weights=None):
  """returns (loss float, probs, g_logits)."""
  logits = _c(logits, np.float32)
  labels = _c(labels, np.float32)
  weights = _c(weights, np.float32)
  probs = np.empty_like(logits)
  g = np.empty_like(logits)
  loss = lib().oracle_sigmoid_ce(_ptr(logits), _ptr(labels), _ptr(weights), c_i64(logits.size),
                                 _ptr(probs), _ptr(g))
  return float(loss), probs, g


# ---------------------------------------------------------------------------
# dense model pieces, numpy fp32
# ---------------------------------------------------------------------------
BN_EPS = 1e-3  # tf.layers.batch_normalization default epsilon
BN_MOMENTUM = 0.99


def dnn_forward(x, layers, training=True, last_no_act=False, last_no_bn=False):
  """layers/dnn.py:50-87: dense -> batch_norm (use_bn) -> relu, per layer.

  layers: list of dicts {W [in,out], b [out], gamma, beta, mean, var} (bn keys optional).
  Returns (y, cache) where cache feeds dnn_backward."""
  cache = []
  n = len(layers)
  for i, L in enumerate(layers):
    z = (x @ L['W'] + L['b']).astype(np.float32)
    use_bn = ('gamma' in L) and not (last_no_bn and i == n - 1)
    act = not (last_no_act and i == n - 1)
    if use_bn:
      if training:
        mu = z.mean(axis=0, dtype=np.float32)
        var = ((z - mu) ** 2).mean(axis=0, dtype=np.float32)
      else:
        mu, var = L['mean'], L['var']
      rstd = (1.0 / np.sqrt(var + np.float32(BN_EPS))).astype(np.float32)
      xhat = ((z - mu) * rstd).astype(np.float32)
      h = (xhat * L['gamma'] + L['beta']).astype(np.float32)
    else:
      mu = var = rstd = xhat = None
      h = z
    y = np.maximum(h, 0).astype(np.float32) if act else h
    cache.append(dict(x=x, z=z, xhat=xhat, rstd=rstd, h=h, use_bn=use_bn, act=act, mu=mu,
                      var=var))
    x = y
  return x, cache


def dnn_backward(gy, layers, cache):
  """returns (gx, grads) with grads[i] = dict(W, b, gamma, beta)."""
  grads = [None] * len(layers)
  for i in reversed(range(len(layers))):
    L, C = layers[i], cache[i]
    g = gy * (C['h'] > 0) if C['act'] else gy
    gr = {}
    if C['use_bn']:
      B = g.shape[0]
      gr['gamma'] = (g * C['xhat']).sum(axis=0, dtype=np.float32)
      gr['beta'] = g.sum(axis=0, dtype=np.float32)
      gxhat = g * L['gamma']
      gz = (C['rstd'] / B) * (B * gxhat - gxhat.sum(axis=0) - C['xhat'] *
                               (gxhat * C['xhat']).sum(axis=0))
      gz = gz.astype(np.float32)
    else:
      gz = g
    gr['W'] = (C['x'].T @ gz).astype(np.float32)
    gr['b'] = gz.sum(axis=0, dtype=np.float32)
    gy = (gz @ L['W'].T).astype(np.float32)
    grads[i] = gr
  return gy, grads


def deepfm_forward(wide, deep, n_field, dim, params, training=True):
  """model/deepfm.py:53-109 with final_dnn.

  wide [B, F] (wide_output_dim=1 columns), deep [B, F*D].
  params: dict(dnn=[...], final=[...], out_W [H,1], out_b [1])."""
  wide_fea = wide.sum(axis=1, keepdims=True, dtype=np.float32)
  fm = fm_fwd(deep, n_field, dim)
  deep_fea, c1 = dnn_forward(deep, params['dnn'], training)
  all_fea = np.concatenate([wide_fea, fm, deep_fea], axis=1).astype(np.float32)
  fin, c2 = dnn_forward(all_fea, params['final'], training)
  logits = (fin @ params['out_W'] + params['out_b']).astype(np.float32)[:, 0]
  return logits, dict(c1=c1, c2=c2, fin=fin, all_fea=all_fea, deep=deep)


def deepfm_backward(g_logits, wide, deep, n_field, dim, params, cache):
  g = g_logits[:, None].astype(np.float32)
  grads = {'out_W': (cache['fin'].T @ g).astype(np.float32), 'out_b': g.sum(axis=0)}
  g_fin = (g @ params['out_W'].T).astype(np.float32)
  g_all, grads['final'] = dnn_backward(g_fin, params['final'], cache['c2'])
  g_wide = np.repeat(g_all[:, :1], wide.shape[1], axis=1).astype(np.float32)
  g_fm = g_all[:, 1:1 + dim]
  g_deep_mlp, grads['dnn'] = dnn_backward(np.ascontiguousarray(g_all[:, 1 + dim:]),
                                          params['dnn'], cache['c1'])
  g_deep = (g_deep_mlp + fm_bwd(deep, np.ascontiguousarray(g_fm), n_field, dim)).astype(np.float32)
  return g_wide, g_deep, grads


# ---- interaction formulas of the other model families (numpy fp32) ---------------------------------


def din_attention(query, keys, lens, mlp_layers):
  """layers/sequence_feature_layer.py:123-189 (= model/multi_tower_din.py:62-97): target attention.

  query [B, D], keys [B, T, D], lens [B]; mlp_layers as for dnn_forward (last layer linear, no BN).
  concat[q, k, q-k, q*k] -> DNN -> scores [B, T]; positions >= len get -2**32 + 1; softmax over T;
  returns the attended history [B, D] (the reference then concatenates the key)."""
  query, keys = query.astype(np.float32), keys.astype(np.float32)
  B, T, D = keys.shape
  q = np.broadcast_to(query[:, None, :], (B, T, D))
  din = np.concatenate([q, keys, q - keys, q * keys], axis=-1).reshape(B * T, 4 * D)
  scores, _ = dnn_forward(din, mlp_layers, training=True, last_no_act=True, last_no_bn=True)
  scores = scores.reshape(B, T)
  mask = np.arange(T)[None, :] < np.asarray(lens).reshape(B, 1)
  scores = np.where(mask, scores, np.float32(-2.0**32 + 1)).astype(np.float32)
  e = np.exp(scores - scores.max(axis=1, keepdims=True), dtype=np.float32)
  p = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
  return np.einsum('bt,btd->bd', p, keys).astype(np.float32)


def cross_v1(x0, weights, biases):
  """model/dcn.py:32-45: x <- x0 * (x . w) + b + x, one (w, b) pair per cross layer."""
  x0 = x0.astype(np.float32)
  x = x0
  for w, b in zip(weights, biases):
    xw = (x * np.asarray(w, np.float32)).sum(axis=1, keepdims=True, dtype=np.float32)
    x = (x0 * xw + np.asarray(b, np.float32) + x).astype(np.float32)
  return x


def cross_v2(x0, x, kernel, bias, diag_scale=0.0, u=None):
  """layers/keras/interaction.py:249-286: x0 * (W x + b [+ diag_scale * x]) + x; W = u @ kernel when the
  low-rank projection `u` [d, p] is given (then kernel is V [p, d])."""
  x0, x = x0.astype(np.float32), x.astype(np.float32)
  h = x if u is None else (x @ np.asarray(u, np.float32)).astype(np.float32)
  prod = (h @ np.asarray(kernel, np.float32) + np.asarray(bias, np.float32)).astype(np.float32)
  if diag_scale:
    prod = prod + np.float32(diag_scale) * x
  return (x0 * prod + x).astype(np.float32)


def dot_interaction(x, self_interaction=False, skip_gather=False):
  """layers/keras/interaction.py:47-128: x [B, F, D] -> pairwise dot products of the lower triangle in
  row-major order (with the diagonal when self_interaction); skip_gather keeps [B, F*F] with the rest zeroed."""
  x = x.astype(np.float32)
  B, F, _ = x.shape
  xa = np.einsum('bfd,bgd->bfg', x, x).astype(np.float32)
  keep = np.tril(np.ones((F, F), bool), 0 if self_interaction else -1)
  if skip_gather:
    return (xa * keep).reshape(B, F * F).astype(np.float32)
  return xa[:, keep]


def mmoe(x, experts, gates):
  """layers/mmoe.py:55-83: experts = list of DNN layer lists (relu on every layer), gates = list of (W, b);
  task output = sum_e softmax(x W + b)[e] * expert_e(x).  Returns one [B, H] array per task."""
  x = x.astype(np.float32)
  ex = np.stack([dnn_forward(x, layers)[0] for layers in experts], axis=1)   # [B, E, H]
  outs = []
  for w, b in gates:
    logit = (x @ np.asarray(w, np.float32) + np.asarray(b, np.float32)).astype(np.float32)
    e = np.exp(logit - logit.max(axis=1, keepdims=True), dtype=np.float32)
    g = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    outs.append((ex * g[:, :, None]).sum(axis=1, dtype=np.float32))
  return outs


def l2_normalize(x, eps=1e-12):
  """tf.nn.l2_normalize over the last axis (model/dssm.py:64-66 via match_model.norm)."""
  x = x.astype(np.float32)
  return (x / np.sqrt(np.maximum((x * x).sum(axis=-1, keepdims=True, dtype=np.float32), np.float32(eps)))).astype(np.float32)


def inbatch_softmax_ce(sim, item_ids=None, weights=None):
  """model/match_model.py:50-69 + model/dssm.py:90-93 + match_model.py:213-226: in-batch duplicates of the
  positive item are pushed to -1e32, softmax over the row, loss = -mean(log(p_bb + 1e-12) * w) / mean(w).
  sim [B, N >= B].  Returns (loss, probs [B, N])."""
  sim = sim.astype(np.float32).copy()
  B = sim.shape[0]
  if item_ids is not None:
    ids = np.asarray(item_ids)[:B]
    dup = (ids[None, :] == ids[:, None]).astype(np.float32) - np.eye(B, dtype=np.float32)
    sim[:, :B] = sim[:, :B] - dup * np.float32(1e32)
  e = np.exp(sim - sim.max(axis=1, keepdims=True), dtype=np.float32)
  probs = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
  w = np.ones(B, np.float32) if weights is None else np.asarray(weights, np.float32)
  hit = probs[np.arange(B), np.arange(B)]
  loss = -np.mean(np.log(hit + np.float32(1e-12)) * w, dtype=np.float32) / np.mean(w, dtype=np.float32)
  return np.float32(loss), probs


# ---- stateless activations of get_activation (utils/activation.py:46-118) --------------------------------------------
_SELU_SCALE, _SELU_ALPHA = 1.0507009873554804934193349852946, 1.6732632423543772848170429916717


def activation(x, name):
  """float64 evaluation of `name`(x) as utils/activation.py:66-118 resolves it (gelu is the reference's own tanh form,
  :46-60; leaky_relu / prelu-without-arguments = tf.nn.leaky_relu, alpha 0.2; swish = x * sigmoid(x), :63-65)."""
  x = np.asarray(x, np.float64)
  if name == 'gelu':
    return x * 0.5 * (1.0 + np.tanh(np.sqrt(2 / np.pi) * (x + 0.044715 * x ** 3)))
  if name in ('leaky_relu', 'prelu'):
    return np.maximum(0.2 * x, x)
  if name == 'elu':
    return np.where(x < 0, np.expm1(x), x)
  if name == 'selu':
    return np.where(x < 0, _SELU_SCALE * _SELU_ALPHA * np.expm1(x), _SELU_SCALE * x)
  if name == 'tanh':
    return np.tanh(x)
  if name == 'swish':
    return x / (1.0 + np.exp(-x))
  if name == 'sigmoid':
    return 1.0 / (1.0 + np.exp(-x))
  if name == 'relu':
    return np.maximum(x, 0.0)
  raise ValueError(name)


def activation_grad(x, name):
  """d activation / dx in float64 (TF's gradient kernels take the negative branch of elu / selu for x < 0 and of
  leaky_relu for x <= 0)."""
  x = np.asarray(x, np.float64)
  if name == 'gelu':
    c = np.sqrt(2 / np.pi)
    t = np.tanh(c * (x + 0.044715 * x ** 3))
    return 0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * c * (1.0 + 3 * 0.044715 * x * x)
  if name in ('leaky_relu', 'prelu'):
    return np.where(x > 0, 1.0, 0.2)
  if name == 'elu':
    return np.where(x < 0, np.exp(x), 1.0)
  if name == 'selu':
    return np.where(x < 0, _SELU_SCALE * _SELU_ALPHA * np.exp(x), _SELU_SCALE)
  if name == 'tanh':
    return 1.0 - np.tanh(x) ** 2
  s = 1.0 / (1.0 + np.exp(-x))
  if name == 'swish':
    return s * (1.0 + x * (1.0 - s))
  if name == 'sigmoid':
    return s * (1.0 - s)
  if name == 'relu':
    return (x > 0).astype(np.float64)
  raise ValueError(name)


# ---- tf.metrics.auc / max_f1 (model/rank_model.py:360-373, core/metrics.py:25-56) -------------------------------------
def tf_thresholds(num_thresholds=200):
  """the threshold list of tf.metrics.auc and of core/metrics.py:33-38, as the float32 constants the graph compares
  float32 predictions with."""
  kepsilon = 1e-7
  t = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
  return np.array([0.0 - kepsilon] + t + [1.0 + kepsilon], np.float32)


def confusion_at_thresholds(labels, probs, num_thresholds=200):
  """tp, fn, tn, fp [T] by the definition: label & (pred > thr[i]) counted sample by sample (TF tiles the predictions
  against the thresholds, metrics_impl._confusion_matrix_at_thresholds); labels go through tf.to_int64 then bool."""
  thr = tf_thresholds(num_thresholds)
  lab = np.asarray(labels).reshape(-1).astype(np.int64) != 0
  p = np.asarray(probs, np.float32).reshape(-1)
  above = p[None, :] > thr[:, None]
  tp = (above & lab[None, :]).sum(1)
  fp = (above & ~lab[None, :]).sum(1)
  return tp, lab.sum() - tp, (~lab).sum() - fp, fp


def auc_tf(labels, probs, num_thresholds=200):
  """tf.metrics.auc(curve='ROC', summation_method='trapezoidal'): float32 accumulators, rec = (tp + 1e-6) /
  (tp + fn + 1e-6), fp_rate = fp / (fp + tn + 1e-6), sum((x[:-1] - x[1:]) * (y[:-1] + y[1:]) / 2)."""
  tp, fn, tn, fp = [a.astype(np.float32) for a in confusion_at_thresholds(labels, probs, num_thresholds)]
  eps = np.float32(1e-6)
  y = (tp + eps) / (tp + fn + eps)
  x = fp / (fp + tn + eps)
  return float(np.sum((x[:-1] - x[1:]) * ((y[:-1] + y[1:]) / np.float32(2.0)), dtype=np.float32))


def max_f1(labels, probs):
  """core/metrics.py:25-56: max over the 200 thresholds of 2 p r / (p + r + 1e-12), p = tp / (tp + fp), r = tp / (tp + fn)
  (tf.metrics.precision / recall: 0 when the denominator is 0)."""
  tp, fn, tn, fp = [a.astype(np.float32) for a in confusion_at_thresholds(labels, probs, 200)]
  with np.errstate(divide='ignore', invalid='ignore'):
    prec = np.where(tp + fp > 0, tp / (tp + fp), np.float32(0))
    rec = np.where(tp + fn > 0, tp / (tp + fn), np.float32(0))
  return float(np.max(2 * prec * rec / (prec + rec + np.float32(1e-12))))


def dice(x, alphas, eps=1e-9):
  """utils/activation.py:13-43 (training): p = sigmoid(batch_norm(x) without centre / scale, epsilon 1e-9, batch
  statistics with the biased variance); alphas * (1 - p) * x + p * x.  Not built in the product (refused by the scope
  check); restated here against the reference function's own output (tests/golden/reference_activations.json)."""
  x = np.asarray(x, np.float32)
  mu = x.mean(0, dtype=np.float32)
  var = ((x - mu) ** 2).mean(0, dtype=np.float32)
  p = 1.0 / (1.0 + np.exp(-((x - mu) / np.sqrt(var + np.float32(eps)))))
  return (np.asarray(alphas, np.float32) * (1.0 - p) * x + p * x).astype(np.float32)
