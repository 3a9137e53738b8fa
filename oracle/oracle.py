"""TEST INFRASTRUCTURE -- numpy/ctypes front end of the CPU oracle (oracle/er_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; the product package (easyrec_b200/) never does.

The sparse path is the C restatement; the dense model pieces (DNN + batch-norm,
DeepFM head, DIN attention, DCN cross, MMoE, DSSM, losses) are restated in numpy
fp32 below, each citing the reference lines it follows.  Parity status: see the
header of er_oracle.c and DESIGN.md -- everything except Fingerprint64's short-string
branches and the two embed_test.py known answers is "parity unpinned" (TensorFlow
cannot be imported in this container).
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
