"""Replay of a REAL TensorFlow dump (tools/dump_tf_golden.py -> tests/golden/tf_golden.npz) on the oracle.

TensorFlow cannot be installed in the build container or on the GPU box, so the dump is absent by default and the
replay skips; on a machine with TF, run the tool once, commit the .npz, and the pieces DESIGN.md §4 lists as "parity
unpinned vs TF" (hash of long strings, sparse Adagrad / Adam applies on duplicated rows, batch norm, sigmoid CE) become
pinned.  `test_replay_plumbing_on_a_self_made_dump` keeps the replay code itself exercised: it builds a dump of the
same schema from the oracle (so it proves nothing about TF, only that the replay reads the schema it claims)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

PATH = os.path.join(os.path.dirname(__file__), 'golden', 'tf_golden.npz')
COMB = {'sum': 0, 'mean': 1, 'sqrtn': 2}


def _pooled(table, lens, ids, w, comb):
  row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
  return O.embedding_fwd(table, ids, row_ptr, COMB[comb], weights=w)[0]


def _opt_steps(g, kind, n_steps=3):
  idx, w = g['opt_idx'], g['opt_w0'].copy()
  V, D = w.shape
  s0 = np.full((V, D), 0.1, np.float32) if kind == 'adagrad' else np.zeros((V, D), np.float32)
  s1 = np.zeros((V, D), np.float32)
  seg = np.arange(idx.size, dtype=np.int32)
  b1p, b2p = np.float32(1.0), np.float32(1.0)
  outs = []
  for k in range(n_steps):
    gr = g['opt_grad%d' % k]
    if kind == 'adagrad':
      O.embedding_bwd(w, s0, None, idx, seg, gr, O.OPT_ADAGRAD, 0.05)
    else:
      b1p, b2p = np.float32(b1p * np.float32(0.9)), np.float32(b2p * np.float32(0.999))
      if kind == 'lazy_adam':
        O.embedding_bwd(w, s0, s1, idx, seg, gr, O.OPT_LAZY_ADAM, 0.01, beta1_power=b1p, beta2_power=b2p)
      else:
        O.embedding_bwd_adam_dense(w, s0, s1, idx, seg, gr, 0.01, beta1_power=b1p, beta2_power=b2p)
    outs.append(w.copy())
  return outs


def _dense(g):
  L = [dict(W=g['dense_param_d0.kernel'], b=g['dense_param_d0.bias'], gamma=g['dense_param_d0.bn.gamma'],
            beta=g['dense_param_d0.bn.beta'])]
  y, cache = O.dnn_forward(g['dense_x'], L)
  logit = (y @ g['dense_param_out.kernel'] + g['dense_param_out.bias'])[:, 0]
  loss, _, g_logit = O.sigmoid_ce(logit, g['dense_labels'])
  gy = g_logit[:, None] @ g['dense_param_out.kernel'].T
  _, grads = O.dnn_backward(gy.astype(np.float32), L, cache)
  return dict(h=cache[0]['z'], bn=cache[0]['h'], logit=logit, loss=loss, g_d0_kernel=grads[0]['W'],
              g_gamma=grads[0]['gamma'], g_beta=grads[0]['beta'], g_out_kernel=y.T @ g_logit[:, None],
              mean=cache[0]['mu'], var=cache[0]['var'])


def _replay(g):
  # integer, bit exact
  for b in g['hash_buckets']:
    assert np.array_equal(O.bucketize(g['hash_ids'], 0, int(b), 0)[0], g['hash_int_%d' % b])
    got = np.array([O.fingerprint64(bytes(s)) % int(b) for s in g['hash_strs']], np.uint64).astype(np.int64)
    assert np.array_equal(got, g['hash_str_%d' % b])
  # lookup + pooling
  for comb in COMB:
    np.testing.assert_allclose(_pooled(g['lookup_table'], g['lookup_lens'], g['lookup_ids'], None, comb),
                               g['lookup_%s' % comb], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(_pooled(g['lookup_table'], g['lookup_lens'], g['lookup_ids'], g['lookup_weights'], comb),
                               g['lookup_%s_weighted' % comb], rtol=1e-5, atol=1e-6)
  # sparse applies on duplicated rows: post-step rows within 1e-6
  for kind in ('adagrad', 'adam') + (() if 'lazy_adam_absent' in g else ('lazy_adam',)):
    for k, w in enumerate(_opt_steps(g, kind)):
      np.testing.assert_allclose(w, g['%s_step%d' % (kind, k)], rtol=0, atol=1e-6, err_msg='%s step %d' % (kind, k))
  # dense + batch norm + relu + sigmoid CE, forward and gradients
  d = _dense(g)
  for k in ('h', 'bn', 'logit', 'g_d0_kernel', 'g_gamma', 'g_beta', 'g_out_kernel'):
    np.testing.assert_allclose(d[k], g['dense_' + k], rtol=1e-4, atol=1e-5, err_msg=k)
  assert abs(float(d['loss']) - float(g['dense_loss'])) < 1e-5
  np.testing.assert_allclose(d['mean'] * 0.01, g['dense_moving_mean'], rtol=1e-4, atol=1e-6)
  np.testing.assert_allclose(1.0 * 0.99 + d['var'] * 0.01, g['dense_moving_var'], rtol=1e-4, atol=1e-6)


  # activations and their gradients (dumps made before these were added do not carry them)
  if 'act_x' in g:
    for name in ('leaky_relu', 'elu', 'selu', 'tanh', 'sigmoid', 'swish', 'gelu'):
      np.testing.assert_allclose(O.activation(g['act_x'], name), g['act_%s' % name], rtol=2e-6, atol=1e-6, err_msg=name)
      np.testing.assert_allclose(O.activation_grad(g['act_x'], name), g['act_%s_grad' % name], rtol=1e-5, atol=2e-6,
                                 err_msg=name + ' gradient')
  # tf.metrics.auc streamed over batches
  if 'auc_labels' in g:
    for T in (200, 500):
      assert abs(O.auc_tf(g['auc_labels'], g['auc_preds'], T) - float(g['auc_%d' % T])) < 2e-6


@pytest.mark.skipif(not os.path.exists(PATH), reason='no TensorFlow dump (tools/dump_tf_golden.py needs a box with TF)')
def test_oracle_matches_the_tensorflow_dump():
  _replay(dict(np.load(PATH, allow_pickle=False)))


def test_replay_plumbing_on_a_self_made_dump():
  rng = np.random.default_rng(1)
  g = {}
  ids = rng.integers(0, 2**40, 64).astype(np.int64)
  strs = np.array([b'', b'abc', b'0123456789abcdefg'])
  g['hash_ids'], g['hash_strs'], g['hash_buckets'] = ids, strs, np.array([10, 1000], np.int64)
  for b in (10, 1000):
    g['hash_int_%d' % b] = O.bucketize(ids, 0, b, 0)[0]
    g['hash_str_%d' % b] = np.array([O.fingerprint64(bytes(s)) % b for s in strs], np.int64)
  V, D, B = 50, 4, 8
  g['lookup_table'] = rng.standard_normal((V, D)).astype(np.float32)
  g['lookup_lens'] = np.array([0, 2, 1, 3, 0, 1, 1, 2])
  n = int(g['lookup_lens'].sum())
  g['lookup_ids'] = rng.integers(0, V, n).astype(np.int64)
  g['lookup_weights'] = rng.uniform(0.1, 2.0, n).astype(np.float32)
  for comb in COMB:
    g['lookup_%s' % comb] = _pooled(g['lookup_table'], g['lookup_lens'], g['lookup_ids'], None, comb)
    g['lookup_%s_weighted' % comb] = _pooled(g['lookup_table'], g['lookup_lens'], g['lookup_ids'], g['lookup_weights'], comb)
  g['opt_idx'] = rng.integers(0, 10, 40).astype(np.int64)
  g['opt_w0'] = rng.standard_normal((10, D)).astype(np.float32)
  for k in range(3):
    g['opt_grad%d' % k] = rng.standard_normal((40, D)).astype(np.float32)
  for kind in ('adagrad', 'adam', 'lazy_adam'):
    for k, w in enumerate(_opt_steps(g, kind)):
      g['%s_step%d' % (kind, k)] = w
  g['dense_x'] = rng.standard_normal((32, 6)).astype(np.float32)
  g['dense_labels'] = (rng.uniform(size=32) < 0.3).astype(np.float32)
  g['dense_param_d0.kernel'] = rng.standard_normal((6, 5)).astype(np.float32)
  g['dense_param_d0.bias'] = np.zeros(5, np.float32)
  g['dense_param_d0.bn.gamma'] = np.ones(5, np.float32)
  g['dense_param_d0.bn.beta'] = np.zeros(5, np.float32)
  g['dense_param_out.kernel'] = rng.standard_normal((5, 1)).astype(np.float32)
  g['dense_param_out.bias'] = np.zeros(1, np.float32)
  d = _dense(g)
  for k in ('h', 'bn', 'logit', 'loss', 'g_d0_kernel', 'g_gamma', 'g_beta', 'g_out_kernel'):
    g['dense_' + k] = d[k]
  g['dense_moving_mean'] = d['mean'] * 0.01
  g['dense_moving_var'] = 0.99 + d['var'] * 0.01
  g['act_x'] = rng.normal(0, 2, 64).astype(np.float32)
  for name in ('leaky_relu', 'elu', 'selu', 'tanh', 'sigmoid', 'swish', 'gelu'):
    g['act_%s' % name], g['act_%s_grad' % name] = O.activation(g['act_x'], name), O.activation_grad(g['act_x'], name)
  g['auc_labels'] = (rng.uniform(size=500) < 0.3).astype(np.float32)
  g['auc_preds'] = rng.uniform(size=500).astype(np.float32)
  for T in (200, 500):
    g['auc_%d' % T] = O.auc_tf(g['auc_labels'], g['auc_preds'], T)
  _replay(g)
