"""Dump golden vectors of the TensorFlow-owned arithmetic on the hot path from a REAL TensorFlow.

The reference's path is a TF graph; the pieces EasyRec does not implement itself - StringToHashBucketFast / AsString,
safe_embedding_lookup_sparse, SparseApplyAdagrad, the sparse applies of tf.train.AdamOptimizer and of EasyRec's lazy
AdamOptimizerS, layers.dense + batch_normalization, losses.sigmoid_cross_entropy, the tf.nn activations, tf.metrics.auc - are restated by `oracle/` and marked
"parity unpinned vs TF" in DESIGN.md §4 where no frozen TF test value exists.  TensorFlow is not installable in the build
container (no network) nor on the GPU box, so this script is for any machine that HAS it (TF 1.15 or 2.x):

    python tools/dump_tf_golden.py tests/golden/tf_golden.npz          # needs: tensorflow, numpy
    python -m pytest tests/test_tf_golden_replay.py                    # replays the dump on the oracle (CPU) and,
                                                                       # with -m gpu, on the kernels

The inputs are generated here from fixed seeds and stored in the .npz next to TF's outputs, so the replay needs neither
TF nor this script.  If `easy_rec` itself is importable (baseline/_ref on PYTHONPATH) the lazy-Adam rule is taken from
easy_rec.python.compat.adam_s.AdamOptimizerS; otherwise that block is skipped and recorded as absent.
"""
import sys

import numpy as np


def main(out_path):
  import tensorflow as tf
  tf1 = tf.compat.v1
  tf1.disable_eager_execution()
  rng = np.random.default_rng(20240)
  out = {'tf_version': np.array(tf.__version__)}

  # ---- A.1 raw value -> table row: as_string + string_to_hash_bucket_fast (feature_column_v2.py:3903-3930) -------
  ids = np.concatenate([rng.integers(0, 2**40, 4096), rng.integers(-2**62, 2**62, 1024),
                        np.array([0, 1, 9, 10, 99, 10**15, 10**16 - 1, 10**17, 2**63 - 1, -1, -2**63])]).astype(np.int64)
  strs = np.array(['', 'a', 'ab', 'abcdefg', 'abcdefgh', '0123456789abcdef', '0123456789abcdefg', 'x' * 32, 'y' * 33,
                   'z' * 64, 'w' * 65, 'q' * 200], dtype=object)
  buckets = np.array([10, 1000, 10_000_000, 2**31 - 1, 2**63 - 1], np.int64)
  with tf1.Session(graph=tf.Graph()) as sess:
    t_ids = tf1.placeholder(tf.int64, [None])
    t_str = tf1.placeholder(tf.string, [None])
    for b in buckets:
      out['hash_int_%d' % b] = sess.run(tf.strings.to_hash_bucket_fast(tf.strings.as_string(t_ids), int(b)), {t_ids: ids})
      out['hash_str_%d' % b] = sess.run(tf.strings.to_hash_bucket_fast(t_str, int(b)), {t_str: strs})
  out['hash_ids'], out['hash_strs'], out['hash_buckets'] = ids, np.array([s.encode() for s in strs]), buckets

  # ---- A.2 lookup + pooling: safe_embedding_lookup_sparse, every combiner, weights, empty rows, invalid ids -------
  V, D, B = 1000, 16, 64
  table = rng.standard_normal((V, D)).astype(np.float32)
  lens = rng.integers(0, 6, B)
  lens[:3] = 0
  n = int(lens.sum())
  sp_ids = rng.integers(0, V, n).astype(np.int64)
  sp_ids[rng.uniform(size=n) < 0.05] = -1           # pruned by safe_embedding_lookup_sparse
  sp_w = rng.uniform(-0.5, 2.0, n).astype(np.float32)  # non-positive weights are pruned under mean / sqrtn
  rows = np.repeat(np.arange(B), lens)
  cols = np.concatenate([np.arange(k) for k in lens]) if n else np.zeros(0, np.int64)
  indices = np.stack([rows, cols], 1).astype(np.int64)
  with tf1.Session(graph=tf.Graph()) as sess:
    t_tab = tf.constant(table)
    sid = tf.SparseTensor(indices, sp_ids, [B, 8])
    sw = tf.SparseTensor(indices, sp_w, [B, 8])
    for comb in ('sum', 'mean', 'sqrtn'):
      out['lookup_%s' % comb] = sess.run(tf.nn.safe_embedding_lookup_sparse(t_tab, sid, None, combiner=comb))
      out['lookup_%s_weighted' % comb] = sess.run(tf.nn.safe_embedding_lookup_sparse(t_tab, sid, sw, combiner=comb))
  out['lookup_table'], out['lookup_lens'], out['lookup_ids'], out['lookup_weights'] = table, lens, sp_ids, sp_w

  # ---- A.4 sparse optimizer rules on duplicated rows (IndexedSlices), three steps ---------------------------------
  idx = rng.integers(0, 50, 200).astype(np.int64)        # many duplicates
  grads = [rng.standard_normal((200, D)).astype(np.float32) * 0.1 for _ in range(3)]
  w0 = rng.standard_normal((50, D)).astype(np.float32) * 0.01

  def run_opt(make_opt, tag):
    with tf1.Session(graph=tf.Graph()) as sess:
      var = tf1.get_variable('w_' + tag, initializer=tf.constant(w0), use_resource=False)
      opt = make_opt()
      g_ph = tf1.placeholder(tf.float32, [200, D])
      step = opt.apply_gradients([(tf.IndexedSlices(g_ph, tf.constant(idx), tf.constant([50, D], tf.int64)), var)])
      sess.run(tf1.global_variables_initializer())
      for k, g in enumerate(grads):
        sess.run(step, {g_ph: g})
        out['%s_step%d' % (tag, k)] = sess.run(var)

  run_opt(lambda: tf1.train.AdagradOptimizer(0.05, initial_accumulator_value=0.1), 'adagrad')
  run_opt(lambda: tf1.train.AdamOptimizer(0.01, beta1=0.9, beta2=0.999, epsilon=1e-8), 'adam')
  try:
    from easy_rec.python.compat.adam_s import AdamOptimizerS
    run_opt(lambda: AdamOptimizerS(0.01, beta1=0.9, beta2=0.999, epsilon=1e-8), 'lazy_adam')
  except Exception as e:   # easy_rec not importable here
    out['lazy_adam_absent'] = np.array(repr(e))
  out['opt_idx'], out['opt_w0'] = idx, w0
  for k, g in enumerate(grads):
    out['opt_grad%d' % k] = g

  # ---- layers.dense + batch_normalization (training) + relu, and losses.sigmoid_cross_entropy ---------------------
  x = rng.standard_normal((256, 24)).astype(np.float32)
  labels = (rng.uniform(size=256) < 0.3).astype(np.float32)
  with tf1.Session(graph=tf.Graph()) as sess:
    tx = tf.constant(x)
    h = tf1.layers.dense(tx, 16, name='d0')
    hb = tf1.layers.batch_normalization(h, training=True, name='d0/bn')
    y = tf.nn.relu(hb)
    logit = tf1.layers.dense(y, 1, name='out')[:, 0]
    loss = tf1.losses.sigmoid_cross_entropy(tf.constant(labels), logit)
    params = {v.name: v for v in tf1.global_variables()}
    gs = tf.gradients(loss, [params['d0/kernel:0'], params['d0/bn/gamma:0'], params['d0/bn/beta:0'], params['out/kernel:0']])
    sess.run(tf1.global_variables_initializer())
    vals = sess.run({'h': h, 'bn': hb, 'logit': logit, 'loss': loss, 'g_d0_kernel': gs[0], 'g_gamma': gs[1],
                     'g_beta': gs[2], 'g_out_kernel': gs[3]})
    for k, v in vals.items():
      out['dense_' + k] = v
    for k, v in params.items():
      out['dense_param_' + k.replace('/', '.').replace(':0', '')] = sess.run(v)
    sess.run(tf1.get_collection(tf1.GraphKeys.UPDATE_OPS))
    out['dense_moving_mean'] = sess.run(params['d0/bn/moving_mean:0'])
    out['dense_moving_var'] = sess.run(params['d0/bn/moving_variance:0'])
  out['dense_x'], out['dense_labels'] = x, labels

  # ---- activations (tf.nn.* as utils/activation.py:get_activation resolves them) with gradients ----------------------
  ax = np.concatenate([rng.normal(0, 2.5, 2048), [0.0, -0.0, 1e-6, -1e-6, 20.0, -20.0]]).astype(np.float32)
  acts = {'leaky_relu': tf.nn.leaky_relu, 'elu': tf.nn.elu, 'selu': tf.nn.selu, 'tanh': tf.tanh, 'sigmoid': tf.nn.sigmoid,
          'swish': tf.nn.swish, 'gelu': lambda t: t * 0.5 * (1.0 + tf.tanh(np.sqrt(2 / np.pi) * (t + 0.044715 * tf.pow(t, 3))))}
  with tf1.Session(graph=tf.Graph()) as sess:
    t = tf1.placeholder(tf.float32, [None])
    for name, fn in acts.items():
      y = fn(t)
      out['act_%s' % name], out['act_%s_grad' % name] = sess.run([y, tf.gradients(tf.reduce_sum(y), t)[0]], {t: ax})
  out['act_x'] = ax

  # ---- tf.metrics.auc at several num_thresholds, streamed over batches (model/rank_model.py:360-373) -----------------
  mlab = (rng.uniform(size=6000) < 0.3).astype(np.float32)
  mpred = np.clip(rng.normal(0.4 + 0.25 * mlab, 0.2), 0, 1).astype(np.float32)
  for T in (200, 500):
    with tf1.Session(graph=tf.Graph()) as sess:
      tl, tp = tf1.placeholder(tf.float32, [None]), tf1.placeholder(tf.float32, [None])
      val, upd = tf1.metrics.auc(tf.cast(tl, tf.int64), tp, num_thresholds=T)
      sess.run(tf1.local_variables_initializer())
      for k in range(0, 6000, 1000):
        sess.run(upd, {tl: mlab[k:k + 1000], tp: mpred[k:k + 1000]})
      out['auc_%d' % T] = sess.run(val)
  out['auc_labels'], out['auc_preds'] = mlab, mpred

  np.savez_compressed(out_path, **out)
  print('wrote %s (%d arrays) from TensorFlow %s' % (out_path, len(out), tf.__version__))


if __name__ == '__main__':
  main(sys.argv[1] if len(sys.argv) > 1 else 'tests/golden/tf_golden.npz')
