"""Golden vectors for the interaction formulas, produced by EXECUTING THE REFERENCE'S OWN CODE.

TensorFlow is not installable here, but these functions only use a handful of tensor ops whose semantics
are unambiguous (stack, square, reduce_sum, subtract, add, matmul, band_part, boolean_mask, where, ...).
The reference source is loaded from /root/reference (never copied), the function bodies are taken as they
are (ast), and they run against `_tf`, a numpy implementation of exactly those ops, in float32:

  layers/fm.py:20-26                         FM.__call__
  model/dcn.py:32-45                         DCN._cross_net      (tf.get_variable -> supplied w / b)
  layers/keras/interaction.py:24-44          keras FM.call
  layers/keras/interaction.py:47-128         DotInteraction.call
  layers/keras/interaction.py:249-286        Cross.call          (Dense -> supplied kernel / bias)
  layers/sequence_feature_layer.py:123-189   SequenceFeatureLayer.target_attention (dnn.DNN -> supplied MLP, no BN)
  model/multi_tower_din.py:62-97             MultiTowerDIN.din
  layers/mmoe.py:55-83                       MMOE.gate / MMOE.__call__
  model/match_model.py:50-69,71-126,213-234  MatchModel._mask_in_batch / _list_wise_sim / _build_list_wise_loss_graph
  core/learning_schedules.py:30-75           exponential_decay_with_burnin
  compat/adam_s.py:185-213                   AdamOptimizerS._apply_sparse_shared (lazy Adam row rule)
  model/deepfm.py:53-109                     DeepFM.build_predict_graph (wide sum | FM | deep DNN -> final DNN -> logit)

Run in the build container (reference mounted):  python tests/golden/make_formula_golden.py
-> tests/golden/reference_formulas.json (inputs + outputs, small shapes), replayed by
tests/test_oracle_golden.py (CPU oracle) and tests/test_gpu_interactions.py (CUDA kernels)."""
import ast
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference/easy_rec/python'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_formulas.json')


class _Scope(object):
  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False


def _make_tf(variables):
  tf = types.SimpleNamespace()
  tf.__version__ = '1.15.5'
  tf.float32 = np.float32
  tf.bool = np.bool_
  tf.name_scope = lambda *a, **k: _Scope()
  tf.stack = lambda xs, axis=0: np.stack([np.asarray(x, np.float32) for x in xs], axis=axis)
  tf.concat = lambda xs, axis: np.concatenate(xs, axis=axis)
  tf.square = lambda x: np.square(x, dtype=np.float32)
  tf.reduce_sum = lambda x, axis=None, keepdims=False, name=None: np.sum(x, axis=axis, keepdims=keepdims, dtype=np.float32)
  tf.subtract = lambda a, b: (a - b).astype(np.float32)
  tf.matmul = lambda a, b, transpose_b=False: np.matmul(a, np.swapaxes(b, -1, -2) if transpose_b else b).astype(np.float32)
  tf.shape = lambda x: x.shape
  tf.ones_like = np.ones_like
  tf.zeros_like = np.zeros_like
  tf.cast = lambda x, dt: x.astype(dt)
  tf.where = lambda condition, x, y: np.where(condition, x, y)
  tf.reshape = lambda x, shape: np.reshape(x, shape)
  tf.boolean_mask = lambda x, mask: np.stack([xi[mi.astype(bool)] for xi, mi in zip(x, mask)])

  def band_part(x, lo, hi):
    n = x.shape[-1]
    i, j = np.arange(n)[:, None], np.arange(n)[None, :]
    keep = ((lo < 0) | (i - j <= lo)) & ((hi < 0) | (j - i <= hi))
    return x * keep.astype(x.dtype)
  tf.linalg = types.SimpleNamespace(band_part=band_part)
  tf.math = types.SimpleNamespace(add=lambda a, b: (a + b).astype(np.float32))
  tf.get_variable = lambda name, dtype=None, shape=None: variables[name]
  tf.errors = types.SimpleNamespace(InvalidArgumentError=ValueError)
  # ---- ops used by the DIN / MMoE / match-model / schedule functions ----
  f32 = np.float32
  tf.int32, tf.int64 = np.int32, np.int64
  tf.tile = lambda x, reps: np.tile(x, reps)
  tf.expand_dims = lambda x, axis: np.expand_dims(x, axis)
  tf.multiply = lambda a, b: (a * b).astype(f32)
  tf.identity = lambda x, name=None: x
  tf.transpose = lambda x: np.transpose(x)
  tf.range = lambda n: np.arange(n)
  tf.squeeze = lambda x: np.squeeze(x)
  tf.log = lambda x: np.log(x, dtype=f32)
  tf.reduce_mean = lambda x, axis=None: np.mean(x, axis=axis, dtype=f32)
  tf.ones = lambda shape, dtype=f32: np.ones(shape, dtype)
  tf.diag = lambda v: np.diag(v)
  tf.to_float = lambda x: x.astype(f32)
  tf.equal = lambda a, b: a == b
  tf.less = lambda a, b: a < b
  tf.constant = lambda v, dtype=None, name=None: v
  tf.maximum = lambda a, b, name=None: np.maximum(a, b, dtype=f32)
  tf.gather_nd = lambda x, idx: x[tuple(np.asarray(idx).T)]
  tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None)
  tf.estimator = types.SimpleNamespace(ModeKeys=types.SimpleNamespace(PREDICT='infer', TRAIN='train'))

  def sequence_mask(lengths, maxlen=None):
    lengths = np.asarray(lengths)
    n = int(lengths.max()) if maxlen is None else maxlen   # tf.sequence_mask default: max over the batch
    return np.arange(n) < lengths[..., None]
  tf.sequence_mask = sequence_mask

  def softmax(x, axis=-1):
    x = np.asarray(x, f32)
    e = np.exp(x - x.max(axis=axis, keepdims=True), dtype=f32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=f32)).astype(f32)
  tf.nn = types.SimpleNamespace(softmax=softmax, relu=lambda x: np.maximum(x, f32(0)))

  def dense(inputs, units, kernel_regularizer=None, name=None):   # tf.layers.dense, linear
    w, b = variables[name + '/kernel'], variables[name + '/bias']
    assert w.shape[1] == units
    return (inputs @ w + b).astype(f32)
  tf.layers = types.SimpleNamespace(dense=dense)

  def exponential_decay(lr, global_step, decay_steps, decay_rate, staircase=False):   # tf.train.exponential_decay
    p = f32(global_step) / f32(decay_steps)
    if staircase:
      p = np.floor(p)
    return f32(lr) * np.power(f32(decay_rate), p, dtype=f32)
  tf.train = types.SimpleNamespace(exponential_decay=exponential_decay)
  return tf


class _NumpyDNN(object):
  """stands in for layers/dnn.py DNN with use_bn=false, no dropout: dense(+bias) -> relu per layer; the last
  layer is linear when last_layer_no_activation (layers/dnn.py:50-87)."""

  def __init__(self, dnn_config, l2_reg, name='dnn', is_training=False, last_layer_no_activation=False,
               last_layer_no_batch_norm=False):
    self.layers = dnn_config[name] if isinstance(dnn_config, dict) else dnn_config
    self.last_linear = last_layer_no_activation

  def __call__(self, x):
    for i, (w, b) in enumerate(self.layers):
      x = (x @ w + b).astype(np.float32)
      if not (self.last_linear and i == len(self.layers) - 1):
        x = np.maximum(x, np.float32(0))
    return x


def _function(path, cls, name):
  """the reference's function object, compiled from its own source text (no package import, no TF)."""
  src = open(os.path.join(REF, path)).read()
  tree = ast.parse(src)
  for node in tree.body:
    if cls is None and isinstance(node, ast.FunctionDef) and node.name == name:
      return compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, path), 'exec'), node.lineno
    if isinstance(node, ast.ClassDef) and node.name == cls:
      for fn in node.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == name:
          mod = ast.Module(body=[fn], type_ignores=[])
          return compile(mod, os.path.join(REF, path), 'exec'), fn.lineno
  raise KeyError('%s.%s not found in %s' % (cls, name, path))


def run(path, cls, name, tf, self_obj, *args, **extra_globals):
  code, line = _function(path, cls, name)
  ns = {'tf': tf}
  ns.update(extra_globals)
  exec(code, ns)
  if cls is None:
    return ns[name](*args), line
  return ns[name](self_obj, *args), line


def _mlp(rng, dims):
  return [(rng.normal(0, 0.4, (a, b)).astype(np.float32), rng.normal(0, 0.1, b).astype(np.float32))
          for a, b in zip(dims[:-1], dims[1:])]


def _mlp_json(layers):
  return [{'w': w.tolist(), 'b': b.tolist()} for w, b in layers]


def main():
  rng = np.random.default_rng(20240)
  out = {'generator': 'tests/golden/make_formula_golden.py', 'cases': {}}
  # ---- layers/fm.py FM.__call__ ----
  feas = [rng.normal(size=(5, 8)).astype(np.float32) for _ in range(6)]
  y, line = run('layers/fm.py', 'FM', '__call__', _make_tf({}), types.SimpleNamespace(_name='fm'), feas)
  out['cases']['fm'] = {'ref': 'layers/fm.py:%d' % line, 'x': np.stack(feas, 1).tolist(), 'y': y.tolist()}
  # ---- model/dcn.py DCN._cross_net ----
  d, n_layers = 12, 3
  x = rng.normal(size=(7, d)).astype(np.float32)
  variables = {}
  for i in range(n_layers):
    variables['cross_layer_%d_w' % i] = rng.uniform(-0.5, 0.5, d).astype(np.float32)
    variables['cross_layer_%d_b' % i] = rng.uniform(-0.5, 0.5, d).astype(np.float32)
  y, line = run('model/dcn.py', 'DCN', '_cross_net', _make_tf(variables), None, x, n_layers)
  out['cases']['dcn_cross'] = {'ref': 'model/dcn.py:%d' % line, 'x': x.tolist(),
                               'w': [variables['cross_layer_%d_w' % i].tolist() for i in range(n_layers)],
                               'b': [variables['cross_layer_%d_b' % i].tolist() for i in range(n_layers)],
                               'y': np.asarray(y, np.float32).tolist()}
  # ---- keras DotInteraction.call ----
  feas = [rng.normal(size=(4, 6)).astype(np.float32) for _ in range(5)]
  for self_int in (False, True):
    me = types.SimpleNamespace(_self_interaction=self_int, _skip_gather=False)
    y, line = run('layers/keras/interaction.py', 'DotInteraction', 'call', _make_tf({}), me, feas)
    out['cases']['dot_interaction_self%d' % int(self_int)] = {
        'ref': 'layers/keras/interaction.py:%d' % line, 'x': np.stack(feas, 1).tolist(),
        'self_interaction': self_int, 'y': np.asarray(y, np.float32).tolist()}
  more_cases(rng, out)
  json.dump(out, open(OUT, 'w'))
  print('wrote', OUT, sorted(out['cases']))


def more_cases(rng, out):
  import logging
  f32 = np.float32
  # ---- keras Cross.call (v2): full-rank with diag_scale, and low-rank U V ----
  d = 10
  x0, x = rng.normal(size=(6, d)).astype(f32), rng.normal(size=(6, d)).astype(f32)
  w, b = rng.normal(0, 0.3, (d, d)).astype(f32), rng.normal(0, 0.1, d).astype(f32)
  me = types.SimpleNamespace(built=True, _projection_dim=None, _diag_scale=0.25,
                             _dense=lambda t: (t @ w + b).astype(f32))
  y, line = run('layers/keras/interaction.py', 'Cross', 'call', _make_tf({}), me, [x0, x])
  out['cases']['keras_cross_full'] = {'ref': 'layers/keras/interaction.py:%d' % line, 'x0': x0.tolist(), 'x': x.tolist(),
                                      'w': w.tolist(), 'b': b.tolist(), 'diag_scale': 0.25, 'y': y.astype(f32).tolist()}
  u, v = rng.normal(0, 0.3, (d, 3)).astype(f32), rng.normal(0, 0.3, (3, d)).astype(f32)
  me = types.SimpleNamespace(built=True, _projection_dim=3, _diag_scale=0.0, _dense_u=lambda t: (t @ u).astype(f32),
                             _dense_v=lambda t: (t @ v + b).astype(f32))
  y, line = run('layers/keras/interaction.py', 'Cross', 'call', _make_tf({}), me, [x0, x])
  out['cases']['keras_cross_lowrank'] = {'ref': 'layers/keras/interaction.py:%d' % line, 'x0': x0.tolist(),
                                         'x': x.tolist(), 'u': u.tolist(), 'v': v.tolist(), 'b': b.tolist(),
                                         'y': y.astype(f32).tolist()}
  # ---- DIN target attention (both restatements in the reference) ----
  B, T, D = 6, 5, 4
  key = rng.normal(size=(B, D)).astype(f32)
  hist = rng.normal(size=(B, T, D)).astype(f32)
  lens = np.array([0, 5, 3, 1, 2, 5], np.int64)       # max == T (sequence_mask has no maxlen); row 0 fully padded
  mlp = _mlp(rng, [4 * D, 8, 1])
  dnn_mod = types.SimpleNamespace(DNN=_NumpyDNN)
  deep_fea = {'key': key, 'hist_seq_emb': hist, 'hist_seq_len': lens, 'aux_hist_seq_emb_list': []}
  me = types.SimpleNamespace(_kernel_regularizer=None, _l2_reg=None, _is_training=True)
  y, line = run('layers/sequence_feature_layer.py', 'SequenceFeatureLayer', 'target_attention', _make_tf({}), me,
                mlp, deep_fea, 'din', dnn=dnn_mod)
  y2, line2 = run('model/multi_tower_din.py', 'MultiTowerDIN', 'din', _make_tf({}), me, mlp, deep_fea, 'din', dnn=dnn_mod)
  assert np.array_equal(y, y2)
  out['cases']['din_target_attention'] = {
      'ref': 'layers/sequence_feature_layer.py:%d, model/multi_tower_din.py:%d' % (line, line2), 'key': key.tolist(),
      'hist': hist.tolist(), 'lens': lens.tolist(), 'mlp': _mlp_json(mlp), 'y': y.astype(f32).tolist()}
  # ---- MMOE.__call__ (+ gate) ----
  B, d_in, E, n_task = 5, 10, 3, 2
  x = rng.normal(size=(B, d_in)).astype(f32)
  experts = {'mmoe/expert_%d' % e: _mlp(rng, [d_in, 8, 6]) for e in range(E)}
  variables = {}
  for t in range(n_task):
    variables['mmoe/gate_%d/dnn/kernel' % t] = rng.normal(0, 0.5, (d_in, E)).astype(f32)
    variables['mmoe/gate_%d/dnn/bias' % t] = rng.normal(0, 0.1, E).astype(f32)
  tf = _make_tf(variables)
  gate_code, _ = _function('layers/mmoe.py', 'MMOE', 'gate')
  ns = {'tf': tf}
  exec(gate_code, ns)
  me = types.SimpleNamespace(_num_expert=E, _num_task=n_task, _expert_dnn_configs=[experts] * E, _l2_reg=None,
                             _name='mmoe', _is_training=True)
  me.gate = lambda unit, fea, name: ns['gate'](me, unit, fea, name)
  ys, line = run('layers/mmoe.py', 'MMOE', '__call__', tf, me, x, dnn=dnn_mod)
  out['cases']['mmoe'] = {
      'ref': 'layers/mmoe.py:%d' % line, 'x': x.tolist(),
      'experts': [_mlp_json(experts['mmoe/expert_%d' % e]) for e in range(E)],
      'gates': [{'w': variables['mmoe/gate_%d/dnn/kernel' % t].tolist(), 'b': variables['mmoe/gate_%d/dnn/bias' % t].tolist()}
                for t in range(n_task)],
      'y': [np.asarray(t_, f32).tolist() for t_ in ys]}
  # ---- match model: list-wise similarity, duplicate-item masking, softmax, list-wise loss ----
  B, H, temperature = 6, 4, 0.2
  def unit(a):
    return (a / np.sqrt((a * a).sum(1, keepdims=True))).astype(f32)
  user, item = unit(rng.normal(size=(B, H))), unit(rng.normal(size=(B, H)))
  item[4] = item[1]                                     # rows 1 and 4 are the same item
  item_ids = np.array([11, 22, 33, 44, 22, 55], np.int64)
  sample_w = np.array([1, 2, 1, 0.5, 1, 1], f32)
  tf = _make_tf({})
  os.environ['tf.estimator.mode'] = 'train'
  me = types.SimpleNamespace(_model_config=types.SimpleNamespace(ignore_in_batch_neg_sam=False), _item_ids=item_ids,
                             _feature_dict={})
  sim, l0 = run('model/match_model.py', 'MatchModel', '_list_wise_sim', tf, me, user, item, os=os, logging=logging)
  sim = (sim / f32(temperature)).astype(f32)           # model/dssm.py:70 `self.sim(...) / temperature`
  logits, l1 = run('model/match_model.py', 'MatchModel', '_mask_in_batch', tf, me, sim)
  probs = tf.nn.softmax(logits)                          # model/dssm.py:93
  loss_type = types.SimpleNamespace(SOFTMAX_CROSS_ENTROPY=4)
  me = types.SimpleNamespace(_loss_type=4, _prediction_dict={'probs': probs, 'user_tower_emb': user, 'item_tower_emb': item},
                             _sample_weight=sample_w, _loss_dict={}, _model_config=None)
  losses, l2 = run('model/match_model.py', 'MatchModel', '_build_list_wise_loss_graph', tf, me, LossType=loss_type,
                   logging=logging)
  out['cases']['match_listwise'] = {
      'ref': 'model/match_model.py:%d,%d,%d' % (l1, l0, l2), 'user': user.tolist(), 'item': item.tolist(),
      'item_ids': item_ids.tolist(), 'sample_weight': sample_w.tolist(), 'temperature': temperature,
      'sim': sim.tolist(), 'probs': np.asarray(probs, f32).tolist(),
      'cross_entropy_loss': float(losses['cross_entropy_loss']), 'reg_pos_loss': float(losses['reg_pos_loss'])}
  # ---- exponential_decay_with_burnin ----
  sched = []
  for base, dsteps, factor, b_lr, b_steps, min_lr, stair in [
      (0.001, 1000, 0.5, 0.0, 0, 1e-5, True), (0.01, 100, 0.9, 0.001, 50, 1e-4, True),
      (0.01, 100, 0.9, 0.0, 30, 0.0, False), (0.05, 7, 0.3, 0.2, 10, 1e-3, False)]:
    steps = [0, 1, 5, 9, 10, 11, 29, 30, 31, 49, 50, 51, 99, 100, 149, 150, 151, 999, 1000, 1001, 2500, 100000]
    vals = []
    for st in steps:
      v, line = run('core/learning_schedules.py', None, 'exponential_decay_with_burnin', _make_tf({}), None,
                    np.int64(st), base, dsteps, factor, b_lr, b_steps, min_lr, stair)
      vals.append(float(v))
    sched.append({'initial_learning_rate': base, 'decay_steps': dsteps, 'decay_factor': factor,
                  'burnin_learning_rate': b_lr, 'burnin_steps': b_steps, 'min_learning_rate': min_lr,
                  'staircase': stair, 'steps': steps, 'lr': vals})
  out['cases']['lr_exponential_decay_with_burnin'] = {'ref': 'core/learning_schedules.py:%d' % line, 'schedules': sched}
  # ---- lazy Adam sparse rule (unique indices with already-summed gradients, as TF hands them over) ----
  V, D, lr, b1, b2, eps = 20, 4, 0.01, 0.9, 0.999, 1e-8

  class Var(object):
    dtype = types.SimpleNamespace(base_dtype=f32)

    def __init__(self, a):
      self.a = a
  var, m, v = Var(rng.normal(0, 0.1, (V, D)).astype(f32)), Var(np.zeros((V, D), f32)), Var(np.zeros((V, D), f32))
  w0 = var.a.copy()

  def scatter_update(ref, idx, vals):
    ref.a[idx] = vals
    return ref

  def scatter_add(ref, idx, vals):
    np.add.at(ref.a, idx, vals.astype(f32))
    return ref
  mods = dict(math_ops=types.SimpleNamespace(cast=lambda x, dt: dt(x), sqrt=lambda x: np.sqrt(x, dtype=f32)),
              array_ops=types.SimpleNamespace(gather=lambda ref, idx: ref.a[idx]),
              state_ops=types.SimpleNamespace(scatter_update=scatter_update),
              control_flow_ops=types.SimpleNamespace(group=lambda *a: None))
  steps = []
  p1, p2 = b1, b2                                      # beta power accumulators start at beta (adam_s.py:117-129)
  for _ in range(3):
    idx = np.sort(rng.choice(V, 7, replace=False)).astype(np.int64)
    g = rng.normal(0, 1, (7, D)).astype(f32)
    me = types.SimpleNamespace(_get_beta_accumulators=lambda: (p1, p2), _lr_t=lr, _beta1_t=b1, _beta2_t=b2,
                               _epsilon_t=eps, get_slot=lambda var_, name: {'m': m, 'v': v}[name])
    _, line = run('compat/adam_s.py', 'AdamOptimizerS', '_apply_sparse_shared', None, me, g, var, idx, scatter_add, **mods)
    p1, p2 = p1 * b1, p2 * b2                          # _finish, adam_s.py:236-245
    steps.append({'indices': idx.tolist(), 'grad': g.tolist(), 'w': var.a.tolist(), 'm': m.a.tolist(), 'v': v.a.tolist()})
  out['cases']['lazy_adam_sparse'] = {'ref': 'compat/adam_s.py:%d' % line, 'lr': lr, 'beta1': b1, 'beta2': b2,
                                      'epsilon': eps, 'w0': w0.tolist(), 'steps': steps}
  # ---- DeepFM.build_predict_graph: the head wiring, with and without final_dnn ----
  B, F, D = 6, 5, 4
  wide = rng.normal(size=(B, F)).astype(f32)
  fm_feas = [rng.normal(size=(B, D)).astype(f32) for _ in range(F)]
  deep = np.concatenate(fm_feas, axis=1)

  class Cfg(dict):
    hidden_units = ()
  fm_code, _ = _function('layers/fm.py', 'FM', '__call__')
  fm_ns = {'tf': _make_tf({})}
  exec(fm_code, fm_ns)

  class RefFM(object):
    def __init__(self, name='fm'):
      self._name = name
    __call__ = fm_ns['__call__']
  for final in (True, False):
    deep_mlp = _mlp(rng, [F * D, 8, 6])
    final_mlp = _mlp(rng, [1 + D + 6, 7]) if final else []
    variables = {}
    for nm, d_in in ((('output', 7),) if final else (('deep_logits', 6),)):
      variables[nm + '/kernel'] = rng.normal(0, 0.4, (d_in, 1)).astype(f32)
      variables[nm + '/bias'] = rng.normal(0, 0.1, 1).astype(f32)
    dnn_cfg, final_cfg = Cfg(deep_feature=deep_mlp), Cfg(final_dnn=final_mlp)
    final_cfg.hidden_units = (7,) if final else ()
    got = {}
    me = types.SimpleNamespace(_num_class=1, _wide_output_dim=1, _wide_features=wide, _fm_features=fm_feas,
                               _deep_features=deep, _l2_reg=None, _is_training=True, _prediction_dict={},
                               _model_config=types.SimpleNamespace(dnn=dnn_cfg, final_dnn=final_cfg),
                               _add_to_prediction_dict=lambda o: got.update(logits=o))
    _, line = run('model/deepfm.py', 'DeepFM', 'build_predict_graph', _make_tf(variables), me,
                  fm=types.SimpleNamespace(FM=RefFM), dnn=types.SimpleNamespace(DNN=_NumpyDNN))
    out['cases']['deepfm_head_final' if final else 'deepfm_head_plain'] = {
        'ref': 'model/deepfm.py:%d' % line, 'wide': wide.tolist(), 'deep': deep.tolist(), 'n_field': F, 'dim': D,
        'dnn': _mlp_json(deep_mlp), 'final_dnn': _mlp_json(final_mlp),
        'head': {k: v.tolist() for k, v in variables.items()}, 'logits': np.asarray(got['logits'], f32).tolist()}


if __name__ == '__main__':
  if not os.path.isdir(REF):
    sys.exit('reference checkout not mounted: nothing to do')
  main()
