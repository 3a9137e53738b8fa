"""Golden vectors for the interaction formulas, produced by EXECUTING THE REFERENCE'S OWN CODE.

TensorFlow is not installable here, but these functions only use a handful of tensor ops whose semantics
are unambiguous (stack, square, reduce_sum, subtract, add, matmul, band_part, boolean_mask, where, ...).
The reference source is loaded from /root/reference (never copied), the function bodies are taken as they
are (ast), and they run against `_tf`, a numpy implementation of exactly those ops, in float32:

  layers/fm.py:20-26                         FM.__call__
  model/dcn.py:32-45                         DCN._cross_net      (tf.get_variable -> supplied w / b)
  layers/keras/interaction.py:24-44          keras FM.call
  layers/keras/interaction.py:47-128         DotInteraction.call

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
  tf.reduce_sum = lambda x, axis=None, keepdims=False: np.sum(x, axis=axis, keepdims=keepdims, dtype=np.float32)
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
  return tf


def _function(path, cls, name):
  """the reference's function object, compiled from its own source text (no package import, no TF)."""
  src = open(os.path.join(REF, path)).read()
  tree = ast.parse(src)
  for node in tree.body:
    if isinstance(node, ast.ClassDef) and node.name == cls:
      for fn in node.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == name:
          mod = ast.Module(body=[fn], type_ignores=[])
          return compile(mod, os.path.join(REF, path), 'exec'), fn.lineno
  raise KeyError('%s.%s not found in %s' % (cls, name, path))


def run(path, cls, name, tf, self_obj, *args):
  code, line = _function(path, cls, name)
  ns = {'tf': tf}
  exec(code, ns)
  return ns[name](self_obj, *args), line


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
  json.dump(out, open(OUT, 'w'))
  print('wrote', OUT, {k: np.asarray(v['y']).shape for k, v in out['cases'].items()})


if __name__ == '__main__':
  if not os.path.isdir(REF):
    sys.exit('reference checkout not mounted: nothing to do')
  main()
