"""Golden vectors for the activations, produced by EXECUTING THE REFERENCE'S OWN CODE (utils/activation.py).

`gelu` (:46-60), `swish` (:63-65) and `dice` (:13-43) are defined in the reference tree itself; `get_activation`
(:66-118) maps the config strings onto them and onto tf.nn functions.  TensorFlow is not installable here: the function
bodies are compiled from the reference's source text (ast, nothing copied) and run against a numpy shim of the few ops
they call (tanh, pow, sigmoid, name_scope; for dice: get_variable and layers.batch_normalization without centre /
scale, batch statistics, biased variance), in float32.  The name -> function map is recorded by giving the shim's
tf.nn.* entries their own names.

Run in the build container (reference mounted):  python tests/golden/make_activation_golden.py
-> tests/golden/reference_activations.json, replayed by tests/test_act_metrics_host.py (oracle, kernel source)."""
import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_formula_golden import _Scope, _function  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_activations.json')
PATH = 'utils/activation.py'


def _named(name):
  def f(x, *a, **k):
    raise RuntimeError('sentinel')
  f.__name__ = name
  return f


def make_tf(alphas=None):
  tf = types.SimpleNamespace()
  tf.__version__ = '1.15.5'
  f32 = np.float32
  tf.name_scope = lambda *a, **k: _Scope()
  tf.tanh = lambda x: np.tanh(x).astype(f32)
  tf.pow = lambda x, p: np.power(x, f32(p)).astype(f32)
  tf.sigmoid = lambda x: (f32(1) / (f32(1) + np.exp(-x))).astype(f32)
  tf.constant_initializer = lambda v: v
  tf.float32 = f32
  tf.get_variable = lambda name, shape, initializer=None, dtype=None: alphas

  def batch_normalization(inputs, axis=-1, epsilon=1e-3, center=True, scale=True, training=True):
    assert training and not center and not scale
    mu = inputs.mean(0, dtype=f32)
    var = ((inputs - mu) ** 2).mean(0, dtype=f32)
    return ((inputs - mu) / np.sqrt(var + f32(epsilon))).astype(f32)
  tf.layers = types.SimpleNamespace(batch_normalization=batch_normalization)
  tf.nn = types.SimpleNamespace(relu=_named('tf.nn.relu'), leaky_relu=_named('tf.nn.leaky_relu'), elu=_named('tf.nn.elu'),
                                selu=_named('tf.nn.selu'), swish=_named('tf.nn.swish'), sigmoid=_named('tf.nn.sigmoid'))
  tf.keras = types.SimpleNamespace(layers=types.SimpleNamespace(PReLU=lambda **kw: _named('tf.keras.layers.PReLU')))
  return tf


def main():
  rng = np.random.default_rng(7)
  x = np.concatenate([rng.normal(0, 2.5, 64), [0.0, -0.0, 1e-6, -1e-6, 8.0, -8.0]]).astype(np.float32)
  out = {'source': 'alibaba/EasyRec easy_rec/python/utils/activation.py executed on a numpy shim', 'x': x.tolist(), 'cases': {}}
  for name in ('gelu', 'swish'):
    code, line = _function(PATH, None, name)
    ns = {'tf': make_tf(), 'np': np}
    exec(code, ns)
    out['cases'][name] = {'ref': '%s:%d' % (PATH, line), 'y': np.asarray(ns[name](x), np.float32).tolist()}
  # dice over a [B, C] matrix with given alphas
  class _T(np.ndarray):   # a float32 array that also answers the one shape call the function makes
    def get_shape(self):
      return self.shape
  xm = rng.normal(0, 1.5, (16, 3)).astype(np.float32).view(_T)
  alphas = np.array([0.0, 0.25, -0.5], np.float32)
  code, line = _function(PATH, None, 'dice')
  ns = {'tf': make_tf(alphas), 'np': np}
  exec(code, ns)
  out['cases']['dice'] = {'ref': '%s:%d' % (PATH, line), 'x': np.asarray(xm).tolist(), 'alphas': alphas.tolist(),
                          'y': np.asarray(ns['dice'](xm), np.float32).tolist()}
  # config string -> function, as get_activation resolves it
  code, line = _function(PATH, None, 'get_activation')
  tf = make_tf()
  tf.tanh = _named('tf.tanh')
  ns = {'tf': tf, 'np': np, 'six': types.SimpleNamespace(string_types=(str,)), 'gelu': _named('gelu'), 'swish': _named('swish'),
        'dice': _named('dice'), 'load_by_path': lambda p: _named('load_by_path(%s)' % p)}
  exec(code, ns)
  table = {}
  for s in ('relu', 'Relu', 'gelu', 'leaky_relu', 'prelu', 'elu', 'selu', 'tanh', 'swish', 'sigmoid', 'linear', '', 'tf.nn.relu',
            'tf.nn.tanh'):
    fn = ns['get_activation'](s)
    table[s] = None if fn is None else fn.__name__
  out['cases']['get_activation'] = {'ref': '%s:%d' % (PATH, line), 'map': table}
  with open(OUT, 'w') as f:
    json.dump(out, f, indent=1)
  print('wrote', OUT, {k: (v if k == 'get_activation' else '...') for k, v in out['cases'].items()}['get_activation'])


if __name__ == '__main__':
  main()
