"""RawFeature.normalizer_fn (input/input.py:133-137, 642-646): a function applied to the raw value after the min-max
normalisation and before the feature is bucketized / projected / fed as a dense input.  The reference resolves the
string with utils/load_class.py:27-50 `load_by_path`: a `lambda x: ...` expression is eval'd with `tf` in scope, a
dotted path (`tf.math.log1p`) is located as a function.  Here the same strings evaluate over torch tensors (the
device-side dense matrix of InputLayer) or numpy arrays (the host-side bucketizer of the readers): a small `tf`
namespace maps the elementwise TensorFlow calls such expressions use.  Configs are trusted input there and here."""
import math
import pydoc

import numpy as np


class _Ops(object):
  """elementwise tf.* / tf.math.* / tf.nn.* calls over one array library"""

  def __init__(self, xp):
    self._xp = xp
    self.math = self
    self.nn = self
    self.float32 = xp.float32

  def _f(self, name, alt=None):
    return getattr(self._xp, name, None) or getattr(self._xp, alt)

  def log(self, x): return self._xp.log(x)                      # noqa: E704
  def log1p(self, x): return self._xp.log1p(x)                  # noqa: E704
  def exp(self, x): return self._xp.exp(x)                      # noqa: E704
  def expm1(self, x): return self._xp.expm1(x)                  # noqa: E704
  def sqrt(self, x): return self._xp.sqrt(x)                    # noqa: E704
  def rsqrt(self, x): return 1.0 / self._xp.sqrt(x)             # noqa: E704
  def square(self, x): return x * x                             # noqa: E704
  def abs(self, x): return self._xp.abs(x)                      # noqa: E704
  def sign(self, x): return self._xp.sign(x)                    # noqa: E704
  def floor(self, x): return self._xp.floor(x)                  # noqa: E704
  def ceil(self, x): return self._xp.ceil(x)                    # noqa: E704
  def round(self, x): return self._xp.round(x)                  # noqa: E704
  def tanh(self, x): return self._xp.tanh(x)                    # noqa: E704
  def sigmoid(self, x): return 1.0 / (1.0 + self._xp.exp(-x))   # noqa: E704
  def relu(self, x): return self.maximum(x, 0.0)                # noqa: E704
  def pow(self, x, y): return x ** y                            # noqa: E704
  def negative(self, x): return -x                              # noqa: E704
  def reciprocal(self, x): return 1.0 / x                       # noqa: E704
  def add(self, a, b): return a + b                             # noqa: E704
  def subtract(self, a, b): return a - b                        # noqa: E704
  def multiply(self, a, b): return a * b                        # noqa: E704
  def divide(self, a, b): return a / b                          # noqa: E704
  def identity(self, x): return x                               # noqa: E704

  def maximum(self, a, b):
    if self._xp is np:
      return np.maximum(a, b)
    return self._xp.clamp(a, min=b) if not hasattr(b, 'shape') else self._xp.maximum(a, b)

  def minimum(self, a, b):
    if self._xp is np:
      return np.minimum(a, b)
    return self._xp.clamp(a, max=b) if not hasattr(b, 'shape') else self._xp.minimum(a, b)

  def clip_by_value(self, x, lo, hi):
    return self.minimum(self.maximum(x, lo), hi)

  def cast(self, x, dtype=None):
    return x

  to_float = identity


def load(path, backend):
  """normalizer_fn string -> function over `backend` arrays ('torch' | 'numpy'); float32 in, float32 out."""
  path = (path or '').strip()
  if not path:
    return None
  if backend == 'torch':
    import torch
    xp = torch
  else:
    xp = np
  tf = _Ops(xp)
  if 'lambda' in path:
    fn = eval(path, {'tf': tf, 'math': math, 'np': np, '__builtins__': {'abs': abs, 'min': min, 'max': max, 'float': float,
                                                                     'int': int, 'pow': pow}})  # noqa: S307
  else:
    parts = path.split('.')
    if parts[0] in ('tf', 'tensorflow'):
      fn = tf
      for p in parts[1:]:
        if p in ('compat', 'v1', 'v2'):
          continue
        fn = getattr(fn, p, None)
        if fn is None:
          raise NotImplementedError('normalizer_fn %r: tf call not in the elementwise set of normalizer.py' % path)
    else:
      fn = pydoc.locate(path)   # a user function by its import path, as load_by_path does
      if fn is None:
        raise ValueError('normalizer_fn %r cannot be located' % path)

  def apply(x):
    y = fn(x)
    return y.to(x.dtype) if backend == 'torch' else np.asarray(y, np.float32)
  return apply
