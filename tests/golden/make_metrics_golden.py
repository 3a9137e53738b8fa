"""Golden values for the grouped AUC, produced by EXECUTING the reference's `_separated_auc_impl`
(core/metrics.py:59-106): its `update_pyfunc` / `value_pyfunc` closures run as they are (sklearn is installed
here), `tf.py_func` is replaced by a recorder that hands the closures back.

  python tests/golden/make_metrics_golden.py -> tests/golden/reference_metrics.json
replayed by tests/test_metrics_golden.py on easyrec_b200.metrics.gauc."""
import ast
import json
import os
import sys
import types
from collections import defaultdict

import numpy as np
from sklearn import metrics as sklearn_metrics

REF = '/root/reference/easy_rec/python'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_metrics.json')


def reference_gauc(labels, predictions, keys, reduction, batch=37):
  src = open(os.path.join(REF, 'core/metrics.py')).read()
  fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == '_separated_auc_impl'][0]
  captured = []
  tf = types.SimpleNamespace(py_func=lambda f, inp, tout: captured.append(f) or f, float32=np.float32)
  ns = {'tf': tf, 'np': np, 'defaultdict': defaultdict, 'sklearn_metrics': sklearn_metrics}
  exec(compile(ast.Module(body=[fn], type_ignores=[]), 'metrics.py', 'exec'), ns)
  value_op, update_op = ns['_separated_auc_impl'](None, None, None, reduction)
  for i in range(0, len(labels), batch):          # the update op runs once per evaluation batch
    update_op(labels[i:i + batch], predictions[i:i + batch], keys[i:i + batch])
  return float(value_op()), fn.lineno


def data():
  rng = np.random.default_rng(12)
  n = 400
  keys = rng.integers(0, 25, n).astype(np.int64)
  keys[keys == 3] = 4                                     # a key that never appears
  labels = (rng.uniform(size=n) < 0.35).astype(np.int64)
  labels[keys == 7] = 1                                   # single-class groups are skipped
  labels[keys == 9] = 0
  preds = np.round(rng.uniform(size=n), 2).astype(np.float32)   # rounded: ties inside groups
  return labels, preds, keys


def main():
  labels, preds, keys = data()
  out = {'generator': 'tests/golden/make_metrics_golden.py', 'labels': labels.tolist(), 'predictions': preds.tolist(),
         'keys': keys.tolist(), 'gauc': {}}
  for reduction in ('mean', 'mean_by_sample_num', 'mean_by_positive_num'):
    out['gauc'][reduction], line = reference_gauc(labels, preds, keys, reduction)
  out['gauc_all_single_class'], _ = reference_gauc(np.ones(10, np.int64), preds[:10], keys[:10], 'mean')
  out['ref'] = 'core/metrics.py:%d' % line
  json.dump(out, open(OUT, 'w'))
  print('wrote', OUT, out['gauc'], out['gauc_all_single_class'])


if __name__ == '__main__':
  if not os.path.isdir(REF):
    sys.exit('reference checkout not mounted: nothing to do')
  main()
