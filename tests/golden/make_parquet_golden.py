"""Golden batches from the reference's OWN Parquet loader, run unmodified in this container.

input/load_parquet.py has no TensorFlow dependency: it is loaded by file path (the `easy_rec` package itself
cannot be imported without TF) and `load_data_proc` (load_parquet.py:139-317) runs in-process with plain
in-memory stand-ins for the multiprocessing queues.  `ParquetInput._to_fea_dict` (input/parquet_input.py:201-239,
the `vals % num_buckets` step) is taken from its source with `ast` and run on the loader's batches.

  python tests/golden/make_parquet_golden.py  ->  tests/golden/reference_parquet_batches.json
replayed by tests/test_parquet_golden.py against easyrec_b200.input.readers.ParquetInput and the oracle's
bucket rule."""
import ast
import importlib.util
import json
import os
import queue
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import parquet_case as case  # noqa: E402

REF = '/root/reference/easy_rec/python'
OUT = os.path.join(HERE, 'reference_parquet_batches.json')


class _Que(object):
  """the queue surface load_data_proc touches."""

  def __init__(self, items=()):
    self.items = list(items)

  def get(self, block=True, timeout=None):
    if not self.items:
      raise queue.Empty()
    return self.items.pop(0)

  def put(self, item, timeout=None):
    self.items.append(item)

  def qsize(self):
    return len(self.items)

  def close(self, wait_send_finish=True):
    pass


def _to_fea_dict():
  src = open(os.path.join(REF, 'input/parquet_input.py')).read()
  for node in ast.parse(src).body:
    if isinstance(node, ast.ClassDef) and node.name == 'ParquetInput':
      for fn in node.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == '_to_fea_dict':
          ns = {}
          exec(compile(ast.Module(body=[fn], type_ignores=[]), 'parquet_input.py', 'exec'), ns)
          return ns['_to_fea_dict'], fn.lineno
  raise KeyError('_to_fea_dict')


def run(drop_remainder):
  spec = importlib.util.spec_from_file_location('ref_load_parquet', os.path.join(REF, 'input/load_parquet.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  with tempfile.TemporaryDirectory() as d:
    paths = case.write_files(d)
    data_que = _Que()
    dense_cfgs = [types.SimpleNamespace(raw_input_dim=k) for _, k in case.DENSE]
    mod.load_data_proc(0, _Que(paths + [None]), data_que, _Que([True]), _Que(), case.BATCH, [case.LABEL],
                       list(case.SPARSE), [n for n, _ in case.DENSE], dense_cfgs, None, drop_remainder, 0, 1, True)
  fn, line = _to_fea_dict()
  me = types.SimpleNamespace(_sparse_fea_names=case.SPARSE, _dense_fea_names=[n for n, _ in case.DENSE], _has_ev=False,
                             _feature_configs=[types.SimpleNamespace(num_buckets=case.NUM_BUCKETS)] * 3 +
                             [types.SimpleNamespace(num_buckets=0)] * 2,
                             _label_fields=[case.LABEL], _reserve_fields=None)
  batches = []
  for item in data_que.items:
    if item is None:
      continue
    lens, raw_vals = item['sparse_fea']
    out = fn(me, item)
    vals, lens2 = out['feature']['sparse_fea']
    assert lens2 is lens
    batches.append({'lens': lens.tolist(), 'raw_vals': np.asarray(raw_vals, np.int64).tolist(),
                    'vals': np.asarray(vals, np.int64).tolist(),
                    'dense_fea': np.asarray(out['feature']['dense_fea'], np.float32).tolist(),
                    'label': np.asarray(out['label'][case.LABEL], np.float32).tolist()})
  return batches, line


def main():
  keep, line = run(False)
  drop, _ = run(True)
  json.dump({'generator': 'tests/golden/make_parquet_golden.py',
             'ref': 'input/load_parquet.py:139-317 (executed), input/parquet_input.py:%d (_to_fea_dict)' % line,
             'batch_size': case.BATCH, 'sparse': case.SPARSE, 'batches': keep, 'n_batches_drop_remainder': len(drop)},
            open(OUT, 'w'))
  print('wrote', OUT, [len(b['label']) for b in keep], len(drop))


if __name__ == '__main__':
  if not os.path.isdir(REF):
    sys.exit('reference checkout not mounted: nothing to do')
  main()
