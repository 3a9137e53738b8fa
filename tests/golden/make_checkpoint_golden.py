"""Golden data for the sharded-table checkpoint layout, produced by EXECUTING the reference saver's own python
functions: `_save_embed` and `_load_embed`, the closures inside EmbeddingParallelSaver._save_dense_embedding /
_load_dense_embedding (compat/embedding_parallel_saver.py:99-173), with `gfile` mapped onto os / glob and
`hvd.rank() / hvd.size()` supplied per worker.

A 41-row, 4-wide table (odd row count: the last worker's shard carries a padding row) is saved by 2 workers and
by 3 workers; every shard is then re-loaded for 1, 2, 3 and 4 workers.  The JSON holds the file names, the sha256
of every part file and every re-sharded result.

  python tests/golden/make_checkpoint_golden.py -> tests/golden/reference_checkpoint.json
replayed by tests/test_checkpoint_golden.py on easyrec_b200.checkpoint (python save, native er_load_embed)."""
import ast
import glob
import hashlib
import json
import logging
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF = '/root/reference/easy_rec/python'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_checkpoint.json')
V, D = 41, 4
VAR = 'input_layer/all_fea/embedding_weights:0'


def table():
  return np.random.default_rng(99).normal(size=(V, D)).astype(np.float32)


def shard(full, rank, world):
  """what worker `rank` of `world` holds: (V + world - 1) // world rows, global row j * world + rank at j."""
  part = np.zeros(((V + world - 1) // world, D), np.float32)
  rows = full[rank::world]
  part[:len(rows)] = rows
  return part


class _NP(object):
  """numpy with the long-removed alias the saver still spells (np.object)."""
  object = object

  def __getattr__(self, k):
    return getattr(np, k)


class _GFile(object):
  Exists = staticmethod(os.path.exists)
  MakeDirs = staticmethod(lambda d: os.makedirs(d, exist_ok=True))
  Glob = staticmethod(lambda p: sorted(glob.glob(p)))
  DeleteRecursively = staticmethod(os.remove)
  GFile = staticmethod(open)


def _closure(outer, inner, rank, world):
  """the nested function `inner` of EmbeddingParallelSaver.<outer>, compiled from the reference source."""
  src = open(os.path.join(REF, 'compat/embedding_parallel_saver.py')).read()
  tree = ast.parse(src)
  helper = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == '_get_embed_part_id']
  cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'EmbeddingParallelSaver'][0]
  meth = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == outer][0]
  fn = [n for n in ast.walk(meth) if isinstance(n, ast.FunctionDef) and n.name == inner][0]
  ns = {'np': _NP(), 'gfile': _GFile, 'logging': logging,
        'hvd': types.SimpleNamespace(rank=lambda: rank, size=lambda: world)}
  exec(compile(ast.Module(body=helper + [fn], type_ignores=[]), 'embedding_parallel_saver.py', 'exec'), ns)
  return ns[inner], fn.lineno


def save_with_reference(ckpt, full, world):
  files = []
  for rank in range(world):
    fn, line = _closure('_save_dense_embedding', '_save_embed', rank, world)
    files.append(str(fn(shard(full, rank, world), ckpt.encode(), VAR.encode())[0]))
  return files, line


def load_with_reference(ckpt, rank, world):
  fn, line = _closure('_load_dense_embedding', '_load_embed', rank, world)
  part_size = (V + world - 1) // world
  return fn(None, D, part_size, rank, world, ckpt.encode(), VAR.encode()), line


def main():
  full = table()
  out = {'generator': 'tests/golden/make_checkpoint_golden.py', 'var_name': VAR, 'rows': V, 'dim': D, 'saves': []}
  d = tempfile.mkdtemp()
  try:
    ckpt = os.path.join(d, 'model.ckpt-100')
    for world_old in (3, 2):           # 3 first: saving with 2 afterwards must delete the stale part-2 file
      files, l_save = save_with_reference(ckpt, full, world_old)
      on_disk = sorted(glob.glob(ckpt + '-embedding/*'))
      entry = {'world': world_old, 'files': [os.path.relpath(f, d) for f in on_disk],
               'sha256': [hashlib.sha256(open(f, 'rb').read()).hexdigest() for f in on_disk], 'loads': []}
      for world_new in (1, 2, 3, 4):
        for rank in range(world_new):
          vals, l_load = load_with_reference(ckpt, rank, world_new)
          entry['loads'].append({'world': world_new, 'rank': rank, 'vals': np.asarray(vals, np.float32).tolist()})
      out['saves'].append(entry)
    out['ref'] = 'compat/embedding_parallel_saver.py:%d (_save_embed), :%d (_load_embed)' % (l_save, l_load)
  finally:
    shutil.rmtree(d)
  json.dump(out, open(OUT, 'w'))
  print('wrote', OUT, [(s['world'], s['files']) for s in out['saves']])


if __name__ == '__main__':
  if not os.path.isdir(REF):
    sys.exit('reference checkout not mounted: nothing to do')
  main()
