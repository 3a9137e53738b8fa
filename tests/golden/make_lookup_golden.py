"""Golden vectors for the lookup / pooling / sharding rules, produced by EXECUTING the reference's own functions
from compat/feature_column/feature_column.py against a numpy shim of the TF ops they call:

  :189-244  embedding_lookup_ragged      unique -> gather -> sparse segment sum / mean / sqrtn (unweighted branch)
  :248-357  embedding_parallel_lookup    packed 'sparse_fea' (ids, lens) form; with hvd.size() == 2 the two ranks
                                         run as two threads and `hvd.alltoall` really exchanges their buffers:
                                         owner = id % N, local row = int64(id / N) (float true-divide), dynamic
                                         partition / stitch, sparse_segment_sum, [B, n_feat * D] layout

  python tests/golden/make_lookup_golden.py -> tests/golden/reference_lookup.json
replayed by tests/test_oracle_golden.py on the oracle (the GPU tests compare the kernels with the oracle)."""
import ast
import json
import os
import sys
import threading
import types

import numpy as np

REF = '/root/reference/easy_rec/python'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_lookup.json')
f32 = np.float32


class Table(np.ndarray):
  def get_shape(self):
    return self.shape


def _unique(x):
  """tf.unique: distinct values in order of first occurrence + index of every input in that list."""
  seen, uniq, idx = {}, [], []
  for v in np.asarray(x).tolist():
    if v not in seen:
      seen[v] = len(uniq)
      uniq.append(v)
    idx.append(seen[v])
  return np.array(uniq, np.int64), np.array(idx, np.int32)


def _segment_sum(data, seg, name=None, n=None):
  seg = np.asarray(seg)
  n = int(seg.max()) + 1 if n is None else n
  out = np.zeros((n,) + data.shape[1:], data.dtype)
  for i, s in enumerate(seg):            # in order, like the CPU kernel
    out[s] += data[i]
  return out


def _sparse_segment(kind):
  def fn(data, indices, segment_ids, name=None, num_segments=None):
    seg = np.asarray(segment_ids)
    g = np.asarray(data)[np.asarray(indices)]
    s = _segment_sum(g, seg, n=num_segments)
    cnt = _segment_sum(np.ones((len(seg), 1), f32), seg, n=s.shape[0])
    if kind == 'mean':
      s = np.where(cnt > 0, s / np.maximum(cnt, 1), 0).astype(f32)
    elif kind == 'sqrtn':
      s = np.where(cnt > 0, s / np.sqrt(np.maximum(cnt, 1)), 0).astype(f32)
    return s
  return fn


def _namespace(hvd):
  dtypes = types.SimpleNamespace(int32=np.int32, int64=np.int64, float32=f32, float16=np.float16, bfloat16=np.float16)

  def dynamic_partition(data, partitions, num):
    return [np.asarray(data)[np.asarray(partitions) == p] for p in range(num)]

  def parallel_dynamic_stitch(indices, data, name=None):
    n = sum(len(i) for i in indices)
    out = np.zeros((n,) + data[0].shape[1:], data[0].dtype)
    for i, d in zip(indices, data):
      out[i] = d
    return out

  def split(x, num_or_size_splits, axis=0):
    if isinstance(num_or_size_splits, int):
      return np.split(x, num_or_size_splits, axis=axis)
    return np.split(x, np.cumsum(num_or_size_splits)[:-1], axis=axis)

  array_ops = types.SimpleNamespace(
      unique=_unique, gather=lambda p, i: np.asarray(p)[np.asarray(i)].view(Table),
      expand_dims=lambda x, a: np.expand_dims(x, a), concat=lambda xs, axis: np.concatenate(xs, axis=axis),
      searchsorted=lambda a, v, side='left': np.searchsorted(a, v, side=side).astype(np.int32),
      size=lambda x: np.asarray(x).size, shape=lambda x: np.array(np.asarray(x).shape[:1], np.int32),
      split=split, reshape=lambda x, s: np.reshape(x, s), transpose=lambda x, perm: np.transpose(x, perm),
      squeeze=lambda x, axis: np.squeeze(x, axis=axis))
  math_ops = types.SimpleNamespace(
      cast=lambda x, dt: np.asarray(x).astype(dt), cumsum=lambda x: np.cumsum(x), range=lambda n: np.arange(int(n)),
      segment_sum=_segment_sum, pow=lambda x, p: np.power(x, p).astype(f32), sqrt=lambda x: np.sqrt(x).astype(f32),
      div_no_nan=lambda a, b, name=None: np.where(b != 0, a / np.where(b != 0, b, 1), 0).astype(f32),
      sparse_segment_sum=_sparse_segment('sum'), sparse_segment_mean=_sparse_segment('mean'),
      sparse_segment_sqrt_n=_sparse_segment('sqrtn'))
  embedding_ops = types.SimpleNamespace(
      embedding_lookup=lambda w, ids, partition_strategy='mod', max_norm=None: np.asarray(w)[np.asarray(ids)].view(Table))

  class _Never(object):
    pass
  return dict(dtypes=dtypes, array_ops=array_ops, math_ops=math_ops, embedding_ops=embedding_ops,
              data_flow_ops=types.SimpleNamespace(dynamic_partition=dynamic_partition,
                                                  parallel_dynamic_stitch=parallel_dynamic_stitch),
              dynamic_variable=types.SimpleNamespace(DynamicVariable=_Never),
              sparse_tensor_lib=types.SimpleNamespace(SparseTensor=_Never), hvd=hvd,
              ops=types.SimpleNamespace())


def _load(name, ns):
  src = open(os.path.join(REF, 'compat/feature_column/feature_column.py')).read()
  for node in ast.parse(src).body:
    if isinstance(node, ast.FunctionDef) and node.name == name:
      exec(compile(ast.Module(body=[node], type_ignores=[]), 'feature_column.py', 'exec'), ns)
      return ns[name], node.lineno
  raise KeyError(name)


class _Exchange(object):
  """hvd.alltoall between threads: rank r's i-th split goes to rank i; returns (received, received sizes)."""

  def __init__(self, n):
    self.n = n
    self.box = {}
    self.barrier = threading.Barrier(n)

  def hvd(self, rank):
    ex = self

    def alltoall(tensor, splits):
      tensor = np.asarray(tensor)
      splits = np.asarray(splits).astype(np.int64)
      parts = np.split(tensor, np.cumsum(splits)[:-1])
      ex.barrier.wait()
      for dst, p in enumerate(parts):
        ex.box[(rank, dst)] = p
      ex.barrier.wait()
      got = [ex.box[(src, rank)] for src in range(ex.n)]
      ex.barrier.wait()
      return np.concatenate(got, axis=0), np.array([len(g) for g in got], np.int32)
    return types.SimpleNamespace(size=lambda: ex.n, rank=lambda: rank, alltoall=alltoall)


def main():
  rng = np.random.default_rng(4242)
  out = {'generator': 'tests/golden/make_lookup_golden.py', 'cases': {}}
  # ---- embedding_lookup_ragged: every combiner, with and without weights ----
  V, D, B = 30, 4, 6
  table = rng.normal(size=(V, D)).astype(f32)
  lens = np.array([2, 0, 3, 1, 0, 2], np.int32)
  ids = rng.integers(0, V, int(lens.sum())).astype(np.int64)
  ids[1] = ids[0]                                         # a repeated id inside one bag
  seg = np.repeat(np.arange(B), lens).astype(np.int64)
  fn, line = _load('embedding_lookup_ragged', _namespace(None))
  cases = []
  for combiner in ('sum', 'mean', 'sqrtn'):
    # unweighted only: the weighted branch calls expand_dims(weights [n], axis=2) on rank-2 embeddings
    # (:212), which TF rejects, so the reference has no defined weighted behaviour on this function
    r_ids = types.SimpleNamespace(value_rowids=lambda: seg, flat_values=ids)
    y = fn(table, r_ids, None, combiner)
    cases.append({'combiner': combiner, 'y': np.asarray(y, f32).tolist()})
  out['cases']['embedding_lookup_ragged'] = {
      'ref': 'compat/feature_column/feature_column.py:%d' % line, 'table': table.tolist(), 'ids': ids.tolist(),
      'lens': lens.tolist(), 'outputs': cases}
  # ---- embedding_parallel_lookup on 2 ranks (and the single-rank branch) ----
  V, D, B, F, N = 41, 4, 4, 3, 2                          # odd V: shards of (V+N-1)//N = 21 rows, rank 1 uses 20
  full = rng.normal(size=(V, D)).astype(f32)
  per_rank = []
  for r in range(N):
    lens = rng.integers(0, 3, F * B).astype(np.int32)
    lens[-1] = 2                                          # last bag not empty (output rows = max segment + 1)
    per_rank.append((rng.integers(0, V, int(lens.sum())).astype(np.int64), lens))
  shard_rows = (V + N - 1) // N
  results = [None] * N
  ex = _Exchange(N)

  def worker(r):
    ns = _namespace(ex.hvd(r))
    fn, _ = _load('embedding_parallel_lookup', ns)
    local = np.zeros((shard_rows, D), f32)
    local[:len(full[r::N])] = full[r::N]                  # local row j of rank r holds global row j * N + r
    tensors = {}
    y = fn(local.view(Table), {'sparse_fea': per_rank[r]}, list(range(F)), True, tensors, B)
    results[r] = (np.asarray(y, f32), {k: np.asarray(v, f32) for k, v in tensors.items()})
  threads = [threading.Thread(target=worker, args=(r,)) for r in range(N)]
  [t.start() for t in threads]
  [t.join() for t in threads]
  assert all(r is not None for r in results)
  single_ns = _namespace(types.SimpleNamespace(size=lambda: 1))
  fn1, line = _load('embedding_parallel_lookup', single_ns)
  singles = [np.asarray(fn1(full.view(Table), {'sparse_fea': per_rank[r]}, list(range(F)), True, None, B), f32)
             for r in range(N)]
  for r in range(N):
    assert np.allclose(results[r][0], singles[r], atol=1e-6)   # sharding must not change the result
    for f in range(F):
      assert np.array_equal(results[r][1][f], results[r][0][:, f * D:(f + 1) * D])
  out['cases']['embedding_parallel_lookup'] = {
      'ref': 'compat/feature_column/feature_column.py:%d' % line, 'world': N, 'batch_size': B, 'n_feature': F,
      'table': full.tolist(), 'shard_rows': shard_rows,
      'ranks': [{'ids': per_rank[r][0].tolist(), 'lens': per_rank[r][1].tolist(), 'y': results[r][0].tolist()}
                for r in range(N)]}
  json.dump(out, open(OUT, 'w'))
  print('wrote', OUT, sorted(out['cases']))


if __name__ == '__main__':
  if not os.path.isdir(REF):
    sys.exit('reference checkout not mounted: nothing to do')
  main()
