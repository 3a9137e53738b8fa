"""CPU: the Parquet packed path against batches produced by the reference's own loader.

tests/golden/reference_parquet_batches.json was written by running input/load_parquet.py `load_data_proc`
UNMODIFIED (it has no TensorFlow dependency) plus `ParquetInput._to_fea_dict` on the two small files of
tests/golden/parquet_case.py.  Checked here, bit exact:
  * batch boundaries and order (per-file full batches first, tails carried across files, partial batch last);
  * the CSR form: lens int32 feature-major [n_feat * B], vals int64 in the same order;
  * `vals % num_buckets` with floored mod on negative ids == the oracle's BUCKET_MOD rule (K1's restatement);
  * dense matrix [B, sum raw_dim] and labels."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import parquet_case as case  # noqa: E402

from easyrec_b200 import _lib, builder  # noqa: E402
from easyrec_b200.config import config_util  # noqa: E402
from easyrec_b200.input import readers  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_parquet_batches.json')))


def _reference_form(il, feats):
  """our batch -> the reference's packed (lens, vals): every sparse feature's B lens, feature-major."""
  B = il.batch_size
  lens, vals = [], []
  sparse = feats['sparse_fea'].numpy().reshape(len(il.sparse_names), B)
  for name in case.SPARSE:
    if name in il.sparse_names:
      lens.append(np.ones(B, np.int32))
      vals.append(sparse[il.sparse_names.index(name)])
    else:
      v, l, w = feats['tag_fea'][name]
      assert w is None
      lens.append(l.numpy())
      vals.append(v.numpy())
  return np.concatenate(lens), np.concatenate(vals)


def test_parquet_reader_reproduces_the_reference_loaders_batches(tmp_path):
  cfg = config_util.get_configs_from_pipeline_file(case.CONFIG)
  il, model, _ = builder.build_model(cfg, case.BATCH, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  # Parquet inputs bucket with `vals % num_buckets` (parquet_input.py:221), not the identity column's clamp
  assert all(il.features[n].bucket_mode == _lib.BUCKET_MOD for n in case.SPARSE)
  paths = case.write_files(str(tmp_path))
  mine = list(readers.ParquetInput(cfg, il, paths))
  full = [b for b in GOLD['batches'] if len(b['label']) == case.BATCH]
  assert len(mine) == len(full) == GOLD['n_batches_drop_remainder'] == 4
  assert len(GOLD['batches']) == 5 and len(GOLD['batches'][-1]['label']) == 1   # the partial batch we skip
  for (feats, labels), g in zip(mine, full):
    lens, vals = _reference_form(il, feats)
    assert lens.dtype == np.int32 and vals.dtype == np.int64
    assert np.array_equal(lens, np.array(g['lens'], np.int32))
    assert np.array_equal(vals, np.array(g['raw_vals'], np.int64))
    assert np.array_equal(feats['dense_fea'].numpy(), np.array(g['dense_fea'], np.float32))
    assert np.array_equal(labels.numpy(), np.array(g['label'], np.float32))
    # the bucket rule the device applies to these ids
    rows, _ = O.bucketize(vals, np.full(vals.size, _lib.BUCKET_MOD), np.full(vals.size, case.NUM_BUCKETS),
                          np.zeros(vals.size, np.int64))
    assert np.array_equal(rows, np.array(g['vals'], np.int64))
  assert min(min(b['raw_vals']) for b in full) < 0 and min(min(b['vals']) for b in full) >= 0
  assert any(0 in b['lens'] for b in full) and any(max(b['lens']) > 1 for b in full)


def test_reference_packed_dict_converts_to_the_input_layer_form(tmp_path):
  """a batch dict in the reference's own form feeds the InputLayer through readers.from_reference_packed."""
  cfg = config_util.get_configs_from_pipeline_file(case.CONFIG)
  il, model, _ = builder.build_model(cfg, case.BATCH, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  mine = list(readers.ParquetInput(cfg, il, case.write_files(str(tmp_path))))
  for (feats, _), g in zip(mine, GOLD['batches']):
    got = readers.from_reference_packed(il, {'sparse_fea': (np.array(g['raw_vals'], np.int64), np.array(g['lens'], np.int32)),
                                             'dense_fea': np.array(g['dense_fea'], np.float32)}, case.SPARSE)
    assert torch.equal(got['sparse_fea'], feats['sparse_fea']) and torch.equal(got['dense_fea'], feats['dense_fea'])
    for name, (v, l, w) in feats['tag_fea'].items():
      assert torch.equal(got['tag_fea'][name][0], v) and torch.equal(got['tag_fea'][name][1], l)
  with pytest.raises(ValueError):
    readers.from_reference_packed(il, {'sparse_fea': (np.zeros(3, np.int64), np.ones(5, np.int32))}, case.SPARSE)


def test_ragged_cells_on_an_id_feature_are_refused(tmp_path):
  import pyarrow as pa
  import pyarrow.parquet as pq
  cfg = config_util.get_configs_from_pipeline_file(case.CONFIG)
  il, model, _ = builder.build_model(cfg, case.BATCH, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  tab = pq.read_table(case.write_files(str(tmp_path))[0]).slice(0, case.BATCH)
  bad = tab.set_column(tab.schema.get_field_index('c_list1'), 'c_list1',
                       pa.array([[1], [], [2, 3], [4]], pa.list_(pa.int64())))
  pq.write_table(bad, str(tmp_path / 'bad.parquet'))
  with pytest.raises(ValueError, match='TagFeature'):
    list(readers.ParquetInput(cfg, il, str(tmp_path / 'bad.parquet')))


def test_golden_file_matches_its_generator_when_the_reference_is_mounted(tmp_path):
  if not os.path.isdir('/root/reference/easy_rec/python'):
    pytest.skip('reference checkout not mounted')
  import make_parquet_golden as gen
  keep, _ = gen.run(False)
  assert keep == GOLD['batches']
