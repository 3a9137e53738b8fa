"""The small Parquet data set behind tests/golden/reference_parquet_batches.json: two files whose row counts are
not multiples of the batch size (so the loader's carry-over between files is exercised), scalar and list id
columns, negative ids (floored mod), empty lists, a dense column stored as 1-element lists (the reference loader reads x[0] of a list-valued dense cell,
load_parquet.py:108-114, so wider dense lists are not representable there), one label."""
import os

import numpy as np

BATCH = 4
NUM_BUCKETS = 1000
SPARSE = ['c_scalar', 'c_list1', 'c_tags']     # Id (scalar), Id (1-element lists), Tag (ragged lists)
DENSE = [('f_one', 1), ('f_list1', 1)]
LABEL = 'is_click'
ROWS = [10, 7]                                 # 10 = 2 full + 2 left, 7 = 1 full + 3 left -> 5 carried, 1 dropped/partial

CONFIG = ('''
data_config { batch_size: %d input_type: ParquetInput label_fields: "is_click"
  input_fields { input_name: "is_click" input_type: FLOAT }
  input_fields { input_name: "f_one" input_type: FLOAT } input_fields { input_name: "f_list1" input_type: FLOAT }
  input_fields { input_name: "c_scalar" input_type: INT64 } input_fields { input_name: "c_list1" input_type: INT64 }
  input_fields { input_name: "c_tags" input_type: INT64 } }
feature_config {
  features { input_names: "c_scalar" feature_type: IdFeature embedding_dim: 8 num_buckets: %d embedding_name: "e" }
  features { input_names: "c_list1" feature_type: IdFeature embedding_dim: 8 num_buckets: %d embedding_name: "e" }
  features { input_names: "c_tags" feature_type: TagFeature embedding_dim: 8 num_buckets: %d embedding_name: "e" combiner: "mean" }
  features { input_names: "f_one" feature_type: RawFeature embedding_dim: 8 }
  features { input_names: "f_list1" feature_type: RawFeature embedding_dim: 8 }
}
model_config { model_class: "DeepFM"
  feature_groups { group_name: "deep" feature_names: ["c_scalar", "c_list1", "c_tags", "f_one", "f_list1"] wide_deep: DEEP }
  feature_groups { group_name: "wide" feature_names: ["c_scalar", "c_list1", "c_tags"] wide_deep: WIDE }
  deepfm { dnn { hidden_units: [16] } final_dnn { hidden_units: [8] } } }
''' % (BATCH, NUM_BUCKETS, NUM_BUCKETS, NUM_BUCKETS)).encode()


def write_files(directory):
  """-> list of parquet paths (deterministic content)."""
  import pyarrow as pa
  import pyarrow.parquet as pq
  rng = np.random.default_rng(77)
  paths = []
  for k, n in enumerate(ROWS):
    scalar = rng.integers(-5000, 2**40, n).astype(np.int64)
    list1 = [[int(v)] for v in rng.integers(0, 2**33, n)]
    tags = [[int(v) for v in rng.integers(-2000, 2**35, int(m))] for m in rng.integers(0, 4, n)]
    tags[0] = []
    scalar[1], scalar[2] = -7, -1                 # floored mod: 993, 999 (and -1 is NOT a missing marker here)
    tags[1] = [-2001, 5, 5]
    tab = pa.table({
        LABEL: pa.array((rng.uniform(size=n) < 0.4).astype(np.float32)),
        'f_one': pa.array(rng.uniform(0, 5, n).astype(np.float32)),
        'f_list1': pa.array([[float(a)] for a in rng.normal(size=n).astype(np.float32)], pa.list_(pa.float32())),
        'c_scalar': pa.array(scalar),
        'c_list1': pa.array(list1, pa.list_(pa.int64())),
        'c_tags': pa.array(tags, pa.list_(pa.int64())),
    })
    path = os.path.join(directory, 'part-%d.parquet' % k)
    pq.write_table(tab, path, row_group_size=3)
    paths.append(path)
  return paths
