"""Write the C2 synthetic workload as files: TSV with the column order of
examples/configs/deepfm_on_criteo.config (label, f1..f13, c1..c26) and Parquet with the schema of
tools/criteo/convert_data.py (is_click, f1..f13 float32, c1..c26 int64).  Same rows as bench.py's batches.

  python tools/make_synthetic.py --rows 81920 --out /tmp/criteo_c2        -> /tmp/criteo_c2.tsv, /tmp/criteo_c2.parquet
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyrec_b200 import workloads  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--rows', type=int, default=81920)
  ap.add_argument('--batch', type=int, default=8192)
  ap.add_argument('--seed', type=int, default=20240)
  ap.add_argument('--out', default='criteo_c2')
  ap.add_argument('--uniform-ids', action='store_true')
  a = ap.parse_args()
  n_b = (a.rows + a.batch - 1) // a.batch
  ids, dense, labels = [], [], []
  for i in range(n_b):
    x, d, l = workloads.criteo_batch(a.batch, a.seed + i, uniform=a.uniform_ids)
    ids.append(x.reshape(26, a.batch).T)   # feature-major -> [B, 26]
    dense.append(d)
    labels.append(l)
  ids = np.concatenate(ids)[:a.rows]
  dense = np.concatenate(dense)[:a.rows]
  labels = np.concatenate(labels)[:a.rows]
  with open(a.out + '.tsv', 'w') as f:
    for r in range(a.rows):
      f.write('%d\t%s\t%s\n' % (int(labels[r]), '\t'.join('%.6g' % v for v in dense[r]),
                                  '\t'.join(str(int(v)) for v in ids[r])))
  try:
    import pyarrow as pa
    import pyarrow.parquet as pq
    cols = {'is_click': labels.astype(np.int32)}
    for j in range(13):
      cols['f%d' % (j + 1)] = dense[:, j].astype(np.float32)
    for j in range(26):
      cols['c%d' % (j + 1)] = ids[:, j].astype(np.int64)
    pq.write_table(pa.table(cols), a.out + '.parquet', row_group_size=a.batch)
  except ImportError:
    print('pyarrow missing: parquet file skipped')
  print('wrote %d rows to %s.tsv / .parquet' % (a.rows, a.out))


if __name__ == '__main__':
  main()
