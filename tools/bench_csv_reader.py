"""Host-side reader throughput on a Criteo-shaped TSV (1 label, 13 floats, 26 int64 ids): native er_csv_parse
engine vs the pure-python restatement.  CPU only.   python tools/bench_csv_reader.py [rows] [threads]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyrec_b200 import builder  # noqa: E402
from easyrec_b200.config import config_util  # noqa: E402
from easyrec_b200.input import readers  # noqa: E402

B = 8192


def config():
  fields = ['input_fields { input_name: "label" input_type: FLOAT }']
  feats, names = [], []
  for i in range(1, 14):
    fields.append('input_fields { input_name: "f%d" input_type: FLOAT }' % i)
    feats.append('features { input_names: "f%d" feature_type: RawFeature embedding_dim: 16 }' % i)
    names.append('f%d' % i)
  for i in range(1, 27):
    fields.append('input_fields { input_name: "c%d" input_type: INT64 }' % i)
    feats.append('features { input_names: "c%d" feature_type: IdFeature embedding_dim: 16 hash_bucket_size: 1000000 '
                 'embedding_name: "embedding" }' % i)
    names.append('c%d' % i)
  fn = ' '.join('feature_names: "%s"' % n for n in names)
  return ('data_config { batch_size: %d input_type: CSVInput separator: "\\t" label_fields: "label" %s }\n'
          'feature_config { %s }\n'
          'model_config { model_class: "DeepFM" feature_groups { group_name: "deep" %s wide_deep: DEEP } '
          'feature_groups { group_name: "wide" %s wide_deep: WIDE } deepfm { dnn { hidden_units: [16] } '
          'final_dnn { hidden_units: [8] } } }' % (B, ' '.join(fields), ' '.join(feats), fn, fn)).encode()


def main():
  rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20 * B
  threads = int(sys.argv[2]) if len(sys.argv) > 2 else min(8, os.cpu_count())
  cfg = config_util.get_configs_from_pipeline_file(config())
  il, _, _ = builder.build_model(cfg, B, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  rng = np.random.default_rng(0)
  with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, 'criteo.tsv')
    lab = rng.integers(0, 2, rows)
    dense = rng.lognormal(size=(rows, 13)).round(4)
    ids = rng.integers(0, 2**40, (rows, 26))
    with open(path, 'w') as f:
      for i in range(rows):
        f.write('%d\t%s\t%s\n' % (lab[i], '\t'.join('%g' % v for v in dense[i]), '\t'.join(map(str, ids[i]))))
    size = os.path.getsize(path)
    for engine, n in (('native', rows), ('python', min(rows, 4 * B))):
      t0 = time.perf_counter()
      got = 0
      for feats, labels in readers.CSVInput(cfg, il, path, engine=engine, n_threads=threads):
        got += labels.shape[0]
        if got >= n:
          break
      dt = time.perf_counter() - t0
      print('%-7s %9d rows in %6.2f s  -> %10.0f rows/s  (%.0f MB/s of text, %d threads)' %
            (engine, got, dt, got / dt, size / rows * got / dt / 1e6, threads if engine == 'native' else 1))


if __name__ == '__main__':
  main()
