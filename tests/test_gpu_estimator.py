"""GPU: the reference's entry point end to end -- EasyRecEstimator(pipeline_config).train / evaluate / predict
(model/easy_rec_estimator.py:62-153, main.py:296-400) over a CSV file and over the same rows as Parquet.
The label depends on one id feature, so a few hundred fused steps must lift AUC well above chance; both input
formats must produce the same model (same batches -> same deterministic step)."""
import numpy as np
import pytest
import torch

from easyrec_b200.config import config_util
from easyrec_b200.estimator import EasyRecEstimator
from easyrec_b200.input import readers

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

CFG = '''
model_dir: "%(dir)s"
train_config { num_steps: 120 log_step_count_steps: 50
  optimizer_config { adagrad_optimizer { learning_rate { constant_learning_rate { learning_rate: 0.05 } } } } }
data_config { batch_size: 256 input_type: %(kind)s separator: "\\t" label_fields: "label"
  input_fields { input_name: "label" input_type: FLOAT }
  input_fields { input_name: "F1" input_type: FLOAT }
  input_fields { input_name: "C1" input_type: INT64 }
  input_fields { input_name: "C2" input_type: INT64 } }
feature_config {
  features { input_names: "F1" feature_type: RawFeature embedding_dim: 16 min_val: 0.0 max_val: 10.0 }
  features { input_names: "C1" feature_type: IdFeature embedding_dim: 16 hash_bucket_size: 2000 }
  features { input_names: "C2" feature_type: IdFeature embedding_dim: 16 num_buckets: 50 }
}
model_config { model_class: "DeepFM"
  feature_groups { group_name: "deep" feature_names: ["F1", "C1", "C2"] wide_deep: DEEP }
  feature_groups { group_name: "wide" feature_names: ["F1", "C1", "C2"] wide_deep: WIDE }
  deepfm { dnn { hidden_units: [64, 32] } final_dnn { hidden_units: [32] } l2_regularization: 1e-6 }
  embedding_regularization: 1e-6 }
'''


def _write(tmp_path, n=256 * 125):
  import pyarrow as pa
  import pyarrow.parquet as pq
  rng = np.random.default_rng(1)
  c1 = rng.integers(0, 300, n).astype(np.int64) * 7919
  c2 = rng.integers(0, 50, n).astype(np.int64)
  f1 = rng.uniform(0, 10, n).astype(np.float32)
  p = 1 / (1 + np.exp(-(((c1 // 7919) % 7 - 3) * 0.9 + (c2 % 3 - 1) * 0.5)))
  lab = (rng.uniform(size=n) < p).astype(np.float32)
  with open(tmp_path / 'train.csv', 'w') as f:
    for i in range(n):
      f.write('%g\t%r\t%d\t%d\n' % (lab[i], float(f1[i]), c1[i], c2[i]))
  pq.write_table(pa.table({'label': lab, 'F1': f1, 'C1': c1, 'C2': pa.array([[int(v)] for v in c2], pa.list_(pa.int64()))}),
                 str(tmp_path / 'train.parquet'), row_group_size=1000)
  return lab


def _run(tmp_path, kind, path):
  cfg = config_util.get_configs_from_pipeline_file((CFG % dict(dir=str(tmp_path / 'm'), kind=kind)).encode())
  est = EasyRecEstimator(cfg, device=DEV, seed=7)
  loss = est.train(lambda: readers.make_input(cfg, est.input_layer, path))
  assert est.global_step == 120 and np.isfinite(loss)
  ev = est.evaluate(lambda: readers.make_input(cfg, est.input_layer, path), steps=20)
  preds = next(iter(est.predict(lambda: readers.make_input(cfg, est.input_layer, path))))
  assert preds['probs'].shape == (256,) and np.all((preds['probs'] >= 0) & (preds['probs'] <= 1))
  ckpt = est.save()
  assert ckpt.endswith('model.ckpt-120.pt')
  # tables + Adagrad slots through the reference's part-file layout and back (native re-shard loader)
  arena = est.input_layer.arenas[16]
  w, acc = arena.weight.clone(), arena.state0.clone()
  assert est.save(embedding_parts=True) == ckpt
  arena.storage.zero_()
  est.restore(ckpt)
  assert torch.equal(arena.weight, w) and torch.equal(arena.state0, acc) and est.global_step == 120
  return ev['auc'], {k: v.detach().clone() for k, v in est.model.state_dict().items()}, est.input_layer.arenas[16].weight.clone()


def test_estimator_trains_from_csv_and_parquet_to_the_same_model(tmp_path):
  _write(tmp_path)
  auc_csv, sd_csv, t_csv = _run(tmp_path, 'CSVInput', str(tmp_path / 'train.csv'))
  auc_pq, sd_pq, t_pq = _run(tmp_path, 'ParquetInput', str(tmp_path / 'train.parquet'))
  assert auc_csv > 0.62, auc_csv          # the label is learnable from C1 / C2 (chance = 0.5)
  assert abs(auc_csv - auc_pq) < 1e-6     # same batches, deterministic step -> same model
  assert torch.equal(t_csv, t_pq)
  for k in sd_csv:
    assert torch.equal(sd_csv[k], sd_pq[k]), k
