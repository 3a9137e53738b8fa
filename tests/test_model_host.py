"""CPU: whole training steps of config-built models - InputLayer, backbone DAG, dense towers, loss, backward,
fused row update, flat dense optimizer, learning-rate schedule - with every kernel entry point replaced by a
double (sparse: the CPU oracle, as in test_input_layer_host.py; dense: plain torch / the oracle's numpy).  The
kernels are checked against the same references on the GPU; here the HOST wiring runs, including the backbone
paths no GPU test builds yet (3-D / pair outputs of the input_layer block, low-rank Cross, keras merge layers,
mixed id + tag groups)."""
import os

import numpy as np
import pytest
import torch

from easyrec_b200 import builder, kernels as K, trainer as T
from easyrec_b200.config import config_util
from easyrec_b200.input import readers
from oracle import oracle as O
import host_doubles
from test_input_layer_host import oracle_kernels  # noqa: F401  (fixture)

BN_EPS, BN_MOM = 1e-3, 0.99


@pytest.fixture
def dense_kernels(monkeypatch, oracle_kernels):  # noqa: F811
  host_doubles.install_dense(monkeypatch.setattr)


CFG = b'''
train_config { optimizer_config { adagrad_optimizer { learning_rate { exponential_decay_learning_rate {
  initial_learning_rate: 0.05 decay_steps: 5 decay_factor: 0.5 min_learning_rate: 0.01 } } } } }
data_config { batch_size: 16 input_type: CSVInput separator: "," label_fields: "label"
  input_fields { input_name: "label" input_type: FLOAT } input_fields { input_name: "uid" input_type: INT64 }
  input_fields { input_name: "item" input_type: STRING } input_fields { input_name: "tags" input_type: STRING }
  input_fields { input_name: "price" input_type: FLOAT } }
feature_config {
  features { input_names: "uid" feature_type: IdFeature embedding_dim: 8 hash_bucket_size: 40 }
  features { input_names: "item" feature_type: IdFeature embedding_dim: 8 hash_bucket_size: 30 }
  features { input_names: "tags" feature_type: TagFeature embedding_dim: 8 num_buckets: 12 separator: "|" combiner: "mean" }
  features { input_names: ["uid", "item"] feature_name: "uid_item" feature_type: ComboFeature embedding_dim: 8 hash_bucket_size: 64 }
  features { input_names: "price" feature_type: RawFeature embedding_dim: 8 boundaries: [1.0, 2.0, 5.0] } }
model_config { model_class: "RankModel"
  feature_groups { group_name: "all" feature_names: ["uid", "item", "tags", "uid_item", "price"] wide_deep: DEEP }
  backbone {
    blocks { name: "mlp" inputs { feature_group_name: "all" } keras_layer { class_name: "MLP" mlp { hidden_units: [16, 8] } } }
    blocks { name: "cube" inputs { feature_group_name: "all" } input_layer { only_output_3d_tensor: true } }
    blocks { name: "pooled" inputs { block_name: "cube" } lambda { expression: "lambda x: tf.reduce_sum(x, axis=1)" } }
    blocks { name: "pair" inputs { feature_group_name: "all" } input_layer { output_2d_tensor_and_feature_list: true } }
    blocks { name: "fm" inputs { block_name: "pair" input_slice: "[1]" } keras_layer { class_name: "FM" fm { use_variant: true } } }
    blocks { name: "cross" inputs { block_name: "pair" input_slice: "[0]" input_fn: "lambda x: [x, x]" }
             keras_layer { class_name: "Cross" st_params { fields { key: "projection_dim" value { number_value: 4 } } } } }
    blocks { name: "added" inputs { block_name: "pooled" } inputs { block_name: "fm" } merge_inputs_into_list: true
             keras_layer { class_name: "Add" } }
    concat_blocks: ["mlp", "added", "cross"]
    top_mlp { hidden_units: [8] }
  }
  model_params { l2_regularization: 1e-4 }
  embedding_regularization: 1e-5 }
'''


def test_backbone_model_trains_on_the_host_with_kernel_doubles(tmp_path, dense_kernels):
  cfg = config_util.get_configs_from_pipeline_file(CFG)
  il, model, opt = builder.build_model(cfg, 16, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  rng = np.random.default_rng(0)
  lines = []
  for i in range(16):
    uid, item = int(rng.integers(0, 6)), 'i%d' % rng.integers(0, 5)
    tags = '|'.join(str(v) for v in rng.integers(0, 12, rng.integers(0, 4)))
    label = int((uid + len(item)) % 2 == 0)
    lines.append('%d,%d,%s,%s,%.2f\n' % (label, uid, item, tags, rng.uniform(0, 8)))
  open(tmp_path / 't.csv', 'w').write(''.join(lines))
  (feats, labels), = list(readers.make_input(cfg, il, str(tmp_path / 't.csv')))
  tr = T.Trainer(model, il, 'adagrad', lr_fn=opt['lr_fn'])
  table0 = il.arenas[8].weight.clone()
  losses = [float(tr.train_step(feats, labels)[0]) for _ in range(12)]
  assert all(np.isfinite(losses)) and losses[-1] < losses[0] - 0.01, losses      # the batch is being fitted
  assert float(tr.dense_opt.lr_dev[0]) == pytest.approx(opt['lr_fn'](11)) == pytest.approx(0.0125)   # staircase decay reached step 11
  moved = (il.arenas[8].weight != table0).any(1)
  assert 5 < int(moved.sum()) < il.arenas[8].n_rows                               # only looked-up rows moved
  shapes = {n: tuple(p.shape) for n, p in model.named_parameters()}
  assert shapes['backbone.mods.cross.dense_u.kernel'] == (40, 4) and shapes['backbone.mods.cross.dense.kernel'] == (4, 40)
  assert not model.backbone.mods['cross'].dense_u.bias.requires_grad


@pytest.mark.parametrize('final', [True, False])
def test_deepfm_wiring_matches_the_oracle_on_the_host(final, dense_kernels):
  """DeepFM from config (with and without final_dnn) against oracle.deepfm_forward / the plain-head formula of
  model/deepfm.py:92-105 on the model's own weights (batch-norm in training mode)."""
  import sys
  import os
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  from test_config import MINI
  text = MINI if final else MINI.replace(b'final_dnn { hidden_units: [16] }', b'')
  cfg = config_util.get_configs_from_pipeline_file(text.replace(b'batch_size: 32', b'batch_size: 8'))
  il, model, _ = builder.build_model(cfg, 8, 'cpu', cpu_generator=torch.Generator().manual_seed(3))
  rng = np.random.default_rng(0)
  with torch.no_grad():
    for p in model.parameters():
      p.add_(torch.from_numpy(rng.normal(0, 0.05, tuple(p.shape)).astype(np.float32)))
  feats = {'sparse_fea': torch.from_numpy(rng.integers(0, 2**40, 8)), 'dense_fea': torch.from_numpy(rng.uniform(0, 10, (8, 1)).astype(np.float32))}
  model.train()
  logits = model(feats).detach().numpy()
  g = il.lookup(feats)
  wide, deep = g['wide'][0].detach().numpy(), g['deep'][0].detach().numpy()[:, :32]

  def layers(dnn):
    return [dict(W=l.kernel.detach().numpy(), b=l.bias.detach().numpy(), gamma=l.gamma.detach().numpy(),
                 beta=l.beta.detach().numpy()) for l in dnn.layers]
  if final:
    params = dict(dnn=layers(model.dnn), final=layers(model.final_dnn), out_W=model.output.kernel.detach().numpy(),
                  out_b=model.output.bias.detach().numpy())
    want, _ = O.deepfm_forward(wide[:, :2], deep, 2, 16, params)
  else:
    h, _ = O.dnn_forward(deep, layers(model.dnn))
    want = (wide[:, :2].sum(1, keepdims=True) + O.fm_fwd(deep, 2, 16).sum(1, keepdims=True) +
            h @ model.output.kernel.detach().numpy() + model.output.bias.detach().numpy())[:, 0]
  np.testing.assert_allclose(logits, want, rtol=1e-4, atol=1e-5)


def test_estimator_flow_on_the_host(tmp_path, dense_kernels):
  """EasyRecEstimator end to end on CPU with kernel doubles: train over a CSV file through the Prefetcher, evaluate
  (AUC + GAUC keyed by an input field), predict, save with the reference's part files, restore into a fresh
  estimator - same predictions."""
  from easyrec_b200.estimator import EasyRecEstimator
  text = ('''
model_dir: "%s"
train_config { num_steps: 240 log_step_count_steps: 100
  optimizer_config { adagrad_optimizer { learning_rate { constant_learning_rate { learning_rate: 0.1 } } } } }
eval_config { metrics_set { auc {} } metrics_set { gauc { uid_field: "grp" } } }
data_config { batch_size: 32 input_type: CSVInput separator: "," label_fields: "label"
  input_fields { input_name: "label" input_type: FLOAT } input_fields { input_name: "grp" input_type: INT64 }
  input_fields { input_name: "c" input_type: STRING } input_fields { input_name: "x" input_type: FLOAT } }
feature_config {
  features { input_names: "grp" feature_type: IdFeature embedding_dim: 8 num_buckets: 4 }
  features { input_names: "c" feature_type: IdFeature embedding_dim: 8 hash_bucket_size: 64 }
  features { input_names: "x" feature_type: RawFeature embedding_dim: 8 min_val: 0.0 max_val: 1.0 } }
model_config { model_class: "DeepFM"
  feature_groups { group_name: "deep" feature_names: ["grp", "c", "x"] wide_deep: DEEP }
  feature_groups { group_name: "wide" feature_names: ["grp", "c"] wide_deep: WIDE }
  deepfm { dnn { hidden_units: [16] } final_dnn { hidden_units: [8] } l2_regularization: 1e-6 } }
''' % str(tmp_path / 'm')).encode()
  rng = np.random.default_rng(4)
  with open(tmp_path / 'd.csv', 'w') as f:
    for _ in range(32 * 12):
      c = int(rng.integers(0, 10))
      f.write('%d,%d,tok%d,%.3f\n' % (int(rng.uniform() < (0.85 if c % 2 else 0.15)), rng.integers(0, 4), c, rng.uniform()))
  est = EasyRecEstimator(text, device='cpu', seed=3)
  make = lambda e: (lambda: readers.make_input(e._pipeline_config, e.input_layer, str(tmp_path / 'd.csv')))  # noqa: E731
  loss = est.train(make(est))
  assert est.global_step == 240 and np.isfinite(loss)          # 20 passes over the 12 batches
  ev = est.evaluate(make(est))
  assert ev['auc'] > 0.75 and 0.5 < ev['gauc'] <= 1.0 and ev["global_step"] == 240   # the label follows the token's parity
  first = next(iter(est.predict(make(est))))['probs']
  path = est.save(embedding_parts=True)
  assert sorted(p for p in os.listdir(path[:-3] + '-embedding') if 'c_embedding' in p and 'wide' not in p) == [
      'embed-input_layer__c_embedding__embedding_weights:0-part-0.bin',
      'embed-input_layer__c_embedding__embedding_weights__Adagrad:0-part-0.bin']
  fresh = EasyRecEstimator(text, device='cpu', seed=99)        # different initial weights
  assert not np.allclose(next(iter(fresh.predict(make(fresh))))['probs'], first)
  fresh.restore(path)
  np.testing.assert_allclose(next(iter(fresh.predict(make(fresh))))['probs'], first, rtol=1e-6, atol=1e-7)
  assert fresh.global_step == 240


@pytest.fixture
def interaction_doubles(monkeypatch, dense_kernels):
  host_doubles.install_interactions(monkeypatch.setattr)


def _gpu_test_configs():
  import test_gpu_models as G
  names = ['DCN_CFG', 'DIN_CFG', 'MMOE_CFG', 'DSSM_CFG', 'DLRM_CFG', 'BACKBONE_DCN_CFG', 'BACKBONE_DLRM_CFG', 'BACKBONE_MTL_CFG',
           'BACKBONE_MATCH_CFG', 'BACKBONE_WIRING_CFG', 'BACKBONE_EMBLAYER_CFG']
  return [(n, getattr(G, n)) for n in names]


@pytest.mark.parametrize('name,text', _gpu_test_configs(), ids=[n for n, _ in _gpu_test_configs()])
def test_every_gpu_test_model_config_trains_on_the_host(name, text, interaction_doubles):
  """The model configs of tests/test_gpu_models.py, run here with kernel doubles: what the GPU job will execute
  at the end of the round must at least build, step and fit a batch on the host side of the same code."""
  cfg = config_util.get_configs_from_pipeline_file(text.encode())
  B = 32
  il, model, opt = builder.build_model(cfg, B, 'cpu', cpu_generator=torch.Generator().manual_seed(1), default_seq_len=20)
  n_labels = max(1, len(cfg.data_config.label_fields))
  feats, labels = readers.DummyInput(il, n_labels=n_labels, seed=5).batch()
  if 'sparse_fea' in feats:   # keep the random ids small so that identity slots do not all collapse onto row 0
    feats['sparse_fea'] = feats['sparse_fea'] % 37
  tr = T.Trainer(model, il, 'adagrad', lr_fn=opt['lr_fn'])
  losses = [float(tr.train_step(feats, labels)[0]) for _ in range(15)]
  assert all(np.isfinite(losses)) and losses[-1] < losses[0], (name, losses)


def test_din_from_csv_sequences_on_the_host(tmp_path, interaction_doubles):
  """MultiTowerDIN fed by the native CSV reader: `|`-separated history, truncated to max_seq_len, key / history
  tables of the sequence group, target attention; the attended vector of a row must ignore everything past its length."""
  cfg = config_util.get_configs_from_pipeline_file(b'''
train_config { optimizer_config { adagrad_optimizer { learning_rate { constant_learning_rate { learning_rate: 0.05 } } } } }
data_config { batch_size: 8 input_type: CSVInput separator: "," label_fields: "clk"
  input_fields { input_name: "clk" input_type: FLOAT } input_fields { input_name: "user_id" input_type: INT64 }
  input_fields { input_name: "item_id" input_type: INT64 } input_fields { input_name: "hist" input_type: STRING } }
feature_config {
  features { input_names: "user_id" feature_type: IdFeature embedding_dim: 8 hash_bucket_size: 50 }
  features { input_names: "item_id" feature_type: IdFeature embedding_dim: 8 hash_bucket_size: 100 }
  features { input_names: "hist" feature_type: SequenceFeature embedding_dim: 8 hash_bucket_size: 100 max_seq_len: 5 separator: "|" } }
model_config { model_class: "MultiTowerDIN"
  feature_groups { group_name: "user" feature_names: ["user_id"] wide_deep: DEEP }
  feature_groups { group_name: "item" feature_names: ["item_id"] wide_deep: DEEP }
  seq_att_groups { group_name: "din" seq_att_map { key: "item_id" hist_seq: "hist" } }
  multi_tower { towers { input: "user" dnn { hidden_units: [16] } } towers { input: "item" dnn { hidden_units: [16] } }
                din_towers { input: "din" dnn { hidden_units: [16, 1] } } final_dnn { hidden_units: [16] } } }
''')
  il, model, opt = builder.build_model(cfg, 8, 'cpu', cpu_generator=torch.Generator().manual_seed(2))
  rows = ['1,1,10,10|11|12', '0,2,20,', '1,3,30,1|2|3|4|5|6|7', '0,4,40,40', '1,5,50,9|9', '0,6,60,60|61', '1,7,70,5', '0,8,80,1|2|3']
  open(tmp_path / 's.csv', 'w').write('\n'.join(rows) + '\n')
  (feats, labels), = list(readers.make_input(cfg, il, str(tmp_path / 's.csv')))
  ids, lens = feats['seq_fea']['hist']
  # a STRING field with a hash_bucket_size: the tokens are bucketed by the reader (Fingerprint64 % 100)
  assert lens.tolist() == [3, 0, 5, 1, 2, 2, 1, 3] and ids[1].tolist() == [0] * 5
  assert ids[2].tolist() == [O.fingerprint64(t) % 100 for t in '12345']            # the FIRST max_seq_len steps
  il.lookup(feats)
  so = il.seq_outputs['din']
  assert tuple(so['hist_seq_emb'].shape) == (8, 5, 8) and tuple(so['key'].shape) == (8, 8)
  assert not so['hist_seq_emb'][1].any() and not so['hist_seq_emb'][0, 3:].any()     # padded steps are zero vectors
  tr = T.Trainer(model, il, 'adagrad', lr_fn=opt['lr_fn'])
  losses = [float(tr.train_step(feats, labels)[0]) for _ in range(15)]
  assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_deepfm_trains_from_the_reference_loaders_parquet_case_on_the_host(tmp_path, dense_kernels):
  """the Parquet data set behind tests/golden/reference_parquet_batches.json (floored-mod buckets, ragged tag lists,
  carry-over across files) through ParquetInput -> InputLayer -> DeepFM for a few optimizer steps."""
  sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
  import sys
  sys.path.insert(0, sys_path)
  import parquet_case as case
  cfg = config_util.get_configs_from_pipeline_file(case.CONFIG)
  il, model, opt = builder.build_model(cfg, case.BATCH, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  batches = list(readers.make_input(cfg, il, case.write_files(str(tmp_path))))
  assert len(batches) == 4
  tr = T.Trainer(model, il, 'adagrad', lr=0.05)
  first = [float(tr.train_step(f, l)[0]) for f, l in batches]
  for _ in range(6):
    last = [float(tr.train_step(f, l)[0]) for f, l in batches]
  assert all(np.isfinite(first + last)) and sum(last) < sum(first)


def test_adam_step_hyper_parameters_follow_adam_s():
  """compat/adam_s.py:117-129,185-192,236-245: at the t-th apply (t = 1, 2, ...) the powers are beta^t and
  lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t); the trainer's step counter starts at 0."""
  import sys
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  from test_config import MINI
  cfg = config_util.get_configs_from_pipeline_file(MINI.replace(b'adam_optimizer', b'lazy_adam_optimizer'))
  il, model, opt = builder.build_model(cfg, 8, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  assert opt['kind'] == 'lazy_adam_optimizer' and next(iter(il.arenas.values())).opt_kind == T._lib.OPT_LAZY_ADAM
  tr = T.Trainer(model, il, 'lazy_adam', lr_fn=opt['lr_fn'], beta1=opt['beta1'], beta2=opt['beta2'])
  for step in (0, 1, 7):
    tr.step = step
    tr._set_hyper()
    o = il.opt_holder['opt']
    t = step + 1
    assert o.kind == T._lib.OPT_LAZY_ADAM and o.lr == pytest.approx(0.001)
    assert o.beta1_power == pytest.approx(0.9**t) and o.beta2_power == pytest.approx(0.999**t)
    # (the config's beta2 is the fp32 proto value 0.99900001..., as it is in the TF graph: 1e-5 relative slack)
    assert float(tr.dense_opt.lr_dev[0]) == pytest.approx(0.001 * (1 - 0.999**t)**0.5 / (1 - 0.9**t), rel=2e-5)
