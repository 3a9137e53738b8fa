"""GPU: the BASELINE model families built from EasyRec-format pipeline configs (text_format, subset schema),
trained a few steps on the fused path; MultiTowerDIN's forward is additionally checked against a plain
PyTorch restatement of the reference graph on the same tables/weights (logits within 1e-4)."""
import numpy as np
import pytest
import torch

from easyrec_b200 import builder
from easyrec_b200.config import config_util
from easyrec_b200.trainer import Trainer
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
B = 256

HEAD = '''
model_dir: "/tmp/x"
train_config { optimizer_config { adagrad_optimizer { learning_rate { constant_learning_rate { learning_rate: 0.05 } } } } }
data_config { batch_size: 256 input_type: DummyInput label_fields: "clk" }
'''

FEATS = '''
feature_config {
  features { input_names: "user_id" feature_type: IdFeature embedding_dim: 16 hash_bucket_size: 1000 }
  features { input_names: "age" feature_type: IdFeature embedding_dim: 16 num_buckets: 10 }
  features { input_names: "item_id" feature_type: IdFeature embedding_dim: 16 hash_bucket_size: 5000 }
  features { input_names: "cate" feature_type: IdFeature embedding_dim: 16 hash_bucket_size: 200 }
  features { input_names: "price" feature_type: RawFeature embedding_dim: 16 min_val: 0 max_val: 100 }
  features { input_names: "hist_items" feature_type: SequenceFeature embedding_dim: 16 hash_bucket_size: 5000 max_seq_len: 20 }
}
'''

DCN_CFG = HEAD + FEATS + '''
model_config { model_class: "DCN"
  feature_groups { group_name: "all" feature_names: ["user_id", "age", "item_id", "cate", "price"] wide_deep: DEEP }
  dcn { deep_tower { input: "all" dnn { hidden_units: [64, 32] } } cross_tower { input: "all" cross_num: 3 }
        final_dnn { hidden_units: [32, 16] } l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

DIN_CFG = HEAD + FEATS + '''
model_config { model_class: "MultiTowerDIN"
  feature_groups { group_name: "user" feature_names: ["user_id", "age"] wide_deep: DEEP }
  feature_groups { group_name: "item" feature_names: ["item_id", "cate", "price"] wide_deep: DEEP }
  seq_att_groups { group_name: "din" seq_att_map { key: "item_id" hist_seq: "hist_items" } }
  multi_tower { towers { input: "user" dnn { hidden_units: [32, 16] } } towers { input: "item" dnn { hidden_units: [32, 16] } }
                din_towers { input: "din" dnn { hidden_units: [32, 16, 1] } } final_dnn { hidden_units: [32, 16] }
                l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

DLRM_CFG = HEAD + FEATS.replace('features { input_names: "price" feature_type: RawFeature embedding_dim: 16 min_val: 0 max_val: 100 }',
                                'features { input_names: "price" feature_type: RawFeature min_val: 0 max_val: 100 }') + '''
model_config { model_class: "DLRM"
  feature_groups { group_name: "sparse" feature_names: ["user_id", "age", "item_id", "cate"] wide_deep: DEEP }
  feature_groups { group_name: "dense" feature_names: ["price"] wide_deep: DEEP }
  dlrm { bot_dnn { hidden_units: [32, 16] } top_dnn { hidden_units: [64, 32] } arch_with_dense_feature: true l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

BACKBONE_DCN_CFG = HEAD + FEATS + '''
model_config { model_class: "RankModel"
  feature_groups { group_name: "all" feature_names: ["user_id", "age", "item_id", "cate", "price"] wide_deep: DEEP }
  backbone {
    blocks { name: "deep" inputs { feature_group_name: "all" } keras_layer { class_name: "MLP" mlp { hidden_units: [64, 32] } } }
    blocks { name: "cross" inputs { feature_group_name: "all" input_fn: "lambda x: [x, x]" }
             recurrent { num_steps: 3 fixed_input_index: 0 keras_layer { class_name: "Cross" } } }
    concat_blocks: ["deep", "cross"]
    top_mlp { hidden_units: [32, 16] }
  }
  model_params { l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

# a23: a backbone `embedding_layer` block (one Embedding(sum vocab, 12) over the bucketized ids of group "ids",
# layers/keras/embedding.py + InputLayer.get_bucketized_features) next to an ordinary input_layer group
BACKBONE_EMBLAYER_CFG = HEAD + FEATS + '''
model_config { model_class: "RankModel"
  feature_groups { group_name: "ids" feature_names: ["user_id", "age", "item_id", "cate"] wide_deep: DEEP }
  feature_groups { group_name: "dense" feature_names: ["price"] wide_deep: DEEP }
  backbone {
    blocks { name: "emb" inputs { feature_group_name: "ids" } embedding_layer { embedding_dim: 12 } }
    blocks { name: "raw" inputs { feature_group_name: "dense" } input_layer { } }
    blocks { name: "mlp" inputs { block_name: "emb" } inputs { block_name: "raw" }
             keras_layer { class_name: "MLP" mlp { hidden_units: [64, 32] } } }
    concat_blocks: ["mlp"]
  }
  model_params { l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

BACKBONE_DLRM_CFG = HEAD + FEATS + '''
model_config { model_class: "RankModel"
  feature_groups { group_name: "sparse" feature_names: ["user_id", "age", "item_id", "cate"] wide_deep: DEEP }
  feature_groups { group_name: "dense" feature_names: ["price"] wide_deep: DEEP }
  backbone {
    blocks { name: "bottom" inputs { feature_group_name: "dense" } keras_layer { class_name: "MLP" mlp { hidden_units: [32, 16] } } }
    blocks { name: "sparse" inputs { feature_group_name: "sparse" } input_layer { only_output_feature_list: true } }
    blocks { name: "dot" inputs { block_name: "bottom" input_fn: "lambda x: [x]" } inputs { block_name: "sparse" }
             keras_layer { class_name: "DotInteraction" } }
    blocks { name: "top" inputs { block_name: "bottom" } inputs { block_name: "dot" } input_concat_axis: 1
             keras_layer { class_name: "MLP" mlp { hidden_units: [32, 16] } } }
    concat_blocks: ["top"]
  }
  model_params { l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

BACKBONE_MTL_CFG = HEAD.replace('label_fields: "clk"', 'label_fields: "clk" label_fields: "buy"') + FEATS + '''
model_config { model_class: "MultiTaskModel"
  feature_groups { group_name: "all" feature_names: ["user_id", "age", "item_id", "cate", "price"] wide_deep: DEEP }
  backbone {
    blocks { name: "all" inputs { feature_group_name: "all" } input_layer { only_output_feature_list: true } }
    blocks { name: "senet" inputs { block_name: "all" } keras_layer { class_name: "SENet" senet { reduction_ratio: 4 } } }
    blocks { name: "mmoe" inputs { block_name: "senet" }
             keras_layer { class_name: "MMoE" mmoe { num_task: 2 num_expert: 3 expert_mlp { hidden_units: [32, 16] } } } }
  }
  model_params {
    task_towers { tower_name: "ctr" label_name: "clk" dnn { hidden_units: [16, 8] } weight: 1.0 }
    task_towers { tower_name: "cvr" label_name: "buy" dnn { hidden_units: [16, 8] } weight: 0.5 }
    l2_regularization: 1e-6 }
  embedding_regularization: 1e-5 }
'''

BACKBONE_MATCH_CFG = HEAD + FEATS + '''
model_config { model_class: "MatchModel"
  feature_groups { group_name: "user" feature_names: ["user_id", "age"] wide_deep: DEEP }
  feature_groups { group_name: "item" feature_names: ["item_id", "cate", "price"] wide_deep: DEEP }
  backbone {
    blocks { name: "user" inputs { feature_group_name: "user" } input_layer { } }
    blocks { name: "item" inputs { feature_group_name: "item" } input_layer { } }
    blocks { name: "user_tower" inputs { block_name: "user" }
             keras_layer { class_name: "MLP" mlp { hidden_units: [32, 16] use_final_bn: false final_activation: "linear" } } }
    blocks { name: "item_tower" inputs { block_name: "item" }
             keras_layer { class_name: "MLP" mlp { hidden_units: [32, 16] use_final_bn: false final_activation: "linear" } } }
    output_blocks: ["user_tower", "item_tower"]
  }
  model_params { l2_regularization: 1e-6 temperature: 0.05 }
  loss_type: SOFTMAX_CROSS_ENTROPY
  embedding_regularization: 1e-5 }
'''

BACKBONE_WIRING_CFG = HEAD + FEATS + '''
model_config { model_class: "RankModel"
  feature_groups { group_name: "all" feature_names: ["user_id", "age", "item_id", "cate", "price"] wide_deep: DEEP }
  backbone {
    blocks { name: "feats" inputs { feature_group_name: "all" } input_layer { only_output_feature_list: true } }
    blocks { name: "parts" inputs { feature_group_name: "all" }
             repeat { num_repeat: 2 input_fn: "lambda x, i: x[:, i * 40:(i + 1) * 40]" output_concat_axis: 1
                      keras_layer { class_name: "MLP" mlp { hidden_units: [16] } } } }
    blocks { name: "scaled" inputs { block_name: "parts" input_slice: "[:, :16]" } lambda { expression: "lambda x: x * 2.0" } }
    blocks { name: "fm" inputs { block_name: "feats" } keras_layer { class_name: "FM" fm { use_variant: true } } }
    concat_blocks: ["parts", "scaled", "fm"]
    top_mlp { hidden_units: [16] }
  }
  model_params { l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

MMOE_CFG = HEAD.replace('label_fields: "clk"', 'label_fields: "clk" label_fields: "buy"') + FEATS + '''
model_config { model_class: "MMoE"
  feature_groups { group_name: "all" feature_names: ["user_id", "age", "item_id", "cate", "price"] wide_deep: DEEP }
  mmoe { expert_dnn { hidden_units: [32, 16] } num_expert: 4
         task_towers { tower_name: "ctr" label_name: "clk" dnn { hidden_units: [16, 8] } loss_type: CLASSIFICATION weight: 1.0 }
         task_towers { tower_name: "cvr" label_name: "buy" dnn { hidden_units: [16, 8] } loss_type: CLASSIFICATION weight: 0.5 }
         l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''

DSSM_CFG = HEAD + FEATS + '''
model_config { model_class: "DSSM"
  feature_groups { group_name: "user" feature_names: ["user_id", "age"] wide_deep: DEEP }
  feature_groups { group_name: "item" feature_names: ["item_id", "cate", "price"] wide_deep: DEEP }
  dssm { user_tower { id: "user_id" dnn { hidden_units: [64, 32, 16] } } item_tower { id: "item_id" dnn { hidden_units: [64, 32, 16] } }
         simi_func: COSINE temperature: 0.1 scale_simi: true l2_regularization: 1e-5 }
  loss_type: SOFTMAX_CROSS_ENTROPY embedding_regularization: 1e-5 }
'''


def make_batch(seed, n_task=1):
  rng = np.random.default_rng(seed)
  ids = np.stack([rng.integers(0, 10**6, B), rng.integers(0, 10, B), rng.integers(0, 10**6, B),
                  rng.integers(0, 500, B)]).astype(np.int64)  # feature-major: user_id, age, item_id, cate
  dense = rng.uniform(0, 100, (B, 1)).astype(np.float32)
  T = 20
  hist = rng.integers(0, 10**6, (B, T)).astype(np.int64)
  lens = rng.integers(0, T + 1, B).astype(np.int32)
  labels = (rng.uniform(size=(B, n_task)) < 0.3).astype(np.float32)
  feats = {'sparse_fea': torch.from_numpy(ids.reshape(-1)).to(DEV), 'dense_fea': torch.from_numpy(dense).to(DEV),
           'seq_fea': {'hist_items': (torch.from_numpy(hist).to(DEV), torch.from_numpy(lens).to(DEV))},
           'item_ids': torch.from_numpy(ids[2]).to(DEV)}
  lab = torch.from_numpy(labels if n_task > 1 else labels[:, 0]).to(DEV)
  return feats, lab, (ids, dense, hist, lens)


@pytest.mark.parametrize('cfg_text,n_task', [(DCN_CFG, 1), (DIN_CFG, 1), (MMOE_CFG, 2), (DSSM_CFG, 1), (DLRM_CFG, 1),
                                             (BACKBONE_DCN_CFG, 1), (BACKBONE_DLRM_CFG, 1), (BACKBONE_MTL_CFG, 2),
                                             (BACKBONE_MATCH_CFG, 1), (BACKBONE_WIRING_CFG, 1), (BACKBONE_EMBLAYER_CFG, 1)])
def test_models_from_pipeline_config_train(cfg_text, n_task):
  torch.backends.cuda.matmul.allow_tf32 = False
  cfg = config_util.get_configs_from_pipeline_file(cfg_text.encode())
  il, model, opt = builder.build_model(cfg, B, DEV, generator=torch.Generator(device=DEV).manual_seed(1),
                                       cpu_generator=torch.Generator().manual_seed(1), default_seq_len=20)
  tr = Trainer(model, il, 'adagrad', lr_fn=opt['lr_fn'])
  feats, lab, _ = make_batch(3, n_task)
  losses = [float(tr.train_step(feats, lab)[0]) for _ in range(25)]
  assert all(np.isfinite(losses))
  assert losses[-1] < losses[0] - 0.02, losses  # the same batch is being fitted
  # a different batch still runs through the same static plan
  feats2, lab2, _ = make_batch(4, n_task)
  assert np.isfinite(float(tr.train_step(feats2, lab2)[0]))


@pytest.mark.parametrize('cfg_text', [DSSM_CFG, MMOE_CFG])
def test_step_replayed_from_a_cuda_graph_equals_the_eager_step(cfg_text):
  """Models whose graphs hold plain torch ops on parameters (DSSM's similarity scale, the multi-task loss sum) capture
  too: the eager steps ahead of the capture run on a side stream, as torch asks for whole-step capture."""
  torch.backends.cuda.matmul.allow_tf32 = False
  n_task = 2 if cfg_text is MMOE_CFG else 1
  cfg = config_util.get_configs_from_pipeline_file(cfg_text.encode())
  losses = []
  for graph in (False, True):
    il, model, opt = builder.build_model(cfg, B, DEV, generator=torch.Generator(device=DEV).manual_seed(1),
                                         cpu_generator=torch.Generator().manual_seed(1), default_seq_len=20)
    tr = Trainer(model, il, 'adagrad', lr_fn=opt['lr_fn'], use_cuda_graph=graph)
    out = []
    for k in range(6):
      feats, lab, _ = make_batch(10 + k, n_task)
      out.append(float(tr.train_step(feats, lab)[0]))
    losses.append(out)
    assert (not graph) or tr._graph is not None
  np.testing.assert_allclose(losses[1], losses[0], rtol=0, atol=1e-6)


def test_din_forward_matches_plain_torch_restatement():
  torch.backends.cuda.matmul.allow_tf32 = False
  cfg = config_util.get_configs_from_pipeline_file(DIN_CFG.encode())
  il, model, _ = builder.build_model(cfg, B, DEV, generator=torch.Generator(device=DEV).manual_seed(1),
                                     cpu_generator=torch.Generator().manual_seed(1), default_seq_len=20)
  model.train()
  feats, lab, (ids, dense, hist, lens) = make_batch(5)
  logits = model(feats).detach()
  arena = il.arenas[16]
  W = arena.weight.detach()

  def table(name):
    off, n, _ = arena.tables[name]
    return W[off:off + n]

  def hashed(v, nb):
    rows, _ = O.bucketize(v.reshape(-1), 0, nb, 0)
    return torch.from_numpy(rows.reshape(v.shape)).to(DEV)

  def bn(x):
    mu, var = x.mean(0), ((x - x.mean(0))**2).mean(0)
    return (x - mu) / torch.sqrt(var + 1e-3)

  def dnn(mod, x, last_plain=False):
    n = len(mod.layers)
    for i, lay in enumerate(mod.layers):
      x = x @ lay.kernel + lay.bias
      if lay.use_bn:
        x = bn(x) * lay.gamma + lay.beta
      if lay.relu:
        x = torch.relu(x)
    return x

  age = torch.from_numpy(np.where((ids[1] < 0) | (ids[1] >= 10), 0, ids[1])).to(DEV)
  user = torch.cat([table('user_id_embedding')[hashed(ids[0], 1000)], table('age_embedding')[age]], 1)
  pn = torch.from_numpy(dense / 100.0).to(DEV)
  item = torch.cat([table('item_id_embedding')[hashed(ids[2], 5000)], table('cate_embedding')[hashed(ids[3], 200)],
                    pn * table('price_embedding')[0][None, :]], 1)
  key = table('din/item_id_embedding')[hashed(ids[2], 5000)]
  he = table('din/hist_items_embedding')[hashed(hist, 5000)]
  T = hist.shape[1]
  mask = torch.arange(T, device=DEV)[None, :] < torch.from_numpy(lens).to(DEV)[:, None]
  he = he * mask[:, :, None]
  cur = key[:, None, :].expand(-1, T, -1)
  din_in = torch.cat([cur, he, cur - he, cur * he], -1).reshape(B * T, -1)
  scores = dnn(model.din_dnn[0], din_in).reshape(B, 1, T)
  scores = torch.where(mask[:, None, :], scores, torch.full_like(scores, -2.0**32 + 1))
  att = (torch.softmax(scores, -1) @ he).reshape(B, -1)
  feas = [dnn(model.tower_dnn[0], bn(user) * model.tower_bn[0].gamma + model.tower_bn[0].beta),
          dnn(model.tower_dnn[1], bn(item) * model.tower_bn[1].gamma + model.tower_bn[1].beta),
          torch.cat([att, key], 1)]
  ref = (dnn(model.final_dnn, torch.cat(feas, 1)) @ model.output.kernel + model.output.bias)[:, 0]
  assert float((logits - ref).abs().max()) < 1e-4


def test_backbone_layers_match_reference_formulas():
  """keras Cross (DCN v2) and DotInteraction of the backbone against their TF formulas in float64."""
  from easyrec_b200 import backbone as BB
  g = torch.Generator(device=DEV).manual_seed(2)
  x0 = torch.randn(300, 48, device=DEV, generator=g)
  x = torch.randn(300, 48, device=DEV, generator=g)
  cross = BB.Cross(48, {'diag_scale': 0.1}, torch.Generator().manual_seed(0)).to(DEV)
  with torch.no_grad():
    cross.dense.bias.copy_(torch.randn(48, device=DEV, generator=g) * 0.1)
  got = cross([x0, x])
  W, b = cross.dense.kernel.double(), cross.dense.bias.double()
  want = x0.double() * (x.double() @ W + b + 0.1 * x.double()) + x.double()
  assert float((got.double() - want).abs().max()) < 1e-5
  feats = [torch.randn(64, 16, device=DEV, generator=g) for _ in range(5)]
  for self_int in (False, True):
    got = BB.DotInteraction({'self_interaction': self_int})(feats)
    f = torch.stack(feats, 1).double()
    xa = f @ f.transpose(1, 2)
    pairs = [(i, j) for i in range(5) for j in range(5) if (j <= i if self_int else j < i)]
    want = torch.stack([xa[:, i, j] for i, j in pairs], 1)   # boolean_mask order: row-major lower triangle
    assert got.shape == want.shape and float((got.double() - want).abs().max()) < 1e-5
