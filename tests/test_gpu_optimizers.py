"""GPU: the optimizers of the BASELINE configs on the fused row update.

  * `adam_optimizer` = tf.train.AdamOptimizer (builders/optimizer_builder.py:61-66): K7's row rule on the rows of the
    batch + er_adam_dense_sweep on all other rows == the oracle's restatement of TF's _apply_sparse_shared, 1 and 10
    steps, <= 1e-6 (every row of the table is compared, touched or not);
  * hyper-parameters read from device memory (er_opt_t.hyper_dev) give bit-identical results to the same values
    passed in the struct, for every optimizer kind;
  * a CUDA-graph-captured training run (Adagrad, lazy Adam and Adam rows, exponentially decaying learning rate)
    produces the same tables / parameters as the eager run of the same batches: the captured graph follows the
    schedule and the beta powers, and no batch is applied twice around the capture.
"""
import numpy as np
import pytest
import torch

from easyrec_b200 import _lib, embedding as E, kernels as K
from easyrec_b200.config import config_util
from easyrec_b200.estimator import EasyRecEstimator
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def t(a):
  return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _plan(V, B, F, dim):
  stride = F * dim
  recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=3, combiner=0, out_buf=0,
               out_stride=stride, out_col=f * dim) for f in range(F)]
  return K.slots_to_device(K.make_slots(recs), DEV), stride


@pytest.mark.parametrize('dim,interleave', [(16, True), (16, False), (6, True), (1, True), (32, True)])
@pytest.mark.parametrize('steps', [1, 10])
def test_tf_adam_rows_plus_dense_sweep_track_the_oracle(dim, interleave, steps):
  rng = np.random.default_rng(dim * 10 + steps)
  V, B, F = 3000, 200, 3
  arena = E.Arena(dim, DEV)
  arena.add_table('t', V)
  arena.materialize(_lib.OPT_ADAM_ROWS, generator=torch.Generator(device=DEV).manual_seed(1), interleave=interleave)
  table = arena.weight.cpu().numpy().copy()
  m = np.zeros((V, dim), np.float32)
  v = np.zeros((V, dim), np.float32)
  sd, stride = _plan(V, B, F, dim)
  pad = (4 - stride % 4) % 4 if dim % 4 == 0 else 0
  ws = K.bwd_workspace(B * F, DEV, dim)
  hyper = K.StepHyper(DEV, 0.9, 0.999)
  for step in range(steps):
    # step 0 touches many rows, later steps few: the early rows must keep decaying
    hi = V if step == 0 else 40
    rows = rng.integers(0, hi, B * F).astype(np.int64)
    rows[rng.integers(0, B * F, 7)] = -1
    gout = rng.normal(0, 0.1, (B, stride + pad)).astype(np.float32)
    lr = 0.01 * (0.7**step)
    hyper.set(lr, step, grad_scale=0.5)
    opt = hyper.opt(_lib.OPT_ADAM_ROWS)
    d_rows = t(rows)
    K.embedding_bwd(arena.weight, arena.state0, arena.state1, dim, d_rows, sd, F, B * F, [t(gout)], opt, ws)
    E.adam_dense_decay(arena, d_rows, opt)
    gseg = np.concatenate([gout[:, f * dim:(f + 1) * dim] for f in range(F)], 0)
    O.embedding_bwd_adam_dense(table, m, v, rows, None, gseg, hyper.lr, beta1_power=float(hyper.b1p),
                               beta2_power=float(hyper.b2p), grad_scale=0.5)
  torch.cuda.synchronize()
  assert not arena.touched.any()
  np.testing.assert_allclose(arena.weight.cpu().numpy(), table, rtol=0, atol=1e-6)
  np.testing.assert_allclose(arena.state0.cpu().numpy(), m, rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(arena.state1.cpu().numpy(), v, rtol=1e-6, atol=1e-9)
  if steps > 1:   # rows only seen at step 0 moved after it: the dense half really ran
    assert (np.abs(m[100:]).sum(1) > 0).sum() > 100


@pytest.mark.parametrize('kind', [_lib.OPT_SGD, _lib.OPT_ADAGRAD, _lib.OPT_LAZY_ADAM, _lib.OPT_ADAM_ROWS])
@pytest.mark.parametrize('dim', [16, 1, 6])
def test_hyper_parameters_from_device_memory_equal_the_struct_path(kind, dim):
  rng = np.random.default_rng(kind * 7 + dim)
  V, B, F = 2000, 300, 2
  sd, stride = _plan(V, B, F, dim)
  pad = (4 - stride % 4) % 4 if dim % 4 == 0 else 0
  table = rng.normal(size=(V, dim)).astype(np.float32)
  rows = (rng.zipf(1.2, B * F) % V).astype(np.int64)
  gout = rng.normal(size=(B, stride + pad)).astype(np.float32)
  ws = K.bwd_workspace(B * F, DEV, dim)
  res = []
  for via_device in (False, True):
    hyper = K.StepHyper(DEV, 0.9, 0.999)
    hyper.set(0.03, 6, grad_scale=0.25)
    if via_device:
      opt = hyper.opt(kind)
      # struct fields deliberately wrong: the kernels must take lr / powers / scale from the device block
      opt.lr, opt.beta1_power, opt.beta2_power, opt.grad_scale = 123.0, 0.5, 0.5, 77.0
    else:
      opt = K.make_opt(kind, hyper.lr, 0.9, 0.999, 1e-8, float(hyper.b1p), float(hyper.b2p), 0.25)
    d_t = t(table)
    s0 = torch.full((V, dim), 0.1, device=DEV) if kind != _lib.OPT_SGD else None
    s1 = torch.full((V, dim), 0.2, device=DEV) if kind in (_lib.OPT_LAZY_ADAM, _lib.OPT_ADAM_ROWS) else None
    K.embedding_bwd(d_t, s0, s1, dim, t(rows), sd, F, B * F, [t(gout)], opt, ws)
    if kind == _lib.OPT_ADAM_ROWS:
      K.adam_dense_sweep(d_t, s0, s1, dim, None, opt)
    res.append([x.cpu().numpy() for x in (d_t, s0, s1) if x is not None])
  for a, b in zip(*res):
    np.testing.assert_array_equal(a, b)


CFG = '''
train_config { optimizer_config { %(opt)s { learning_rate { exponential_decay_learning_rate {
  initial_learning_rate: 0.02 decay_steps: 3 decay_factor: 0.6 min_learning_rate: 0.0001 } } } } }
data_config { batch_size: 512 input_type: CSVInput separator: "," label_fields: "label"
  input_fields { input_name: "label" input_type: FLOAT } input_fields { input_name: "F1" input_type: FLOAT }
  input_fields { input_name: "C1" input_type: INT64 } input_fields { input_name: "C2" input_type: INT64 } }
feature_config {
  features { input_names: "F1" feature_type: RawFeature embedding_dim: 16 min_val: 0.0 max_val: 10.0 }
  features { input_names: "C1" feature_type: IdFeature embedding_dim: 16 hash_bucket_size: 5000 }
  features { input_names: "C2" feature_type: IdFeature embedding_dim: 16 num_buckets: 50 } }
model_config { model_class: "DeepFM"
  feature_groups { group_name: "deep" feature_names: ["F1", "C1", "C2"] wide_deep: DEEP }
  feature_groups { group_name: "wide" feature_names: ["F1", "C1", "C2"] wide_deep: WIDE }
  deepfm { dnn { hidden_units: [64, 32] } final_dnn { hidden_units: [32] } l2_regularization: 1e-6 }
  embedding_regularization: 1e-6 }
'''


def _batches(n, B=512):
  rng = np.random.default_rng(11)
  out = []
  for _ in range(n):
    ids = np.concatenate([rng.integers(0, 10**9, B), rng.integers(0, 50, B)]).astype(np.int64)
    dense = rng.uniform(0, 10, (B, 1)).astype(np.float32)
    labels = (rng.uniform(size=B) < 0.3).astype(np.float32)
    out.append(({'sparse_fea': t(ids), 'dense_fea': t(dense)}, t(labels)))
  return out


@pytest.mark.parametrize('opt', ['adagrad_optimizer', 'lazy_adam_optimizer', 'adam_optimizer'])
def test_cuda_graph_run_equals_eager_run_for_every_optimizer(opt):
  """8 steps (2 eager warm-up steps + capture + 5 replays) against 8 eager steps on the same batches: the decayed
  learning rate (3 decay boundaries inside the run) and Adam's beta powers reach the captured kernels through device
  memory; every batch is applied exactly once.  Deterministic kernels -> bit-identical state."""
  cfg = config_util.get_configs_from_pipeline_file((CFG % dict(opt=opt)).encode())
  batches = _batches(8)
  runs = []
  for graph in (False, True):
    est = EasyRecEstimator(cfg, device=DEV, seed=3, use_cuda_graph=graph)
    losses = []
    for f, l in batches:
      loss, _ = est.trainer.train_step(f, l)
      losses.append(float(loss))
    torch.cuda.synchronize()
    if graph:
      assert est.trainer._graph is not None and est.trainer.launches_per_step > 10
    runs.append((losses, {d: a.storage.clone() for d, a in est.input_layer.arenas.items()},
                 est.trainer.dense_opt.flat_p.clone(), est.trainer.dense_opt.s0.clone(),
                 {k: v.clone() for k, v in est.model.state_dict().items()}))
  (l0, a0, p0, s0, sd0), (l1, a1, p1, s1, sd1) = runs
  assert l0 == l1, (l0, l1)
  for d in a0:
    assert torch.equal(a0[d], a1[d]), 'arena %d differs' % d
  assert torch.equal(p0, p1) and torch.equal(s0, s1)
  for k in sd0:
    assert torch.equal(sd0[k], sd1[k]), k     # incl. batch-norm moving statistics: no double application
  assert l0[-1] != l0[0]


def test_global_norm_clipping_on_the_kernels_eager_and_captured():
  """train_config.gradient_clipping_by_norm through the real K7 (emit form over per-column virtual rows) and the
  device-resident gradient scale: plain SGD makes every update linear in its gradient, so clipped = scale * unclipped
  for every dense parameter and table row; the captured step follows the eager one."""
  import sys, os
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  from test_round2_host import CLIP_CFG
  from easyrec_b200.estimator import EasyRecEstimator
  torch.backends.cuda.matmul.allow_tf32 = False
  rng = np.random.default_rng(0)
  B = 16
  batches = []
  for _ in range(5):
    ids = np.stack([rng.integers(0, 6, B), rng.integers(0, 6, B), rng.integers(0, 1000, B)]).astype(np.int64)
    batches.append(({'sparse_fea': torch.from_numpy(ids.reshape(-1)).to(DEV),
                     'dense_fea': torch.from_numpy(rng.uniform(0, 2, (B, 1)).astype(np.float32)).to(DEV)},
                    torch.from_numpy((rng.uniform(size=B) < 0.4).astype(np.float32)).to(DEV)))
  plain = EasyRecEstimator(CLIP_CFG % b'', device=DEV, seed=11)
  clip = EasyRecEstimator(CLIP_CFG % b'gradient_clipping_by_norm: 0.05', device=DEV, seed=11)
  p0 = plain.trainer.dense_opt.flat_p.clone()
  t0 = {d: a.weight.clone() for d, a in plain.input_layer.arenas.items()}
  plain.trainer.train_step(*batches[0])
  clip.trainer.train_step(*batches[0])
  norm = float(clip.trainer.last_grad_norm)
  assert norm > 0.05
  scale = 0.05 / norm
  torch.testing.assert_close(clip.trainer.dense_opt.flat_p - p0, (plain.trainer.dense_opt.flat_p - p0) * scale,
                             rtol=2e-4, atol=2e-7)
  for d, a in clip.input_layer.arenas.items():
    want = (plain.input_layer.arenas[d].weight - t0[d]) * scale
    assert float(want.abs().max()) > 1e-5
    torch.testing.assert_close(a.weight - t0[d], want, rtol=2e-4, atol=2e-8)
  # captured: the same five steps eager and from the graph (two eager steps, capture, replays)
  runs = []
  for graph in (False, True):
    e = EasyRecEstimator(CLIP_CFG % b'gradient_clipping_by_norm: 0.05', device=DEV, seed=11, use_cuda_graph=graph)
    losses = [float(e.trainer.train_step(*b)[0]) for b in batches]
    runs.append((losses, e.trainer.dense_opt.flat_p.clone(), float(e.trainer.last_grad_norm)))
    assert (not graph) or e.trainer._graph is not None
  np.testing.assert_allclose(runs[1][0], runs[0][0], rtol=0, atol=1e-6)
  torch.testing.assert_close(runs[1][1], runs[0][1], rtol=0, atol=1e-7)
  assert abs(runs[1][2] - runs[0][2]) < 1e-5 * max(1.0, runs[0][2])
