"""GPU: the CUDA path against golden vectors produced by EXECUTING the reference's own function bodies on a
numpy shim of the tf ops they call (tests/golden/make_formula_golden.py -> reference_formulas.json): DIN target
attention, MMoE, keras Cross (full-rank / low-rank), list-wise match loss with duplicate-item masking, lazy Adam.
fp32 with a different reduction order (and the 3xTF32 tensor-core GEMM for the dense layers): 1e-5 abs / 1e-4 rel."""
import json
import os

import numpy as np
import pytest
import torch

from easyrec_b200 import _lib, backbone as BB, interactions as I, kernels as K, layers as L

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = dict(rtol=1e-4, atol=1e-5)
CASES = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_formulas.json')))['cases']


def t(a, dtype=torch.float32):
  return torch.tensor(a, dtype=dtype, device=DEV)


def _load_mlp(layers_json, last_linear):
  """layers.DNN without batch norm carrying the golden weights."""
  dims = [len(layers_json[0]['w'])] + [len(l['b']) for l in layers_json]
  dnn = L.DNN(dims[0], dims[1:], use_bn=False, last_layer_no_activation=last_linear,
              last_layer_no_batch_norm=last_linear).to(DEV)
  with torch.no_grad():
    for layer, l in zip(dnn.layers, layers_json):
      layer.kernel.copy_(t(l['w']))
      layer.bias.copy_(t(l['b']))
  return dnn


def test_din_attention_reproduces_target_attention():
  c = CASES['din_target_attention']
  key, hist = t(c['key']), t(c['hist'])
  lens = t(c['lens'], torch.int32)
  dnn = _load_mlp(c['mlp'], last_linear=True)
  att = I.din_attention(key, hist, lens, dnn)
  got = torch.cat([att, key], dim=1)            # din_output = concat([hist_din_emb, cur_id])
  assert torch.allclose(got, t(c['y']), **TOL)
  # a fully padded history attends uniformly: mean of the (padded) keys
  assert torch.allclose(att[0], hist[0].mean(0), **TOL)


def test_mmoe_reproduces_layers_mmoe():
  c = CASES['mmoe']
  x = t(c['x'])
  experts = torch.stack([_load_mlp(e, last_linear=False)(x) for e in c['experts']], dim=1).contiguous()
  for g, want in zip(c['gates'], c['y']):
    gate = L.Dense(len(g['w']), len(g['b'])).to(DEV)
    with torch.no_grad():
      gate.kernel.copy_(t(g['w']))
      gate.bias.copy_(t(g['b']))
    assert torch.allclose(I.mmoe_mix(gate(x), experts), t(want), **TOL)


@pytest.mark.parametrize('case', ['keras_cross_full', 'keras_cross_lowrank'])
def test_backbone_cross_reproduces_keras_cross(case):
  c = CASES[case]
  d = len(c['b'])
  if 'u' in c:
    cross = BB.Cross(d, {'projection_dim': float(len(c['u'][0]))}).to(DEV)
    with torch.no_grad():
      cross.dense_u.kernel.copy_(t(c['u']))
      cross.dense.kernel.copy_(t(c['v']))
  else:
    cross = BB.Cross(d, {'diag_scale': c['diag_scale']}).to(DEV)
    with torch.no_grad():
      cross.dense.kernel.copy_(t(c['w']))
  with torch.no_grad():
    cross.dense.bias.copy_(t(c['b']))
  assert torch.allclose(cross([t(c['x0']), t(c['x'])]), t(c['y']), **TOL)


def test_inbatch_softmax_ce_reproduces_match_model_listwise_loss():
  c = CASES['match_listwise']
  user, item = t(c['user']), t(c['item'])
  # towers are already unit-norm in the golden case: l2_normalize must leave them alone
  assert torch.allclose(I.l2_normalize(user), user, **TOL)
  sim = (user @ item.t()) / c['temperature']
  assert torch.allclose(sim, t(c['sim']), **TOL)
  loss, p_diag = I.inbatch_softmax_ce(t(c['sim']), t(c['item_ids'], torch.int64), t(c['sample_weight']))
  assert abs(float(loss) - c['cross_entropy_loss']) < 1e-5 * max(1.0, abs(c['cross_entropy_loss']))
  assert torch.allclose(p_diag, torch.diagonal(t(c['probs'])), **TOL)


def _lazy_adam_state(c):
  w = t(c['w0'])
  return w, torch.zeros_like(w), torch.zeros_like(w)


def test_fused_backward_update_reproduces_adam_s_lazy_adam():
  """K7 (dedup + fused row update) on one single-valued slot whose lookups are the golden unique rows."""
  c = CASES['lazy_adam_sparse']
  w, m, v = _lazy_adam_state(c)
  V, dim = w.shape
  p1, p2 = c['beta1'], c['beta2']
  for st in c['steps']:
    rows = t(st['indices'], torch.int64)
    n = rows.numel()
    g = t(st['grad'])
    slots = K.make_slots([dict(num_buckets=V, row_offset=0, seg_begin=0, n_seg=n, bucket_mode=3, combiner=0,
                               out_buf=0, out_stride=dim, out_col=0)])
    opt = K.make_opt(_lib.OPT_LAZY_ADAM, c['lr'], beta1=c['beta1'], beta2=c['beta2'], eps=c['epsilon'],
                     beta1_power=p1, beta2_power=p2)
    K.embedding_bwd(w, m, v, dim, rows, K.slots_to_device(slots, DEV), 1, n, [g], opt, K.bwd_workspace(n, DEV, dim))
    p1, p2 = p1 * c['beta1'], p2 * c['beta2']
    torch.cuda.synchronize()
    np.testing.assert_allclose(m.cpu().numpy(), np.array(st['m'], np.float32), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), np.array(st['v'], np.float32), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(w.cpu().numpy(), np.array(st['w'], np.float32), rtol=2e-6, atol=1e-8)


def test_sparse_apply_reproduces_adam_s_lazy_adam():
  """er_sparse_apply (already deduplicated gradient, the data-parallel path) on the same golden steps."""
  c = CASES['lazy_adam_sparse']
  w, m, v = _lazy_adam_state(c)
  dim = w.shape[1]
  p1, p2 = c['beta1'], c['beta2']
  for st in c['steps']:
    rows = t(st['indices'], torch.int64)
    n_uniq = torch.tensor([rows.numel()], dtype=torch.int32, device=DEV)
    opt = K.make_opt(_lib.OPT_LAZY_ADAM, c['lr'], beta1=c['beta1'], beta2=c['beta2'], eps=c['epsilon'],
                     beta1_power=p1, beta2_power=p2)
    K.sparse_apply(w, m, v, dim, rows, t(st['grad']), n_uniq, opt)
    p1, p2 = p1 * c['beta1'], p2 * c['beta2']
    torch.cuda.synchronize()
    np.testing.assert_allclose(m.cpu().numpy(), np.array(st['m'], np.float32), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), np.array(st['v'], np.float32), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(w.cpu().numpy(), np.array(st['w'], np.float32), rtol=2e-6, atol=1e-8)
