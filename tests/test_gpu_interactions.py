"""GPU: fused interaction kernels (DIN, DCN cross, MMoE, DSSM pieces) vs plain PyTorch fp32 references that
restate the reference TF code line by line.  Tolerance 1e-5 abs / 1e-5 rel (fp32, different reduction order)."""
import pytest
import torch

from easyrec_b200 import interactions as I

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = dict(rtol=2e-5, atol=2e-5)


def _grads(fn, inputs, gout):
  xs = [x.detach().clone().requires_grad_(True) for x in inputs]
  y = fn(*xs)
  y.backward(gout)
  return y.detach(), [x.grad for x in xs]


@pytest.mark.parametrize('B,T,D', [(64, 50, 16), (7, 3, 5), (4096, 50, 16), (33, 70, 32)])
def test_din_attention_matches_tf_formula(B, T, D):
  g = torch.Generator(device=DEV).manual_seed(B + T)
  q = torch.randn(B, D, device=DEV, generator=g)
  keys = torch.randn(B, T, D, device=DEV, generator=g)
  lens = torch.randint(0, T + 1, (B,), device=DEV, generator=g, dtype=torch.int32)
  lens[0] = 0  # fully padded history -> uniform softmax (sequence_feature_layer.py:172-177)
  W = torch.randn(4 * D, 1, device=DEV, generator=g) * 0.3
  gout = torch.randn(B, D, device=DEV, generator=g)

  def ref(q, keys, W):
    # sequence_feature_layer.py:150-189
    cur = q[:, None, :].expand(B, T, D)
    din = torch.cat([cur, keys, cur - keys, cur * keys], dim=-1)
    scores = (din @ W).reshape(B, 1, T)
    mask = (torch.arange(T, device=DEV)[None, :] < lens[:, None])[:, None, :]
    scores = torch.where(mask, scores, torch.full_like(scores, -2.0**32 + 1))
    scores = torch.softmax(scores, dim=-1)
    return (scores @ keys).reshape(B, D)

  def fused(q, keys, W):
    return I.din_attention(q, keys, lens, lambda x: x @ W)

  y0, g0 = _grads(ref, [q, keys, W], gout)
  y1, g1 = _grads(fused, [q, keys, W], gout)
  assert torch.allclose(y1, y0, **TOL)
  for a, b in zip(g1, g0):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('B,D', [(8192, 624), (17, 5), (300, 1280)])
def test_cross_layer_matches_dcn_v1(B, D):
  g = torch.Generator(device=DEV).manual_seed(D)
  x0 = torch.randn(B, D, device=DEV, generator=g)
  xl = torch.randn(B, D, device=DEV, generator=g)
  w = torch.randn(D, device=DEV, generator=g) * 0.05
  b = torch.randn(D, device=DEV, generator=g) * 0.1
  gout = torch.randn(B, D, device=DEV, generator=g)

  def ref(x0, xl, w, b):  # model/dcn.py:41-44
    xw = (xl * w).sum(dim=1, keepdim=True)
    return x0 * xw + b + xl

  y0, g0 = _grads(ref, [x0, xl, w, b], gout)
  y1, g1 = _grads(I.cross_layer, [x0, xl, w, b], gout)
  assert torch.allclose(y1, y0, rtol=1e-4, atol=1e-4)
  for a, b_ in zip(g1, g0):
    assert torch.allclose(a, b_, rtol=2e-4, atol=2e-3 if B > 1000 else 2e-4)


@pytest.mark.parametrize('B,E,H', [(16384, 4, 64), (11, 3, 7), (100, 40, 33)])
def test_mmoe_mix_matches(B, E, H):
  g = torch.Generator(device=DEV).manual_seed(E)
  gate = torch.randn(B, E, device=DEV, generator=g)
  experts = torch.randn(B, E, H, device=DEV, generator=g)
  gout = torch.randn(B, H, device=DEV, generator=g)

  def ref(gate, experts):  # layers/mmoe.py:73-83
    return (experts * torch.softmax(gate, dim=1)[:, :, None]).sum(dim=1)

  y0, g0 = _grads(ref, [gate, experts], gout)
  y1, g1 = _grads(I.mmoe_mix, [gate, experts], gout)
  assert torch.allclose(y1, y0, **TOL)
  for a, b in zip(g1, g0):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_l2_normalize_and_inbatch_softmax():
  g = torch.Generator(device=DEV).manual_seed(9)
  B, H = 512, 32
  u = torch.randn(B, H, device=DEV, generator=g)
  it = torch.randn(B, H, device=DEV, generator=g)
  item_ids = torch.randint(0, 200, (B,), device=DEV, generator=g)  # many duplicate items in the batch
  gu = torch.randn(B, H, device=DEV, generator=g)
  y0, g0 = _grads(lambda x: torch.nn.functional.normalize(x, dim=-1, eps=1e-6), [u], gu)
  y1, g1 = _grads(I.l2_normalize, [u], gu)
  assert torch.allclose(y1, y0, **TOL) and torch.allclose(g1[0], g0[0], rtol=1e-4, atol=1e-5)

  temperature = 0.05

  def ref_loss(u, it):  # dssm.py:64-71, match_model.py:50-69, 213-234
    un = torch.nn.functional.normalize(u, dim=-1)
    inn = torch.nn.functional.normalize(it, dim=-1)
    sim = un @ inn.t() / temperature
    dup = (item_ids[None, :] == item_ids[:, None]).float() - torch.eye(B, device=DEV)
    logits = sim - dup * 1e32
    probs = torch.softmax(logits, dim=1)
    hit = probs.diagonal()
    return -(torch.log(hit + 1e-12)).mean()

  def fused_loss(u, it):
    sim = I.l2_normalize(u) @ I.l2_normalize(it).t() / temperature
    return I.inbatch_softmax_ce(sim, item_ids)[0]

  one = torch.ones((), device=DEV)
  l0, gr0 = _grads(ref_loss, [u, it], one)
  l1, gr1 = _grads(fused_loss, [u, it], one)
  assert abs(float(l0) - float(l1)) < 1e-4
  for a, b in zip(gr1, gr0):
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)


def test_kernels_reproduce_the_reference_code_outputs():
  """FM, DCN cross and DLRM dot interaction against golden vectors produced by executing the reference's own
  functions (tests/golden/make_formula_golden.py; layers/fm.py, model/dcn.py, layers/keras/interaction.py)."""
  import json
  import os
  from easyrec_b200 import backbone as BB, kernels as K
  cases = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_formulas.json')))['cases']
  c = cases['fm']
  x = torch.tensor(c['x'], device=DEV)                      # [B, F, D]
  B, F, D = x.shape
  want = torch.tensor(c['y'], device=DEV)
  assert torch.allclose(K.fm_fwd(x.reshape(B, F * D).contiguous(), F, D), want, rtol=1e-5, atol=1e-5)
  y, _ = K.fm_block_fwd(x.reshape(B, F * D).contiguous(), F, D)
  assert torch.allclose(y, want, rtol=1e-5, atol=1e-5)
  c = cases['dcn_cross']
  x0 = torch.tensor(c['x'], device=DEV)
  xl = x0
  for w, b in zip(c['w'], c['b']):
    xl = I.cross_layer(x0, xl.contiguous(), torch.tensor(w, device=DEV), torch.tensor(b, device=DEV))
  assert torch.allclose(xl, torch.tensor(c['y'], device=DEV), rtol=1e-5, atol=1e-5)
  for key in ('dot_interaction_self0', 'dot_interaction_self1'):
    c = cases[key]
    x = torch.tensor(c['x'], device=DEV)                    # [B, F, D]
    got = BB.DotInteraction({'self_interaction': c['self_interaction']})([x[:, i] for i in range(x.shape[1])])
    assert torch.allclose(got, torch.tensor(c['y'], device=DEV), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('B,n,d', [(64, 27, 16), (5, 3, 7), (8192, 27, 16), (130, 40, 32)])
def test_gram_matches_bmm(B, n, d):
  """DLRM / DotInteraction pairwise dot products (model/dlrm.py:52-61 einsum 'bne,bme->bnm') and their gradient"""
  g = torch.Generator(device=DEV).manual_seed(B + n)
  x = torch.randn(B, n, d, device=DEV, generator=g)
  gout = torch.randn(B, n, n, device=DEV, generator=g)
  y0, (gx0,) = _grads(lambda x: torch.einsum('bne,bme->bnm', x, x), [x], gout)
  y1, (gx1,) = _grads(I.gram, [x], gout)
  torch.testing.assert_close(y1, y0, **TOL)
  torch.testing.assert_close(gx1, gx0, rtol=1e-4, atol=1e-4)
  assert torch.equal(y1, y1.transpose(1, 2))   # same operands, same order of additions: exactly symmetric


@pytest.mark.parametrize('B,C,H', [(512, 512, 32), (100, 260, 16), (4096, 4096, 32)])
def test_in_batch_similarity_matmul_matches_torch(B, C, H):
  """MatchModel's U I^T (model/match_model.py:92-97) on er_gemm (3xTF32): fp32-level accuracy, both gradients"""
  torch.backends.cuda.matmul.allow_tf32 = False
  g = torch.Generator(device=DEV).manual_seed(B)
  u = torch.randn(B, H, device=DEV, generator=g)
  i = torch.randn(C, H, device=DEV, generator=g)
  gout = torch.randn(B, C, device=DEV, generator=g)
  y0, (gu0, gi0) = _grads(lambda u, i: u.double() @ i.double().t(), [u, i], gout.double())
  y1, (gu1, gi1) = _grads(I.matmul_nt, [u, i], gout)
  torch.testing.assert_close(y1.double(), y0, rtol=1e-5, atol=2e-5)
  # K = C (or B) products of N(0,1) pairs, |result| ~ sqrt(K): 3xTF32 with fp32 accumulation, <= ~1e-5 of that scale
  torch.testing.assert_close(gu1.double(), gu0.double(), rtol=1e-5, atol=1e-5 * C ** 0.5 + 2e-4)
  torch.testing.assert_close(gi1.double(), gi0.double(), rtol=1e-5, atol=1e-5 * B ** 0.5 + 2e-4)
