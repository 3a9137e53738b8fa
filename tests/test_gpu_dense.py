"""GPU: fused dense epilogue (bias + batch-norm + relu, fwd/bwd) vs a plain PyTorch fp32 reference of
the same ops and vs the numpy oracle (oracle.dnn_forward/backward).  Tolerances: 2e-5 abs/rel (fp32
reductions over the batch in a different order)."""
import numpy as np
import pytest
import torch

from easyrec_b200 import layers as L
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def torch_ref_dnn(x, layers, training=True):
  """plain torch fp32: dense -> BN(batch stats, biased var, eps 1e-3) -> relu."""
  for lay in layers:
    z = x @ lay['W'] + lay['b']
    if 'gamma' in lay:
      if training:
        mu = z.mean(0)
        var = ((z - mu)**2).mean(0)
      else:
        mu, var = lay['mean'], lay['var']
      z = (z - mu) / torch.sqrt(var + 1e-3) * lay['gamma'] + lay['beta']
    x = torch.relu(z) if lay['act'] else z
  return x


@pytest.mark.parametrize('B,dims,last_plain', [(8192, [624, 256, 128, 64], False), (777, [81, 256, 33, 5], True),
                                               (64, [10, 7], False)])
def test_fused_dnn_matches_torch_and_oracle(B, dims, last_plain):
  torch.backends.cuda.matmul.allow_tf32 = False
  g = torch.Generator().manual_seed(3)
  dnn = L.DNN(dims[0], dims[1:], last_layer_no_activation=last_plain, last_layer_no_batch_norm=last_plain,
              generator=g).to(DEV)
  dnn.train()
  rng = np.random.default_rng(0)
  for lay in dnn.layers:  # non-trivial parameters
    with torch.no_grad():
      lay.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, lay.n_out).astype(np.float32)))
      if lay.use_bn:
        lay.gamma.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, lay.n_out).astype(np.float32)))
        lay.beta.copy_(torch.from_numpy(rng.normal(0, 0.2, lay.n_out).astype(np.float32)))
  x = torch.from_numpy(rng.normal(size=(B, dims[0])).astype(np.float32)).to(DEV).requires_grad_(True)
  gy = torch.from_numpy(rng.normal(size=(B, dims[-1])).astype(np.float32)).to(DEV)
  y = dnn(x)
  y.backward(gy)
  # ---- torch reference ----
  ref_layers = []
  for lay in dnn.layers:
    d = {'W': lay.kernel.detach().clone().requires_grad_(True), 'b': lay.bias.detach().clone().requires_grad_(True),
         'act': lay.relu}
    if lay.use_bn:
      d['gamma'] = lay.gamma.detach().clone().requires_grad_(True)
      d['beta'] = lay.beta.detach().clone().requires_grad_(True)
    ref_layers.append(d)
  xr = x.detach().clone().requires_grad_(True)
  yr = torch_ref_dnn(xr, ref_layers)
  yr.backward(gy)
  # Both sides are fp32 with different summation orders (tcgen05 3xTF32 GEMM vs cuBLAS SGEMM, ~1e-6 relative
  # each); batch-norm backward subtracts batch means (cancellation) and a pre-activation within that noise of
  # zero flips its ReLU mask, which changes single elements discretely.  So: the bulk must agree tightly
  # and nothing may be off by more than a few 1e-4.
  def close(a, b, rtol, atol, what, frac=0.999, slack=20.0):
    d = (a - b).abs()
    lim = atol + rtol * b.abs()
    ok = bool((d <= lim).float().mean() >= frac) and bool((d <= slack * lim).all())
    assert ok, '%s: %.4f%% outside tol, worst %.3g x tol (max abs diff %.3g)' % (
        what, 100 * float((d > lim).float().mean()), float((d / lim).max()), float(d.max()))
  close(y, yr, 2e-4, 2e-5, 'y')
  close(x.grad, xr.grad, 2e-4, 2e-5, 'x.grad')
  for li, (lay, d) in enumerate(zip(dnn.layers, ref_layers)):
    close(lay.kernel.grad, d['W'].grad, 2e-4, 2e-4, 'kernel grad %d' % li)
    if lay.use_bn:
      close(lay.gamma.grad, d['gamma'].grad, 2e-4, 2e-4, 'gamma grad %d' % li)
      close(lay.beta.grad, d['beta'].grad, 2e-4, 2e-4, 'beta grad %d' % li)
      assert float(lay.bias.grad.abs().max()) == 0.0  # identically zero under batch norm
      assert float(d['b'].grad.abs().max()) < 1e-3   # ... which torch evaluates as rounding noise
    else:
      assert torch.allclose(lay.bias.grad, d['b'].grad, rtol=2e-4, atol=2e-4)
  # ---- numpy oracle ----
  ol = []
  for lay in dnn.layers:
    d = {'W': lay.kernel.detach().cpu().numpy(), 'b': lay.bias.detach().cpu().numpy()}
    if lay.use_bn:
      d['gamma'] = lay.gamma.detach().cpu().numpy()
      d['beta'] = lay.beta.detach().cpu().numpy()
    ol.append(d)
  yo, cache = O.dnn_forward(x.detach().cpu().numpy(), ol, True, last_no_act=last_plain, last_no_bn=last_plain)
  np.testing.assert_allclose(y.detach().cpu().numpy(), yo, rtol=2e-4, atol=2e-5)
  gxo, _ = O.dnn_backward(gy.cpu().numpy(), ol, cache)
  # ReLU is discontinuous in its derivative: a pre-activation within fp32 noise of 0 may flip the mask in one
  # implementation and not the other, which changes that sample's whole input gradient.  Compare the samples
  # whose activations all stay clear of zero (the others are a ~1e-4 fraction).
  safe = np.ones(B, bool)
  for c in cache:
    if c['act']:
      safe &= (np.abs(c['h']) > 1e-4).all(axis=1)
  assert safe.mean() > 0.9
  np.testing.assert_allclose(x.grad.cpu().numpy()[safe], gxo[safe], rtol=5e-4, atol=2e-4)


def test_moving_statistics_and_inference_mode():
  torch.backends.cuda.matmul.allow_tf32 = False
  lay = L.DenseLayer(16, 8, use_bn=True, relu=True, generator=torch.Generator().manual_seed(1)).to(DEV)
  x = torch.randn(512, 16, device=DEV)
  lay.train()
  z = x @ lay.kernel + lay.bias
  mu, var = z.mean(0), ((z - z.mean(0))**2).mean(0)
  lay(x)
  assert torch.allclose(lay.moving_mean, 0.01 * mu, atol=1e-6)
  assert torch.allclose(lay.moving_var, 0.99 + 0.01 * var, atol=1e-6)
  lay.eval()
  y = lay(x)
  ref = torch.relu((z - lay.moving_mean) / torch.sqrt(lay.moving_var + 1e-3) * lay.gamma + lay.beta)
  assert torch.allclose(y, ref, atol=1e-5)
