"""Per-layer accuracy of the fused dense+BN+ReLU layer (forward y, input gradient, kernel gradient) against a
float64 evaluation, next to torch fp32 (cuBLAS SGEMM) on the same inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from easyrec_b200 import layers as L

torch.backends.cuda.matmul.allow_tf32 = False
DEV = 'cuda:0'


def ref_layer(x, W, b, gamma, beta, relu):
  z = x @ W + b
  mu = z.mean(0)
  var = ((z - mu)**2).mean(0)
  h = (z - mu) / torch.sqrt(var + 1e-3) * gamma + beta
  return (torch.relu(h) if relu else h), h


def stats(name, mine, t32, t64, mask=None):
  if mask is not None:
    mine, t32, t64 = mine[mask], t32[mask], t64[mask]
  em = (mine.double() - t64).abs()
  et = (t32.double() - t64).abs()
  sc = float(t64.abs().mean())
  print('  %-8s scale %.3g | mine: rms %.2e p99.9 %.2e max %.2e | torch32: rms %.2e p99.9 %.2e max %.2e' % (
      name, sc, float(em.pow(2).mean().sqrt()), float(torch.quantile(em.flatten()[:4000000], 0.999)), float(em.max()),
      float(et.pow(2).mean().sqrt()), float(torch.quantile(et.flatten()[:4000000], 0.999)), float(et.max())))


B = 8192
rng = np.random.default_rng(0)
for (kin, kout) in [(624, 256), (256, 128), (128, 64)]:
  print('layer %d -> %d' % (kin, kout))
  lay = L.DenseLayer(kin, kout, True, True, torch.Generator().manual_seed(1)).to(DEV)
  lay.train()
  with torch.no_grad():
    lay.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, kout).astype(np.float32)))
    lay.gamma.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, kout).astype(np.float32)))
    lay.beta.copy_(torch.from_numpy(rng.normal(0, 0.2, kout).astype(np.float32)))
  x = torch.from_numpy(rng.normal(size=(B, kin)).astype(np.float32)).to(DEV)
  if kin != 624:
    x = torch.relu(x)      # hidden layers see post-ReLU inputs
  gy = torch.from_numpy(rng.normal(size=(B, kout)).astype(np.float32)).to(DEV)
  xm = x.clone().requires_grad_(True)
  ym = lay(xm)
  ym.backward(gy)
  res = {}
  for tag, dt in (('t32', torch.float32), ('t64', torch.float64)):
    xr = x.detach().to(dt).clone().requires_grad_(True)
    W = lay.kernel.detach().to(dt).clone().requires_grad_(True)
    yr, h = ref_layer(xr, W, lay.bias.detach().to(dt), lay.gamma.detach().to(dt), lay.beta.detach().to(dt), True)
    yr.backward(gy.to(dt))
    res[tag] = (yr.detach(), xr.grad, W.grad, h.detach())
  safe = res['t64'][3].abs().min(dim=1).values > 1e-4
  print('  safe rows %.3f' % float(safe.float().mean()))
  stats('y', ym.detach(), res['t32'][0], res['t64'][0], safe)
  stats('x.grad', xm.grad, res['t32'][1], res['t64'][1], safe)
  stats('W.grad', lay.kernel.grad, res['t32'][2], res['t64'][2])
