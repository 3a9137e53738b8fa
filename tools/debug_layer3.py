import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from easyrec_b200 import layers as L, kernels as K
torch.backends.cuda.matmul.allow_tf32 = False
DEV = 'cuda:0'
B = 8192
rng = np.random.default_rng(0)
stash = {}
orig = K.bias_bn_act_bwd
def patched(*a, **k):
  r = orig(*a, **k)
  stash['gz'] = r[0]
  stash['args'] = a
  return r
K.bias_bn_act_bwd = patched
orig_gemm = K.gemm
calls = []
def pg(a, b, bias=None, out=None):
  r = orig_gemm(a, b, bias=bias, out=out)
  calls.append((a, b, r))
  return r
K.gemm = pg
SET = os.environ.get('SET', '1') == '1'
for (kin, kout) in [(624, 256), (256, 128), (128, 64)]:
  del calls[:]
  lay = L.DenseLayer(kin, kout, True, True, torch.Generator().manual_seed(1)).to(DEV)
  lay.train()
  if SET:
    with torch.no_grad():
      lay.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, kout).astype(np.float32)))
      lay.gamma.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, kout).astype(np.float32)))
      lay.beta.copy_(torch.from_numpy(rng.normal(0, 0.2, kout).astype(np.float32)))
  x = torch.from_numpy(rng.normal(size=(B, kin)).astype(np.float32)).to(DEV)
  if kin != 624:
    x = torch.relu(x)
  gy = torch.from_numpy(rng.normal(size=(B, kout)).astype(np.float32)).to(DEV)
  xm = x.clone().requires_grad_(True)
  y = lay(xm)
  y.backward(gy)
  torch.cuda.synchronize()
  print('layer %d->%d' % (kin, kout))
  # float64 BN backward from the kernel's own inputs
  z, bias, gamma, yy, gyy, mean, rstd = stash['args'][:7]
  zb = (z + bias).double()
  xhat = (zb - mean.double()) * rstd.double()
  gp = gyy.double() * (yy > 0)
  a_ = gp.sum(0); b_ = (gp * xhat).sum(0)
  gz_ref = gamma.double() * rstd.double() * (gp - a_ / B - xhat * b_ / B)
  e = (stash['gz'].double() - gz_ref).abs()
  print('  gz vs f64 formula on the same inputs: rms %.2e max %.2e' % (float(e.pow(2).mean().sqrt()), float(e.max())))
  mu64 = zb.mean(0); var64 = ((zb - mu64) ** 2).mean(0)
  print('  save_mean err %.2e  save_rstd rel err %.2e' % (float((mean.double() - mu64).abs().max()),
        float(((rstd.double() - 1 / torch.sqrt(var64 + 1e-3)) * torch.sqrt(var64 + 1e-3)).abs().max())))
  for i, (a, b, r) in enumerate(calls):
    ref = a.detach().double() @ b.detach().double()
    e = (r.detach().double() - ref).abs()
    print('  gemm call %d: a %s %s b %s %s -> rms %.2e max %.2e (scale %.3g)' % (i, tuple(a.shape), a.stride(), tuple(b.shape), b.stride(),
          float(e.pow(2).mean().sqrt()), float(e.max()), float(ref.abs().mean())))
  e = (lay.kernel.grad.double() - (calls[0][0].detach().double() @ calls[0][1].detach().double())).abs()
  print('  kernel.grad vs f64 of the same operands: rms %.2e max %.2e' % (float(e.pow(2).mean().sqrt()), float(e.max())))
