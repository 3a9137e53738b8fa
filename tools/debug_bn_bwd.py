import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyrec_b200 import kernels as K
g = torch.Generator(device='cuda').manual_seed(0)
B = 8192
for units in (64, 128, 256, 48):
  z = torch.randn(B, units, device='cuda', generator=g)
  bias = torch.randn(units, device='cuda', generator=g) * 0.1
  gamma = torch.rand(units, device='cuda', generator=g) + 0.5
  beta = torch.randn(units, device='cuda', generator=g) * 0.2
  gy = torch.randn(B, units, device='cuda', generator=g)
  zb = (z + bias).double()
  mu = zb.mean(0); var = ((zb - mu) ** 2).mean(0); rstd = 1 / torch.sqrt(var + 1e-3)
  xhat = (zb - mu) * rstd
  h = xhat * gamma.double() + beta.double()
  y = torch.relu(h)
  gp = gy.double() * (h > 0)
  a = gp.sum(0); b = (gp * xhat).sum(0)
  gz_ref = gamma.double() * rstd * (gp - a / B - xhat * b / B)
  ws = K.dense_workspace(B, units, 'cuda')
  gz, gbias, ggamma, gbeta = K.bias_bn_act_bwd(z, bias, gamma, y.float(), gy, mu.float(), rstd.float(), True, ws)
  e = (gz.double() - gz_ref).abs()
  print('units %3d: gz rms err %.2e max %.2e | colsum(gz) max %.2e (ref %.2e) | gbeta err %.2e ggamma err %.2e' % (
      units, float(e.pow(2).mean().sqrt()), float(e.max()), float(gz.double().sum(0).abs().max()),
      float(gz_ref.sum(0).abs().max()), float((gbeta.double() - a).abs().max()), float((ggamma.double() - b).abs().max())))
