"""Dense layers of the interaction stage (K6 of SURVEY.md): fp32, TF-default semantics.

  DNN   layers/dnn.py:50-87   dense(glorot-uniform, zero bias) -> batch_norm -> relu [-> dropout]
  BN    tf.layers.batch_normalization defaults: momentum 0.99, epsilon 1e-3, batch statistics in
        training, *biased* variance for both the normalisation and the moving average (2-D
        inputs take TF's non-fused path)

The GEMMs are plain library matmuls (cuBLAS SGEMM through torch, TF32 disabled so logits stay
within 1e-4 of an fp32 CPU run); everything sparse around them is liber_b200.so.
"""
import math

import torch
from torch import nn

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


class TFBatchNorm(nn.Module):

  def __init__(self, units):
    super().__init__()
    self.gamma = nn.Parameter(torch.ones(units))
    self.beta = nn.Parameter(torch.zeros(units))
    self.register_buffer('moving_mean', torch.zeros(units))
    self.register_buffer('moving_var', torch.ones(units))

  def forward(self, x):
    if self.training:
      mu = x.mean(0)
      var = ((x - mu)**2).mean(0)
      with torch.no_grad():
        self.moving_mean.mul_(BN_MOMENTUM).add_(mu.detach(), alpha=1 - BN_MOMENTUM)
        self.moving_var.mul_(BN_MOMENTUM).add_(var.detach(), alpha=1 - BN_MOMENTUM)
    else:
      mu, var = self.moving_mean, self.moving_var
    return (x - mu) * torch.rsqrt(var + BN_EPS) * self.gamma + self.beta


class Dense(nn.Module):
  """tf.layers.dense: kernel [in, out] glorot-uniform, bias zeros."""

  def __init__(self, n_in, n_out, generator=None):
    super().__init__()
    limit = math.sqrt(6.0 / (n_in + n_out))
    w = torch.empty(n_in, n_out)
    w.uniform_(-limit, limit, generator=generator)
    self.kernel = nn.Parameter(w)
    self.bias = nn.Parameter(torch.zeros(n_out))

  def forward(self, x):
    return torch.addmm(self.bias, x, self.kernel)


class DNN(nn.Module):

  def __init__(self, n_in, hidden_units, use_bn=True, last_layer_no_activation=False,
               last_layer_no_batch_norm=False, generator=None):
    super().__init__()
    self.dense = nn.ModuleList()
    self.bn = nn.ModuleList()
    self.act = []
    n = len(hidden_units)
    for i, u in enumerate(hidden_units):
      self.dense.append(Dense(n_in, u, generator))
      bn = use_bn and (i + 1 < n or not last_layer_no_batch_norm)
      self.bn.append(TFBatchNorm(u) if bn else nn.Identity())
      self.act.append(i + 1 < n or not last_layer_no_activation)
      n_in = u
    self.out_dim = n_in

  def forward(self, x):
    for d, b, a in zip(self.dense, self.bn, self.act):
      x = b(d(x))
      if a:
        x = torch.relu(x)
    return x

  def kernels(self):
    return [d.kernel for d in self.dense]
