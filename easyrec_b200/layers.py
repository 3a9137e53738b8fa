"""Dense layers of the interaction stage (K6 of SURVEY.md): fp32, TF-default semantics.

  DNN   layers/dnn.py:50-87   dense(glorot-uniform, zero bias) -> batch_norm -> relu [-> dropout]
  BN    tf.layers.batch_normalization defaults: momentum 0.99, epsilon 1e-3, batch statistics in
        training, *biased* variance for both the normalisation and the moving average (2-D
        inputs take TF's non-fused path)

Per layer: er_gemm (tcgen05 tensor cores, 3xTF32 operand split so logits stay within 1e-4 of an fp32
CPU run; forward, dX and dW read X / W[in,out] / dY in place) + liber_b200's fused bias/batch-norm/ReLU
epilogue (2 launches forward, 2 backward, deterministic statistics).  There is no torch fallback.
"""
import math

import torch
from torch import nn

from easyrec_b200 import kernels as K

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


class _DenseBNAct(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, kernel, bias, gamma, beta, moving_mean, moving_var, training, relu, ws):
    x = K.gemm_ready(x)
    z = K.gemm(x, kernel)
    y, mean, rstd = K.bias_bn_act_fwd(z, bias, gamma, beta, moving_mean, moving_var, BN_EPS,
                                      BN_MOMENTUM, training, relu, ws)
    ctx.relu = relu
    ctx.ws = ws
    ctx.has_bn = gamma is not None
    ctx.save_for_backward(x, kernel, bias, gamma, z, y, mean, rstd)
    return y

  @staticmethod
  def backward(ctx, gy):
    x, kernel, bias, gamma, z, y, mean, rstd = ctx.saved_tensors
    gz, gbias, ggamma, gbeta = K.bias_bn_act_bwd(z, bias, gamma, y, gy.contiguous(), mean, rstd,
                                                 ctx.relu, ctx.ws)
    gk = K.gemm(x.t(), gz)
    gx = K.gemm(gz, kernel.t()) if ctx.needs_input_grad[0] else None
    return gx, gk, gbias, ggamma, gbeta, None, None, None, None, None


class DenseLayer(nn.Module):
  """tf.layers.dense (kernel [in, out] glorot-uniform, bias zeros) [+ batch_norm] [+ relu]."""

  def __init__(self, n_in, n_out, use_bn, relu, generator=None):
    super().__init__()
    limit = math.sqrt(6.0 / (n_in + n_out))
    w = torch.empty(n_in, n_out)
    w.uniform_(-limit, limit, generator=generator)
    self.kernel = nn.Parameter(w)
    self.bias = nn.Parameter(torch.zeros(n_out))
    self.use_bn = use_bn
    self.relu = relu
    if use_bn:
      self.gamma = nn.Parameter(torch.ones(n_out))
      self.beta = nn.Parameter(torch.zeros(n_out))
      self.register_buffer('moving_mean', torch.zeros(n_out))
      self.register_buffer('moving_var', torch.ones(n_out))
    self.n_out = n_out
    self._ws = None
    self._ws_batch = -1

  def forward(self, x):
    if self._ws is None or self._ws_batch != x.shape[0] or self._ws.device != x.device:
      self._ws = K.dense_workspace(x.shape[0], self.n_out, x.device)
      self._ws_batch = x.shape[0]
    if self.use_bn:
      return _DenseBNAct.apply(x, self.kernel, self.bias, self.gamma, self.beta, self.moving_mean,
                               self.moving_var, self.training, self.relu, self._ws)
    return _DenseBNAct.apply(x, self.kernel, self.bias, None, None, None, None, self.training,
                             self.relu, self._ws)


class Dense(DenseLayer):
  """plain tf.layers.dense: no batch norm, no activation (logit heads)."""

  def __init__(self, n_in, n_out, generator=None):
    super().__init__(n_in, n_out, use_bn=False, relu=False, generator=generator)


class DNN(nn.Module):
  """layers/dnn.py:50-87.  Inputs of rank 3 ([B, T, d], DIN attention) are flattened to rows, which
  is exactly what tf.layers.batch_normalization does on the last axis."""

  def __init__(self, n_in, hidden_units, use_bn=True, last_layer_no_activation=False,
               last_layer_no_batch_norm=False, generator=None):
    super().__init__()
    self.layers = nn.ModuleList()
    n = len(hidden_units)
    for i, u in enumerate(hidden_units):
      bn = use_bn and (i + 1 < n or not last_layer_no_batch_norm)
      act = i + 1 < n or not last_layer_no_activation
      self.layers.append(DenseLayer(n_in, u, bn, act, generator))
      n_in = u
    self.out_dim = n_in

  def forward(self, x):
    shape = x.shape
    if x.dim() == 3:
      x = x.reshape(-1, shape[-1])
    for layer in self.layers:
      x = layer(x)
    if len(shape) == 3:
      x = x.reshape(shape[0], shape[1], -1)
    return x

  def kernels(self):
    return [layer.kernel for layer in self.layers]


class _BNOnly(torch.autograd.Function):
  """tf.layers.batch_normalization on a feature matrix (no dense, no activation):
  model/multi_tower_din.py:103-107.  Same fused kernels with an identity pre-activation."""

  @staticmethod
  def forward(ctx, x, gamma, beta, moving_mean, moving_var, training, ws):
    x = x.contiguous()
    y, mean, rstd = K.bias_bn_act_fwd(x, None, gamma, beta, moving_mean, moving_var, BN_EPS, BN_MOMENTUM,
                                      training, False, ws)
    ctx.ws = ws
    ctx.save_for_backward(x, gamma, y, mean, rstd)
    return y

  @staticmethod
  def backward(ctx, gy):
    x, gamma, y, mean, rstd = ctx.saved_tensors
    gx, _, ggamma, gbeta = K.bias_bn_act_bwd(x, None, gamma, y, gy.contiguous(), mean, rstd, False, ctx.ws)
    return gx, ggamma, gbeta, None, None, None, None


class BatchNorm(nn.Module):

  def __init__(self, units):
    super().__init__()
    self.gamma = nn.Parameter(torch.ones(units))
    self.beta = nn.Parameter(torch.zeros(units))
    self.register_buffer('moving_mean', torch.zeros(units))
    self.register_buffer('moving_var', torch.ones(units))
    self.units = units
    self._ws = None
    self._ws_batch = -1

  def forward(self, x):
    if self._ws is None or self._ws_batch != x.shape[0] or self._ws.device != x.device:
      self._ws = K.dense_workspace(x.shape[0], self.units, x.device)
      self._ws_batch = x.shape[0]
    return _BNOnly.apply(x, self.gamma, self.beta, self.moving_mean, self.moving_var, self.training, self._ws)
