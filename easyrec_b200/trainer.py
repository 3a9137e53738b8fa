"""Training step driver: the torch-side replacement of the reference's hot loop
(`sess.run(train_op)` under EasyRecEstimator._train_model_fn, model/easy_rec_estimator.py:155-472,
and optimize_loss, compat/optimizers.py:89-450).

One step = K1 bucketize -> K2 gather+pool -> interaction + dense MLP -> loss -> backward ->
K7 dedup + fused row update -> one fused dense-optimizer launch.  The whole step is shape-static,
so it is captured once into a CUDA graph and replayed (launch latency, not HBM, bounds a
batch-8192 step: SURVEY.md section 8d); the learning rate lives in device memory so a schedule
does not need a re-capture.
"""
import ctypes

import numpy as np
import torch

from easyrec_b200 import _lib
from easyrec_b200 import kernels as K
from easyrec_b200 import layers as L


class FlatDenseOptimizer(object):
  """All dense parameters as views of one flat fp32 buffer; gradients are gathered into a flat
  buffer (one multi-tensor copy) and applied by ONE er_dense_apply launch.

  kind 'adagrad': tf.train.AdagradOptimizer (acc0 = 0.1, protos/optimizer.proto:79)
  kind 'adam' / 'lazy_adam': tf.train.AdamOptimizer dense rule, lr_t = lr*sqrt(1-b2^t)/(1-b1^t)
  l2 per tensor: kernel_regularizer l2_regularizer(scale) applied as g += scale * w."""

  def __init__(self, named_params, kind, lr, l2_of=None, beta1=0.9, beta2=0.999, eps=1e-8,
               adagrad_init=0.1):
    self.names = [n for n, _ in named_params]
    self.params = [p for _, p in named_params]
    dev = self.params[0].device
    sizes = [p.numel() for p in self.params]
    self.sizes = sizes
    # every tensor starts on a 16-byte boundary so er_gemm reads the kernels in place (float4 loads)
    total = sum((n + 3) // 4 * 4 for n in sizes)
    self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
    self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
    off = 0
    segs = np.zeros(len(sizes), dtype=_lib.DENSE_SEG_DTYPE)
    self.grad_views = []
    for i, (n, p) in enumerate(named_params):
      v = self.flat_p[off:off + sizes[i]].view_as(p)
      v.copy_(p.data)
      p.data = v
      self.grad_views.append(self.flat_g[off:off + sizes[i]].view_as(p))
      # dense layers write dW straight into this slice (layers._DenseBNAct.backward) when the kernel is
      # used once per step, so the big gradients skip the gather copy
      p._er_grad_out = self.grad_views[-1] if p.dim() == 2 else None
      p._er_uses = 0
      segs[i]['offset'] = off
      segs[i]['n'] = sizes[i]
      segs[i]['l2'] = float(l2_of(n, p)) if l2_of else 0.0
      segs[i]['lr_mult'] = 1.0
      off += (sizes[i] + 3) // 4 * 4
    self.segs_dev = torch.from_numpy(segs.view(np.uint8).reshape(-1).copy()).to(dev)
    self.n_segs = len(sizes)
    self.max_n = max(sizes)
    self.kind = {'adagrad': _lib.OPT_ADAGRAD, 'adam': _lib.OPT_ADAM_ROWS, 'lazy_adam': _lib.OPT_ADAM_ROWS,
                 'sgd': _lib.OPT_SGD}[kind]
    self.s0 = self.s1 = None
    if self.kind == _lib.OPT_ADAGRAD:
      self.s0 = torch.full((total,), adagrad_init, dtype=torch.float32, device=dev)
    elif self.kind == _lib.OPT_ADAM_ROWS:
      self.s0 = torch.zeros(total, dtype=torch.float32, device=dev)
      self.s1 = torch.zeros(total, dtype=torch.float32, device=dev)
    self.b1, self.b2, self.eps = beta1, beta2, eps
    self.lr_dev = torch.tensor([float(lr)], dtype=torch.float32, device=dev)
    self.reg_loss = torch.zeros(1, dtype=torch.float32, device=dev)
    self.grad_scale = 1.0

  def set_lr(self, lr, step):
    if self.kind == _lib.OPT_ADAM_ROWS:
      t = step + 1
      lr = float(lr) * (1 - self.b2**t)**0.5 / (1 - self.b1**t)
    self.lr_dev.fill_(float(lr))

  def zero_grad(self):
    for p in self.params:
      p.grad = None
      p._er_uses = 0

  def gather_grads(self):
    """p.grad (fresh autograd tensors) -> the flat gradient buffer (one multi-tensor copy)."""
    srcs, dsts = [], []
    for p, v in zip(self.params, self.grad_views):
      if p.grad is None:
        v.zero_()
      elif p.grad.data_ptr() != v.data_ptr():   # else: already produced in place
        srcs.append(p.grad)
        dsts.append(v)
    if srcs:
      torch._foreach_copy_(dsts, srcs)

  def apply(self):
    lib = _lib.load()
    opt = K.make_opt(self.kind, 0.0, self.b1, self.b2, self.eps, grad_scale=self.grad_scale)
    self.reg_loss.zero_()
    _lib.check(
        lib.er_dense_apply(self.flat_p.data_ptr(), self.flat_g.data_ptr(), K._p(self.s0), K._p(self.s1),
                           self.segs_dev.data_ptr(), self.n_segs, self.max_n, ctypes.byref(opt),
                           self.lr_dev.data_ptr(), self.reg_loss.data_ptr(),
                           torch.cuda.current_stream().cuda_stream), 'er_dense_apply')


class Trainer(object):

  def __init__(self, model, input_layer, dense_optimizer='adagrad', lr=0.01, lr_fn=None,
               use_cuda_graph=False, world_size=1, beta1=0.9, beta2=0.999, adagrad_init=0.1):
    self.model = model
    self.input_layer = input_layer
    self.lr = lr
    self.lr_fn = lr_fn or (lambda step: lr)
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    l2 = getattr(model, 'l2_of', None)
    self.dense_opt = FlatDenseOptimizer(named, dense_optimizer, lr, l2_of=l2, beta1=beta1, beta2=beta2,
                                        adagrad_init=adagrad_init)
    self.world = world_size
    self.dp = None
    if world_size > 1:
      from easyrec_b200.distributed import DataParallel
      self.dp = DataParallel(input_layer, self.dense_opt, world_size)
    self.step = 0
    self.use_cuda_graph = use_cuda_graph
    self._graph = None
    self._graph2 = None
    self._step_pending = []
    self._static = None
    self._loss = None
    self._probs = None

  def _set_hyper(self):
    lr = self.lr_fn(self.step)
    self.input_layer.set_optimizer_step(lr, self.step)
    self.dense_opt.set_lr(lr, self.step)

  # The step is three segments; only the middle one talks to other ranks, so with world > 1 the
  # CUDA graph is captured as two graphs around eager NCCL calls.
  def _segment_compute(self, features, labels):
    """lookup -> model -> loss -> backward -> dense grads into the flat buffer."""
    self.dense_opt.zero_grad()
    logits = self.model(features)
    loss, probs = self.model.loss(logits, labels)
    with L.defer_dw_join():   # kernel-gradient GEMMs overlap the rest of the backward chain
      loss.backward()
    self.dense_opt.gather_grads()
    self._step_pending = list(self.input_layer._pending)
    return loss.detach(), probs

  def _segment_exchange(self):
    if self.dp is not None:
      self.dp.exchange(self._step_pending)   # flat all-reduce + all-gather of K7 inputs
      self.dp.join_presort()

  def _segment_update(self, loss):
    if self.dp is not None:
      self.dp.apply_sparse(self._step_pending, self.input_layer.opt_holder['opt'])
      self.input_layer._pending = []
    else:
      self.input_layer.backward_update()   # K7: dedup + fused row update, on this thread/stream
    self.dense_opt.apply()                 # one launch: l2 + adagrad/adam over the flat buffer
    # reported loss = data loss + embedding regularisation (autograd) + dense l2 (from the apply)
    return loss + self.dense_opt.reg_loss[0]

  def _segment_pre(self, features):
    if self.dp is not None:
      self.dp.pre_exchange(features)   # eager: K1 + all-gather of rows + global dedup sort on a side stream

  def _step_body(self, features, labels):
    self._segment_pre(features)
    loss, probs = self._segment_compute(features, labels)
    self._segment_exchange()
    return self._segment_update(loss), probs

  def train_step(self, features, labels):
    """features/labels: device tensors.  Returns (loss [scalar tensor], probs [B])."""
    self.model.train()
    self._set_hyper()
    if not self.use_cuda_graph:
      out = self._step_body(features, labels)
      self.step += 1
      return out
    if self._graph is None:
      self._capture(features, labels)
    else:
      for k, v in features.items():
        self._static[k].copy_(v, non_blocking=True)
      self._static['__labels'].copy_(labels, non_blocking=True)
    if self._graph2 is not None:
      self._segment_pre(self._static_feats)
    self._graph.replay()
    if self._graph2 is not None:   # world > 1: collectives between the two captured segments
      self._segment_exchange()
      self._graph2.replay()
    self.step += 1
    return self._loss, self._probs

  def _capture(self, features, labels):
    # The fused row update reads its hyper-parameters from kernel arguments, which a CUDA graph
    # freezes; graphs are therefore only used with a constant learning rate and adagrad/sgd rows
    # (the dense optimizer reads lr from device memory and follows any schedule).
    kind = next(iter(self.input_layer.arenas.values())).opt_kind
    if kind not in (_lib.OPT_ADAGRAD, _lib.OPT_SGD):
      raise _lib.ErError('CUDA-graph capture needs step-invariant row-update arguments '
                         '(adagrad/sgd); adam rows carry beta powers per step')
    self._static = {k: v.clone() for k, v in features.items()}
    self._static['__labels'] = labels.clone()
    feats = {k: self._static[k] for k in features}
    self._static_feats = feats
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      for _ in range(2):  # warm-up on the side stream (allocator, cuBLAS handles)
        self._step_body(feats, self._static['__labels'])
    torch.cuda.current_stream().wait_stream(s)
    self._graph = torch.cuda.CUDAGraph()
    n0 = _lib.load().er_launch_count()
    if self.dp is None:
      with torch.cuda.graph(self._graph):
        self._loss, self._probs = self._step_body(feats, self._static['__labels'])
    else:
      self._segment_pre(feats)   # eager, before the capture: the captured lookup reuses these rows
      with torch.cuda.graph(self._graph):
        loss, self._probs = self._segment_compute(feats, self._static['__labels'])
      self._segment_exchange()   # eager: NCCL stays out of the capture
      self._graph2 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self._graph2, pool=self._graph.pool()):
        self._loss = self._segment_update(loss)
    # kernels of liber_b200.so inside one replay of the graph(s)
    self.launches_per_step = int(_lib.load().er_launch_count() - n0)
