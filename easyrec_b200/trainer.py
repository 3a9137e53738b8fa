"""Training step driver: the torch-side replacement of the reference's hot loop
(`sess.run(train_op)` under EasyRecEstimator._train_model_fn, model/easy_rec_estimator.py:155-472,
and optimize_loss, compat/optimizers.py:89-450).

One step = K1 bucketize -> K2 gather+pool -> interaction + dense MLP -> loss -> backward ->
K7 dedup + fused row update (inside backward) -> dense optimizer.  The whole step is
shape-static, so it can be captured once into a CUDA graph and replayed (launch latency, not
HBM, bounds a batch-8192 step: SURVEY.md section 8d).
"""
import torch

from easyrec_b200 import _lib


class TFAdagrad(torch.optim.Optimizer):
  """tf.train.AdagradOptimizer (dense apply): acc += g^2; w -= lr * g * rsqrt(acc); acc0 = 0.1
  (protos/optimizer.proto:79).  foreach implementation, capturable (lr lives on the device)."""

  def __init__(self, params, lr, initial_accumulator_value=0.1):
    super().__init__(params, dict(lr=lr))
    self.params = [p for g in self.param_groups for p in g['params']]
    self.acc = [torch.full_like(p, initial_accumulator_value) for p in self.params]
    dev = self.params[0].device
    self.lr_t = torch.tensor(float(lr), device=dev)

  def set_lr(self, lr):
    self.lr_t.fill_(float(lr))

  @torch.no_grad()
  def step(self):
    ps = [p for p in self.params if p.grad is not None]
    if not ps:
      return
    accs = [a for p, a in zip(self.params, self.acc) if p.grad is not None]
    gs = [p.grad for p in ps]
    torch._foreach_addcmul_(accs, gs, gs, value=1.0)
    upd = torch._foreach_div(gs, torch._foreach_sqrt(accs))
    torch._foreach_mul_(upd, self.lr_t)
    torch._foreach_sub_(ps, upd)


class TFAdam(torch.optim.Optimizer):
  """tf.train.AdamOptimizer (dense apply): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
  m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t * m / (sqrt(v) + eps)."""

  def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    super().__init__(params, dict(lr=lr))
    self.params = [p for g in self.param_groups for p in g['params']]
    self.m = [torch.zeros_like(p) for p in self.params]
    self.v = [torch.zeros_like(p) for p in self.params]
    dev = self.params[0].device
    self.b1, self.b2, self.eps = beta1, beta2, eps
    self.lr_t = torch.tensor(float(lr), device=dev)   # already bias-corrected by set_lr
    self.t = 0
    self.base_lr = lr
    self.set_lr(lr, 0)

  def set_lr(self, lr, step):
    t = step + 1
    self.lr_t.fill_(float(lr) * (1 - self.b2**t)**0.5 / (1 - self.b1**t))

  @torch.no_grad()
  def step(self):
    idx = [i for i, p in enumerate(self.params) if p.grad is not None]
    if not idx:
      return
    ps = [self.params[i] for i in idx]
    gs = [p.grad for p in ps]
    ms = [self.m[i] for i in idx]
    vs = [self.v[i] for i in idx]
    torch._foreach_mul_(ms, self.b1)
    torch._foreach_add_(ms, gs, alpha=1 - self.b1)
    torch._foreach_mul_(vs, self.b2)
    torch._foreach_addcmul_(vs, gs, gs, value=1 - self.b2)
    den = torch._foreach_sqrt(vs)
    torch._foreach_add_(den, self.eps)
    upd = torch._foreach_div(ms, den)
    torch._foreach_mul_(upd, self.lr_t)
    torch._foreach_sub_(ps, upd)


class Trainer(object):

  def __init__(self, model, input_layer, dense_optimizer='adagrad', lr=0.01, lr_fn=None,
               use_cuda_graph=False, world_size=1):
    self.model = model
    self.input_layer = input_layer
    self.lr = lr
    self.lr_fn = lr_fn or (lambda step: lr)
    params = [p for p in model.parameters() if p.requires_grad]
    if dense_optimizer == 'adagrad':
      self.dense_opt = TFAdagrad(params, lr)
    elif dense_optimizer in ('adam', 'lazy_adam'):
      self.dense_opt = TFAdam(params, lr)
    else:
      raise ValueError(dense_optimizer)
    self.dp = None
    if world_size > 1:
      from easyrec_b200.distributed import DataParallel
      self.dp = DataParallel(input_layer, params, world_size)
    self.step = 0
    self.use_cuda_graph = use_cuda_graph
    self._graph = None
    self._static = None
    self._loss = None
    self._probs = None

  def _set_hyper(self):
    lr = self.lr_fn(self.step)
    self.input_layer.set_optimizer_step(lr, self.step)
    if isinstance(self.dense_opt, TFAdam):
      self.dense_opt.set_lr(lr, self.step)
    else:
      self.dense_opt.set_lr(lr)

  def _step_body(self, features, labels):
    for p in self.dense_opt.params:
      p.grad = None
    logits = self.model(features)
    loss, probs = self.model.loss(logits, labels)
    loss.backward()
    if self.dp is not None:
      self.dp.sync_dense_grads()                                       # flat all-reduce (mean)
      self.dp.sparse_backward_update(self.input_layer.opt_holder['opt'])  # all-gather + K7
    else:
      self.input_layer.backward_update()   # K7: dedup + fused row update, on this thread/stream
    self.dense_opt.step()
    return loss.detach(), probs

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
    self._graph.replay()
    self.step += 1
    return self._loss, self._probs

  def _capture(self, features, labels):
    # The fused row update reads its hyper-parameters from kernel arguments, which a CUDA graph
    # freezes; graphs are therefore only used with a constant learning rate and adagrad/sgd rows.
    kind = next(iter(self.input_layer.arenas.values())).opt_kind
    if kind not in (_lib.OPT_ADAGRAD, _lib.OPT_SGD):
      raise _lib.ErError('CUDA-graph capture needs step-invariant row-update arguments '
                         '(adagrad/sgd); adam rows carry beta powers per step')
    self._static = {k: v.clone() for k, v in features.items()}
    self._static['__labels'] = labels.clone()
    feats = {k: self._static[k] for k in features}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
      for _ in range(2):  # warm-up on the side stream (allocator, cuBLAS handles)
        self._step_body(feats, self._static['__labels'])
    torch.cuda.current_stream().wait_stream(s)
    self._graph = torch.cuda.CUDAGraph()
    n0 = _lib.load().er_launch_count()
    with torch.cuda.graph(self._graph):
      self._loss, self._probs = self._step_body(feats, self._static['__labels'])
    # kernels of liber_b200.so inside one replay of the graph
    self.launches_per_step = int(_lib.load().er_launch_count() - n0)
