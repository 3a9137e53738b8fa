"""Finer bisect of CUDA-graph capture failures: which custom op / which thread."""
import sys
import threading

import torch

sys.path.insert(0, '.')
from easyrec_b200 import embedding as E, kernels as K, workloads  # noqa: E402
from easyrec_b200.trainer import Trainer  # noqa: E402

which = sys.argv[1]
dev = 'cuda:0'
B = 8192
x0 = torch.randn(B, 624, device=dev)
gy = torch.randn(B, 16, device=dev)
lab = (torch.rand(B, device=dev) < 0.25).float()
info = {}


class Probe(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x):
    info['fwd_stream'] = torch.cuda.current_stream().cuda_stream
    return x * 2

  @staticmethod
  def backward(ctx, g):
    info['bwd_stream'] = torch.cuda.current_stream().cuda_stream
    info['bwd_thread'] = threading.current_thread().name
    return g * 2


def body():
  if which == 'fm':
    x = x0.clone().requires_grad_(True)
    y = E.fm(x, 39, 16)
    y.sum().backward()
    return x.grad.sum()
  if which == 'sigmoid':
    lg = x0[:, 0].clone().requires_grad_(True)
    l, _ = E.sigmoid_cross_entropy(lg, lab)
    l.backward()
    return lg.grad.sum()
  if which == 'probe':
    x = x0.clone().requires_grad_(True)
    Probe.apply(x).sum().backward()
    return x.grad.sum()
  if which == 'fm_main':
    gx = K.fm_bwd(x0, gy, 39, 16)
    return gx.sum()
  if which == 'fm_thread':
    out = {}
    st = torch.cuda.current_stream()

    def run():
      with torch.cuda.stream(st):
        out['g'] = K.fm_bwd(x0, gy, 39, 16)

    t = threading.Thread(target=run)
    t.start()
    t.join()
    return out['g'].sum()
  if which == 'emb_manual':
    return tr._step_body(feats, lab8)[0]


if which == 'emb_manual':
  torch.backends.cuda.matmul.allow_tf32 = False
  il, model = workloads.build_deepfm_criteo(B, 1000003, dev)
  tr = Trainer(model, il, 'adagrad', lr=0.01)
  ids, dense, labels = workloads.criteo_batch(B, 1)
  feats = {'sparse_fea': torch.from_numpy(ids).to(dev), 'dense_fea': torch.from_numpy(dense).to(dev)}
  lab8 = torch.from_numpy(labels).to(dev)
  tr._set_hyper()
  model.train()

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
  for _ in range(3):
    body()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
  with torch.cuda.graph(g):
    info['capture_stream'] = torch.cuda.current_stream().cuda_stream
    out = body()
  g.replay()
  torch.cuda.synchronize()
  print('CAPTURE OK', which, float(out), info)
except Exception as e:
  print('CAPTURE FAIL', which, str(e).split('\n')[0], info)
