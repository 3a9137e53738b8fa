"""Bisect which part of the training step breaks CUDA-graph capture (run on a GPU box)."""
import sys
import torch
sys.path.insert(0, '.')
from easyrec_b200 import workloads, kernels as K, embedding as E, _lib
from easyrec_b200.trainer import Trainer

which = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else 'global'
dev = 'cuda:0'
torch.backends.cuda.matmul.allow_tf32 = False
B, V = 8192, 1000003
il, model = workloads.build_deepfm_criteo(B, V, dev)
tr = Trainer(model, il, 'adagrad', lr=0.01)
ids, dense, labels = workloads.criteo_batch(B, 1)
feats = {'sparse_fea': torch.from_numpy(ids).to(dev), 'dense_fea': torch.from_numpy(dense).to(dev)}
lab = torch.from_numpy(labels).to(dev)
tr._set_hyper()
model.train()


def body():
  if which == 'lookup':
    return il.lookup(feats)['deep'][0].sum()
  if which == 'fwd':
    return model(feats).sum()
  if which == 'fwdloss':
    lg = model(feats)
    return model.loss(lg, lab)[0]
  if which == 'bwd':
    for p in tr.dense_opt.params:
      p.grad = None
    lg = model(feats)
    l = model.loss(lg, lab)[0]
    l.backward()
    il.backward_update()
    return l.detach()
  if which == 'bwd_noemb':
    x = torch.randn(B, 624, device=dev, requires_grad=True)
    y = model.dnn(x).sum()
    y.backward()
    return y.detach()
  if which == 'full':
    return tr._step_body(feats, lab)[0]


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
  for _ in range(3):
    body()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
  with torch.cuda.graph(g, capture_error_mode=mode):
    out = body()
  g.replay()
  torch.cuda.synchronize()
  print('CAPTURE OK', which, mode, float(out))
except Exception as e:
  print('CAPTURE FAIL', which, mode, str(e).split('\n')[0])
