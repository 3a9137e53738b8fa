"""er_gemm vs torch.mm kernel time without Python launch overhead: each op is captured into a CUDA graph of
20 back-to-back launches and the graph is replayed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyrec_b200 import kernels as K

torch.backends.cuda.matmul.allow_tf32 = False
B = int(os.environ.get('B', 8192))
shapes = [(624, 256), (256, 128), (128, 64), (84, 256)]
REP = 20


def graph_time(fn):
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
      for _ in range(REP):
        fn()
  torch.cuda.synchronize()
  for _ in range(2):
    g.replay()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    g.replay()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / (5 * REP) * 1e3


for kin, kout in shapes:
  x = torch.randn(B, kin, device='cuda')
  w = torch.randn(kin, kout, device='cuda')
  gz = torch.randn(B, kout, device='cuda')
  for name, a, b in [('fwd', x, w), ('dX ', gz, w.t()), ('dW ', x.t(), gz)]:
    out = torch.empty(a.shape[0], b.shape[1], device='cuda')
    t_er = graph_time(lambda: K.gemm(a, b, out=out))
    t_th = graph_time(lambda: torch.mm(a, b, out=out))
    fl = 2.0 * a.shape[0] * a.shape[1] * b.shape[1]
    print('%s [%5d x %4d x %5d]  er_gemm %7.1f us (%6.1f TFLOP/s fp32-equiv)   torch.mm %7.1f us'
          % (name, a.shape[0], b.shape[1], a.shape[1], t_er, fl / t_er * 1e-6, t_th), flush=True)
