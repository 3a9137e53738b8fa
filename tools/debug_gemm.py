"""Layout probe for er_gemm: A = identity picks out B's elements, B = identity picks out A's."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyrec_b200 import kernels as K


def run(M, N, Kd, a_mn, b_mn, probe):
  dev = 'cuda'
  if probe == 'B':   # A = I (M x K), B[k,n] = code
    A = torch.zeros(M, Kd, device=dev)
    idx = torch.arange(min(M, Kd), device=dev)
    A[idx, idx] = 1.0
    B = (torch.arange(Kd, device=dev)[:, None] * 1.0 + torch.arange(N, device=dev)[None, :] / 1024.0).float()
  else:
    B = torch.zeros(Kd, N, device=dev)
    idx = torch.arange(min(N, Kd), device=dev)
    B[idx, idx] = 1.0
    A = (torch.arange(M, device=dev)[:, None] * 1.0 + torch.arange(Kd, device=dev)[None, :] / 1024.0).float()
  a = A.t().contiguous().t() if a_mn else A.contiguous()        # a_mn: m contiguous
  b = B.contiguous() if b_mn else B.t().contiguous().t()        # b_mn: n contiguous
  C = K.gemm(a, b)
  ref = A.double() @ B.double()
  err = (C.double() - ref).abs().max().item()
  tag = 'M%d N%d K%d a_mn=%d b_mn=%d probe=%s' % (M, N, Kd, a_mn, b_mn, probe)
  print('%-48s max err %.3e %s' % (tag, err, 'OK' if err < 1e-5 else 'BAD'))
  if err >= 1e-5:
    c = C[:4, :8].cpu()
    print('   got (int part = row idx of probed operand, frac*1024 = col idx):')
    for r in range(4):
      print('   ', ' '.join('%9.4f' % v for v in c[r].tolist()), '  | want', ' '.join('%9.4f' % v for v in ref[r, :8].tolist()))
    bad = ((C.double() - ref).abs() > 1e-5).nonzero()
    print('   n_bad %d of %d; first bad idx %s; rows bad: %s cols bad: %s' % (
        bad.shape[0], C.numel(), bad[0].tolist(), sorted(set(bad[:, 0].tolist()))[:12], sorted(set(bad[:, 1].tolist()))[:12]))


for shape in [(32, 32, 8), (32, 32, 32), (128, 128, 32), (128, 128, 128), (200, 72, 136)]:
  for a_mn in (0, 1):
    for b_mn in (0, 1):
      for probe in ('B', 'A'):
        run(shape[0], shape[1], shape[2], a_mn, b_mn, probe)
torch.cuda.synchronize()
