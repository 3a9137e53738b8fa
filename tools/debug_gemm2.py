import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyrec_b200 import kernels as K
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device='cuda').manual_seed(0)


def run(M, N, Kd, a_mn, b_mn):
  A = torch.randn(M, Kd, device='cuda', generator=g)
  B = torch.randn(Kd, N, device='cuda', generator=g)
  a = A.t().contiguous().t() if a_mn else A.contiguous()
  b = B.contiguous() if b_mn else B.t().contiguous().t()
  C = K.gemm(a, b)
  ref = A.double() @ B.double()
  e = (C.double() - ref).abs()
  sc = float(ref.abs().mean())
  bad = (e > 1e-3 * sc).nonzero()
  msg = ''
  if bad.numel():
    msg = ' BAD n=%d rows %s cols %s' % (bad.shape[0], sorted(set(bad[:, 0].tolist()))[:8], sorted(set(bad[:, 1].tolist()))[:8])
  print('M%5d N%4d K%5d a_mn=%d b_mn=%d  rms %.2e max %.2e (rel to mean|ref|)%s' % (
      M, N, Kd, a_mn, b_mn, float(e.pow(2).mean().sqrt()) / sc, float(e.max()) / sc, msg), flush=True)


for shape in [(128, 64, 8192), (128, 64, 256), (128, 64, 512), (128, 64, 2048), (128, 128, 8192), (256, 64, 8192),
              (128, 32, 8192), (128, 96, 4096)]:
  run(shape[0], shape[1], shape[2], 1, 1)
for shape in [(8192, 128, 64), (8192, 128, 32), (8192, 128, 96), (8192, 128, 128), (8192, 256, 64), (1024, 64, 64)]:
  run(shape[0], shape[1], shape[2], 0, 0)
  run(shape[0], shape[1], shape[2], 0, 1)
