"""Warm, in-graph per-kernel durations of the C2 DeepFM training step (CUPTI via torch.profiler; unlike an
ncu launch list the kernels run back to back with warm caches, as in the timed bench)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from easyrec_b200 import workloads
from easyrec_b200.trainer import Trainer

torch.backends.cuda.matmul.allow_tf32 = False
B, V = 8192, int(os.environ.get('VOCAB', 10_000_000))
dev = 'cuda:0'
graph = os.environ.get('GRAPH', '1') == '1'
il, model = workloads.build_deepfm_criteo(B, V, dev, seed=20240)
tr = Trainer(model, il, 'adagrad', lr=0.01, use_cuda_graph=graph)
host = [workloads.criteo_batch(B, 20240 + i) for i in range(8)]
devb = [({'sparse_fea': torch.from_numpy(a).to(dev), 'dense_fea': torch.from_numpy(b).to(dev)},
         torch.from_numpy(c).to(dev)) for a, b, c in host]
for i in range(10):
  tr.train_step(*devb[i % 8])
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  for i in range(N):
    tr.train_step(*devb[i % 8])
  torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for e in prof.events():
  if e.device_type == torch.autograd.DeviceType.CUDA:
    tot[e.name] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
    cnt[e.name] += 1
total = sum(tot.values())
print('sum of kernel time %.1f us/step over %.1f launches/step (graph=%s)' % (total / N, sum(cnt.values()) / N, graph))
for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(os.environ.get('TOP', 60))]:
  print('%8.1f us/step %5.1f%% %5.1f x/step %7.1f us each  %s' % (v / N, 100 * v / total, cnt[name] / N, v / cnt[name], name[:100]))
