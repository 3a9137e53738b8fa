"""Per-kernel durations of the multi-GPU C2 step (CUPTI via torch.profiler), rank 0 prints.
torchrun --nproc-per-node N tools/profile_dp.py          (EP=1: row-sharded tables instead of replicated ones)"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from easyrec_b200 import workloads
from easyrec_b200.estimator import EasyRecEstimator

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = 'cuda:%d' % lr
dist.init_process_group('nccl', device_id=torch.device(dev))
torch.backends.cuda.matmul.allow_tf32 = False
B, V = 8192, int(os.environ.get('VOCAB', 10_000_000))
est = EasyRecEstimator(workloads.c2_config_text(V, B), device=dev, seed=20240, use_cuda_graph=os.environ.get('GRAPH', '1') == '1',
                       world_size=world, rank=rank, embedding_parallel=os.environ.get('EP', '0') == '1')
host = [workloads.criteo_batch(B, 20240 + rank * 1000 + i) for i in range(8)]
devb = [({'sparse_fea': torch.from_numpy(a).to(dev), 'dense_fea': torch.from_numpy(b).to(dev)}, torch.from_numpy(c).to(dev))
        for a, b, c in host]
for i in range(10):
  est.trainer.train_step(*devb[i % 8])
torch.cuda.synchronize()
dist.barrier()
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for i in range(N):
    est.trainer.train_step(*devb[i % 8])
  torch.cuda.synchronize()
if rank == 0:
  tot, cnt = collections.defaultdict(float), collections.Counter()
  for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
      tot[e.name] += e.device_time
      cnt[e.name] += 1
  print('EP=%s' % os.environ.get('EP', '0'), 'world %d: sum of kernel time %.1f us/step over %.1f launches/step' % (world, sum(tot.values()) / N, sum(cnt.values()) / N))
  for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:28]:
    print('%8.1f us/step %5.1f x/step %7.1f us each  %s' % (v / N, cnt[name] / N, v / cnt[name], name[:100]))
dist.barrier()
sys.stdout.flush()
os._exit(0)
