"""2 GPUs over NCCL: EmbeddingParallel through the product surface (row-sharded arenas + ShardedLookup all-to-all, real
kernels) trains the same model as replicated data parallel on the same per-rank batches - see tests/ep_helpers.py.
Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
WORLD = 2


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, port, ret):
  import torch.distributed as dist
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dev = 'cuda:%d' % rank
  dist.init_process_group('nccl', rank=rank, world_size=WORLD, device_id=torch.device(dev))
  torch.backends.cuda.matmul.allow_tf32 = False
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  import ep_helpers
  from easyrec_b200.estimator import EasyRecEstimator

  def make(cfg, ep):
    return EasyRecEstimator(cfg, device=dev, seed=5, world_size=WORLD, rank=rank, embedding_parallel=ep)
  ret[rank] = ep_helpers.run(make, dev, rank, WORLD, steps=5, atol=5e-6)
  dist.barrier()
  os._exit(0)


@pytest.mark.timeout(400)
def test_embedding_parallel_equals_replicated_data_parallel_on_2_gpus():
  if torch.cuda.device_count() < WORLD:
    pytest.skip('needs %d GPUs' % WORLD)
  import torch.multiprocessing as mp
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(_free_port(), ret), nprocs=WORLD, join=True)
  assert len(ret) == WORLD
