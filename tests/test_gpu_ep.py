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

  def make(cfg, ep, graph=False):
    return EasyRecEstimator(cfg, device=dev, seed=5, world_size=WORLD, rank=rank, embedding_parallel=ep,
                            use_cuda_graph=graph and ep is not False)
  ret[rank] = ep_helpers.run(make, dev, rank, WORLD, steps=5, atol=5e-6)
  # and with the whole step - the three all-to-alls included - replayed from one CUDA graph (two eager steps, capture,
  # replays), against the eager replicated model
  worst = ep_helpers.run(lambda cfg, ep: make(cfg, ep, graph=True), dev, rank, WORLD, steps=7, atol=5e-6)
  ret[rank] = max(ret[rank], worst)
  # the id exchange of the next batch prefetched beside the current step: eager, and replayed from the graph
  for graph in (False, True):
    worst = ep_helpers.run(lambda cfg, ep: make(cfg, ep, graph=graph), dev, rank, WORLD, steps=8, atol=5e-6, lookahead=True)
    ret[rank] = max(ret[rank], worst)
  dist.barrier()
  os._exit(0)


def test_shard_group_kernel_groups_distinct_rows_by_owner():
  """er_shard_group (K8): every lookup's position holds its row inside its owner's block, each (owner, row) pair
  appears once, counts are the distinct rows per owner; a block that is too small reports the lost lookups."""
  import numpy as np
  from easyrec_b200 import kernels as K
  dev = 'cuda:0'
  rng = np.random.default_rng(3)
  for n, world, cap in ((5000, 4, 2048), (212992, 8, 40000), (3000, 2, 64)):
    rows = (rng.zipf(1.2, n) % 100000).astype(np.int64)
    owner = (rows % world).astype(np.int32)
    local = rows // world
    local[rng.uniform(size=n) < 0.03] = -1            # dropped lookups
    t_rows, t_owner = torch.from_numpy(local).to(dev), torch.from_numpy(owner).to(dev)
    send = torch.empty(world * cap, dtype=torch.int64, device=dev)
    pos = torch.empty(n, dtype=torch.int64, device=dev)
    counts = torch.empty(world + 1, dtype=torch.int32, device=dev)
    ws = K.shard_group_workspace(n, dev)
    for _ in range(2):   # the workspace is re-initialised by every call
      K.shard_group(t_rows, t_owner, world, cap, send, pos, counts, ws)
    send, pos, counts = send.cpu().numpy(), pos.cpu().numpy(), counts.cpu().numpy()
    live = local >= 0
    want_counts = np.array([np.unique(local[live & (owner == o)]).size for o in range(world)])
    assert np.array_equal(counts[:world], want_counts)
    assert (pos[~live] == -1).all()
    kept = live & (pos >= 0)
    assert np.array_equal(send[pos[kept]], local[kept]) and np.array_equal(pos[kept] // cap, owner[kept])
    for o in range(world):
      blk = send[o * cap:(o + 1) * cap]
      k = min(int(counts[o]), cap)
      assert (blk[k:] == -1).all() and np.unique(blk[:k]).size == k
    if (want_counts <= cap).all():
      assert counts[world] == 0 and kept.sum() == live.sum()
    else:
      assert counts[world] == (live & (pos < 0)).sum() > 0


@pytest.mark.timeout(400)
def test_embedding_parallel_equals_replicated_data_parallel_on_2_gpus():
  if torch.cuda.device_count() < WORLD:
    pytest.skip('needs %d GPUs' % WORLD)
  import torch.multiprocessing as mp
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(_free_port(), ret), nprocs=WORLD, join=True)
  assert len(ret) == WORLD
