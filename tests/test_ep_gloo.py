"""CPU, world 2 and 8 over gloo, kernels replaced by the oracle-backed doubles: EmbeddingParallel through the product surface
(EasyRecEstimator with train_distribute: EmbeddingParallelStrategy -> row-sharded arenas, ShardedLookup all-to-all
around every lookup, gradient rows to the owners, 1/N gradient scale, dense all-reduce) trains the same model as
replicated data parallel on the same per-rank batches (compat/feature_column/feature_column.py:248-357,
compat/optimizers.py:294-345)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, port, ret, world):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  import ep_helpers
  import host_doubles
  host_doubles.install_all()
  torch.use_deterministic_algorithms(True)
  torch.utils.deterministic.fill_uninitialized_memory = True
  from easyrec_b200.estimator import EasyRecEstimator

  def make(cfg, ep):
    return EasyRecEstimator(cfg, device='cpu', seed=5, world_size=world, rank=rank, embedding_parallel=ep)
  ret[rank] = ep_helpers.run(make, 'cpu', rank, world)
  # ... and with the next batch's id exchange prefetched beside the current step (train_step(next_features=...))
  ret[rank] = max(ret[rank], ep_helpers.run(make, 'cpu', rank, world, steps=5, lookahead=True))
  dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('world', [2, 8])   # (8 = the scaling bench's largest run: 8 owners, per-peer blocks, rank-major sums)
def test_embedding_parallel_equals_replicated_data_parallel_gloo(world):
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(_free_port(), ret, world), nprocs=world, join=True)
  assert len(ret) == world
