"""CPU, world_size 2 and 4 over gloo: a config-built model trained data-parallel through Trainer(world_size=N)
with the kernel doubles of tests/host_doubles.py - flat dense all-reduce, all-gather of every arena's K7 inputs,
the same fused update on every rank.  Replicas see different batches and must stay bit-identical (tables,
optimizer slots, dense parameters), and the loss must go down."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, port, ret, world):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  sys.path.insert(0, HERE)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  import host_doubles
  host_doubles.install_all()
  torch.use_deterministic_algorithms(True)
  torch.utils.deterministic.fill_uninitialized_memory = True
  import test_gpu_models as G
  from easyrec_b200 import builder, trainer as T
  from easyrec_b200.config import config_util
  from easyrec_b200.input import readers
  cfg = config_util.get_configs_from_pipeline_file(G.BACKBONE_DCN_CFG.encode())
  B = 16
  il, model, opt = builder.build_model(cfg, B, 'cpu', cpu_generator=torch.Generator().manual_seed(1))   # replicated tables
  tr = T.Trainer(model, il, 'adagrad', lr_fn=opt['lr_fn'], world_size=world)
  feats, labels = readers.DummyInput(il, seed=100 + rank).batch()      # a different batch on every rank
  feats['sparse_fea'] = feats['sparse_fea'] % 37
  losses = [float(tr.train_step(feats, labels)[0]) for _ in range(10)]
  assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
  digest = [float(a.storage.double().sum()) for a in il.arenas.values()]
  digest += [float(p.detach().double().sum()) for p in model.parameters()]          # views of the flat buffer
  digest += [float(v.double().abs().sum()) for v in tr.dense_opt.grad_views]        # the all-reduced gradients
  ret[rank] = tuple(digest)
  dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('world', [2, 4])
def test_data_parallel_model_training_keeps_replicas_identical(world):
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(_free_port(), ret, world), nprocs=world, join=True)
  assert len(ret) == world and len(set(ret.values())) == 1, dict(ret)
