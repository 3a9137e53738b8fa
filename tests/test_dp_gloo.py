"""CPU, world_size 2 and 8 over gloo: the host logic of data-parallel training (easyrec_b200/distributed.py).

The CUDA kernel itself cannot run here; what is checked is everything around it: the replicated slot
plan (GlobalCall), the all-gathered K7 inputs, the 1/world scaling and the dense flat all-reduce.  The
gathered buffers are interpreted through the global slot plan and fed to the CPU oracle; the result must
equal the oracle run on the concatenated global batch.
"""
import os
import socket

import numpy as np
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


def _gseg_from_plan(slots_np, grads, dim):
  n_seg = int((slots_np['seg_begin'] + slots_np['n_seg']).max())
  out = np.zeros((n_seg, dim), np.float32)
  for s in slots_np:
    buf = grads[int(s['out_buf'])].reshape(-1)
    for k in range(int(s['n_seg'])):
      o = k * int(s['out_stride']) + int(s['out_col'])
      out[int(s['seg_begin']) + k] = buf[o:o + dim]
  return out


def _worker(rank, port, ret, WORLD):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=WORLD)
  from easyrec_b200 import _lib, embedding as E
  from easyrec_b200.distributed import DataParallel
  from oracle import oracle as O
  B, F, D, V = 16, 3, 4, 50
  arena = E.Arena(D, 'cpu')
  arena.add_table('t0', 30)
  arena.add_table('t1', 20)
  arena.materialize(_lib.OPT_ADAGRAD, generator=torch.Generator().manual_seed(5))
  slots = [E.Slot('a', 't0', _lib.BUCKET_NONE, 30), E.Slot('b', 't1', _lib.BUCKET_NONE, 20),
           E.Slot('c', 't0', _lib.BUCKET_NONE, 30, out_buf=1)]
  call = E.ArenaCall(arena, slots, B, [2 * D, D], single_valued=True)

  class FakeIL(object):
    calls = {D: call}
    _pending = []

  from easyrec_b200.trainer import FlatDenseOptimizer
  w1 = torch.nn.Parameter(torch.ones(5, 3) * (rank + 1))
  w2 = torch.nn.Parameter(torch.ones(7) * (rank + 2))
  dopt = FlatDenseOptimizer([('w1', w1), ('w2', w2)], 'adagrad', 0.01)
  assert w1.data_ptr() == dopt.flat_p.data_ptr()  # parameters are views of the flat buffer
  dp = DataParallel(FakeIL, dopt, WORLD)
  # ---- dense: mean over replicas = all-reduce(sum) x grad_scale ----
  w1.grad = torch.full((5, 3), float(rank + 1))
  w2.grad = torch.full((7,), float(10 * (rank + 1)))
  dopt.gather_grads()
  dp.sync_dense_grads()
  # (each tensor starts on a 16-byte boundary of the flat buffer; the padding stays zero)
  mean_rank = (WORLD + 1) / 2.0   # mean of rank + 1 over the replicas
  assert torch.allclose(dopt.grad_views[0] * dopt.grad_scale, torch.full((5, 3), mean_rank))
  assert torch.allclose(dopt.grad_views[1] * dopt.grad_scale, torch.full((7,), 10.0 * mean_rank))
  assert dopt.grad_views[1].data_ptr() % 16 == dopt.flat_g.data_ptr() % 16 and float(dopt.flat_g[15]) == 0.0
  # ---- sparse: gathered inputs through the replicated slot plan == global batch ----
  rng = np.random.default_rng(100 + rank)
  offs = np.repeat(np.array([0, 30, 0]), B)
  rows = torch.from_numpy(offs + np.concatenate([rng.integers(0, 30, B), rng.integers(0, 20, B),
                                                 rng.integers(0, 30, B)]).astype(np.int64))
  w = torch.from_numpy(rng.uniform(0.5, 1.5, F * B).astype(np.float32))
  outs = call.alloc_outputs()
  for o in outs:
    o.grad = torch.from_numpy(rng.normal(size=tuple(o.shape)).astype(np.float32))
  g = dp.gather_sparse(call, rows, w, outs)
  gseg = _gseg_from_plan(g.slots_np, [x.numpy() for x in g.grads], D)
  t_a, acc_a = arena.weight.numpy().copy(), arena.state0.numpy().copy()
  O.embedding_bwd(t_a, acc_a, None, g.rows.numpy(), None, gseg, O.OPT_ADAGRAD, 0.05, weights=g.weights.numpy(),
                  grad_scale=1.0 / WORLD)
  # reference: every rank's batch concatenated by hand, plain per-rank slot plans
  all_rows = [torch.empty_like(rows) for _ in range(WORLD)]
  all_w = [torch.empty_like(w) for _ in range(WORLD)]
  dist.all_gather(all_rows, rows)
  dist.all_gather(all_w, w)
  gsegs = []
  for r in range(WORLD):
    per = []
    for i, o in enumerate(outs):
      lst = [torch.empty_like(o.grad) for _ in range(WORLD)]
      dist.all_gather(lst, o.grad)
      per.append(lst[r].numpy())
    gsegs.append(_gseg_from_plan(call.slots_np, per, D))
  t_b, acc_b = arena.weight.numpy().copy(), arena.state0.numpy().copy()
  O.embedding_bwd(t_b, acc_b, None, np.concatenate([x.numpy() for x in all_rows]), None, np.concatenate(gsegs),
                  O.OPT_ADAGRAD, 0.05, weights=np.concatenate([x.numpy() for x in all_w]), grad_scale=1.0 / WORLD)
  assert np.array_equal(t_a, t_b) and np.array_equal(acc_a, acc_b)
  assert (t_a != arena.weight.numpy()).any()
  ret[rank] = float(t_a.sum())
  dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('world', [2, 8])
def test_data_parallel_host_logic_gloo(world):
  port = _free_port()
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(port, ret, world), nprocs=world, join=True)
  assert len(ret) == world and len(set(ret.values())) == 1  # replicas agree bit for bit
