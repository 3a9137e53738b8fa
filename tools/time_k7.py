"""K7 timing on the C2 batch: whole er_embedding_bwd (sort + sums + row update) and the part after the sort."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyrec_b200 import _lib, kernels as K, workloads
B, V, DIM = 8192, 10_000_000, 16
dev = 'cuda:0'
il, model = workloads.build_deepfm_criteo(B, V, dev, seed=20240)
call, arena = il.calls[DIM], il.arenas[DIM]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
gout = torch.randn(B, call.out_strides[0], device=dev) * 1e-3
rows_l, w_l = [], []
for i in range(4):
  ids, dense, _ = workloads.criteo_batch(B, 20240 + i)
  feats = {'sparse_fea': torch.from_numpy(ids).to(dev), 'dense_fea': torch.from_numpy(dense).to(dev)}
  cids, w = il._gather_inputs(DIM, feats['sparse_fea'], il.normalize_dense(feats['dense_fea']))
  rows_l.append(K.bucketize(cids, call.slots_dev, call.n_slots, call.n_seg).clone())
  w_l.append(w.clone())
opt = K.make_opt(_lib.OPT_ADAGRAD, 0.01)
st = torch.cuda.current_stream()


def timeit(fn, pre=None, iters=20):
  tot = 0.0
  for it in range(iters):
    flush.fill_(float(it))
    if pre:
      pre(it)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); fn(it); e1.record(st); e1.synchronize()
    tot += e0.elapsed_time(e1)
  return tot / iters * 1e3


def full(it):
  K.embedding_bwd(arena.weight, arena.state0, None, DIM, rows_l[it % 4], call.slots_dev, call.n_slots, call.n_seg,
                  [gout], opt, call.ws, weights=w_l[it % 4])


def presort(it):
  K.embedding_bwd_presort(rows_l[it % 4], arena.n_rows, DIM, call.ws, call.slots_dev, call.n_slots)


def after_sort(it):
  K.embedding_bwd(arena.weight, arena.state0, None, DIM, rows_l[it % 4], call.slots_dev, call.n_slots, call.n_seg,
                  [gout], opt, call.ws, weights=w_l[it % 4], sorted_from=(call.ws, DIM))


for _ in range(3):
  full(0)
print('K7 whole %.1f us | sort alone %.1f us | sums + row update after the sort %.1f us' % (
    timeit(full), timeit(presort), timeit(after_sort, pre=presort)))
