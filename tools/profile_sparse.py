"""Sparse kernels only (C2 shape), for ncu launch lists / full captures.
usage: python tools/profile_sparse.py [iters] [vocab] [uniform]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from easyrec_b200 import _lib, kernels as K, workloads

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
uniform = len(sys.argv) > 3 and sys.argv[3] == 'uniform'
dev = 'cuda:0'
B = 8192
il, model = workloads.build_deepfm_criteo(B, vocab, dev)
call, arena = il.calls[16], il.arenas[16]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
gout = torch.randn(B, call.out_strides[0], device=dev) * 1e-3
outs = call.alloc_outputs()
opt = K.make_opt(_lib.OPT_ADAGRAD, 0.01)
for it in range(iters):
  ids, dense, labels = workloads.criteo_batch(B, 100 + it, uniform=uniform)
  feats = {'sparse_fea': torch.from_numpy(ids).to(dev), 'dense_fea': torch.from_numpy(dense).to(dev)}
  dn = il.normalize_dense(feats['dense_fea'])
  cids, w = il._gather_inputs(16, feats['sparse_fea'], dn)
  flush.fill_(1.0)
  torch.cuda.synchronize()
  rows = K.bucketize(cids, call.slots_dev, call.n_slots, call.n_seg)
  K.embedding_fwd(arena.weight, 16, rows, call.slots_dev, call.n_slots, call.n_seg, outs, weights=w)
  flush.fill_(2.0)
  K.embedding_bwd(arena.weight, arena.state0, None, 16, rows, call.slots_dev, call.n_slots, call.n_seg,
                  [gout], opt, call.ws, weights=w)
  torch.cuda.synchronize()
print('done', int(torch.unique(rows).numel()), 'unique rows of', rows.numel())
