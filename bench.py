#!/usr/bin/env python
"""bench.py -- samples/sec of one DeepFM training step on the fused sm_100a sparse path.

Contract: `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1) prints ONE JSON
line on rank 0.  A step = K1 hash/bucketize -> K2 gather+pool -> FM + MLPs -> sigmoid CE ->
backward -> K7 dedup + fused adagrad row update -> dense optimizer, on one batch of the
BASELINE.json config-2 shape (26 sparse + 13 dense, V rows x 16 fp32, batch 8192 per GPU).

  value     : samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       : samples/s through Trainer.train_step with pinned HOST batches: H2D of ids/dense/
              labels and D2H of the loss inside the timed region
  roofline  : the dominant own HBM-bound kernel (K7 run-sum + row update, else K2 gather),
              algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the CPU oracle port of the same step on a bounded sample (rank 0, N=1)

`--impl reference` times the CPU oracle port (TensorFlow, hence the real reference, cannot be
installed in this image: see DESIGN.md) with all host threads on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'samples_per_sec'
BATCH = 8192
N_SPARSE, N_DENSE, DIM = 26, 13, 16


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=2000)
  ap.add_argument('--warmup', type=int, default=50)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--vocab', type=int, default=10_000_000)
  ap.add_argument('--batch', type=int, default=BATCH)
  ap.add_argument('--uniform-ids', action='store_true')
  ap.add_argument('--no-graph', action='store_true')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--kernel-iters', type=int, default=30)
  return ap.parse_args()


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
  """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

  def __init__(self, index=0):
    super().__init__(daemon=True)
    self.index = index
    self.rows = []
    self.stop_flag = False
    self.proc = None

  def run(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
           '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        self.rows.append((time.time(), [x.strip() for x in line.split(',')]))
        if self.stop_flag:
          break
    except Exception:
      pass

  def mark(self):
    """start of the timed region: samples taken before it (warm-up) only count if none falls inside"""
    self.t_mark = time.time()

  def finish(self):
    t_end = time.time() + 0.06   # one more sampling period: the last line may still be in the pipe
    while time.time() < t_end and not any(t >= getattr(self, 't_mark', 0.0) for t, _ in self.rows):
      time.sleep(0.01)
    self.stop_flag = True
    if self.proc is not None:
      try:
        self.proc.terminate()
      except Exception:
        pass
    sm, mx, reasons = [], 0.0, set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    t_mark = getattr(self, 't_mark', 0.0)
    inside = [r for t, r in self.rows if t >= t_mark]
    for r in (inside or [r for _, r in self.rows]):
      try:
        sm.append(float(r[0]))
        mx = max(mx, float(r[1]))
        for n, v in zip(names, r[2:6]):
          if v.lower().startswith('active'):
            reasons.add(n)
      except Exception:
        continue
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def algorithmic_bytes(L, S, U, D, k):
  """fwd = L*(8+R) + S*R ; bwd+update = S*R + 8*L + U*(8 + k*R)."""
  R = 4 * D
  return L * (8 + R) + S * R, S * R + 8 * L + U * (8 + k * R)


# --------------------------------------------------------------------------------------
def cpu_step_oracle(state, ids, dense, labels, V, B, lr=0.01):
  """One training step of the same DeepFM on the CPU oracle (numpy + oracle/er_oracle.c)."""
  from easyrec_b200 import workloads
  from oracle import oracle as O
  rows_id, _ = O.bucketize(ids, 0, V, N_DENSE)
  rows = np.concatenate([np.repeat(np.arange(N_DENSE, dtype=np.int64), B), rows_id])
  mn = np.array(workloads.CRITEO_MIN, np.float32)
  mx = np.array(workloads.CRITEO_MAX, np.float32)
  dn = ((dense - mn) / (mx - mn)).astype(np.float32)
  w = np.concatenate([dn.T.reshape(-1), np.ones(N_SPARSE * B, np.float32)])
  F = N_SPARSE + N_DENSE
  rp = np.arange(F * B + 1, dtype=np.int32)
  deep_seg, _ = O.embedding_fwd(state['t16'], rows, rp, 0, weights=w)
  wide_seg, _ = O.embedding_fwd(state['t1'], rows, rp, 0, weights=w)
  deep = np.ascontiguousarray(deep_seg.reshape(F, B, DIM).transpose(1, 0, 2).reshape(B, F * DIM))
  wide = np.ascontiguousarray(wide_seg.reshape(F, B).T)
  logits, cache = O.deepfm_forward(wide, deep, F, DIM, state['params'])
  loss, _, g_logits = O.sigmoid_ce(logits, labels)
  g_wide, g_deep, grads = O.deepfm_backward(g_logits, wide, deep, F, DIM, state['params'], cache)
  g_deep = g_deep + np.float32(state['emb_reg']) * deep
  g_wide = g_wide + np.float32(state['emb_reg']) * wide
  gd = np.ascontiguousarray(g_deep.reshape(B, F, DIM).transpose(1, 0, 2).reshape(F * B, DIM))
  gw = np.ascontiguousarray(g_wide.T.reshape(F * B, 1))
  O.embedding_bwd(state['t16'], state['a16'], None, rows, None, gd, O.OPT_ADAGRAD, lr, weights=w)
  O.embedding_bwd(state['t1'], state['a1'], None, rows, None, gw, O.OPT_ADAGRAD, lr, weights=w)
  # dense adagrad (tf.train.AdagradOptimizer), incl. l2 on kernels
  for tag in ('dnn', 'final'):
    for L_, G_ in zip(state['params'][tag], grads[tag]):
      for k in ('W', 'b', 'gamma', 'beta'):
        g = G_[k] + (np.float32(state['l2']) * L_[k] if k == 'W' else 0)
        acc = state['acc'].setdefault((tag, id(L_), k), np.full_like(L_[k], 0.1))
        acc += g * g
        L_[k] -= (lr * g / np.sqrt(acc)).astype(np.float32)
  return loss


def make_cpu_state(V, seed=0):
  rng = np.random.default_rng(seed)
  F = N_SPARSE + N_DENSE

  def mk(i, o):
    lim = np.sqrt(6.0 / (i + o))
    return {'W': rng.uniform(-lim, lim, (i, o)).astype(np.float32), 'b': np.zeros(o, np.float32),
            'gamma': np.ones(o, np.float32), 'beta': np.zeros(o, np.float32)}

  params = {'dnn': [mk(F * DIM, 256), mk(256, 128), mk(128, 64)],
            'final': [mk(1 + DIM + 64, 256), mk(256, 128), mk(128, 64)],
            'out_W': rng.uniform(-0.3, 0.3, (64, 1)).astype(np.float32), 'out_b': np.zeros(1, np.float32)}
  n16 = V + N_DENSE
  return {'t16': (rng.standard_normal((n16, DIM), dtype=np.float32) * 0.0025),
          'a16': np.full((n16, DIM), 0.1, np.float32),
          't1': (rng.standard_normal((n16, 1), dtype=np.float32) * 0.01),
          'a1': np.full((n16, 1), 0.1, np.float32), 'params': params, 'acc': {}, 'l2': 1e-5,
          'emb_reg': 1e-5}


def run_cpu(args, steps, warmup, vocab):
  """CPU oracle port of the step; returns (samples/s, threads, seconds)."""
  from easyrec_b200 import workloads
  from oracle import oracle as O
  B = args.batch
  state = make_cpu_state(vocab)
  batches = [workloads.criteo_batch(B, 1000 + i, uniform=args.uniform_ids) for i in range(4)]
  # the port is given the thread count it runs fastest with on this host (more threads than the small
  # per-step matrices can use only adds fork/join cost): one step per candidate, best kept
  try:
    from threadpoolctl import threadpool_limits
  except Exception:
    threadpool_limits = None
  n_all = max(O.num_threads(), os.cpu_count() or 1)
  cands = sorted({n_all, max(1, n_all // 2), max(1, n_all // 4), min(n_all, 16), min(n_all, 8)}, reverse=True)
  cpu_step_oracle(state, *batches[0], vocab, B)   # first touch of the tables
  best, best_t = n_all, None
  for n in cands:
    O.set_num_threads(n)
    ctx = threadpool_limits(limits=n) if threadpool_limits else None
    t0 = time.perf_counter()
    cpu_step_oracle(state, *batches[1], vocab, B)
    t = time.perf_counter() - t0
    if ctx is not None:
      ctx.restore_original_limits() if hasattr(ctx, 'restore_original_limits') else ctx.unregister()
    if best_t is None or t < best_t:
      best, best_t = n, t
  O.set_num_threads(best)
  ctx = threadpool_limits(limits=best) if threadpool_limits else None
  for i in range(warmup):
    cpu_step_oracle(state, *batches[i % 4], vocab, B)
  t0 = time.perf_counter()
  for i in range(steps):
    cpu_step_oracle(state, *batches[i % 4], vocab, B)
  dt = time.perf_counter() - t0
  return B * steps / dt, best, dt


def main():
  args = parse()
  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  B = args.batch
  workload = 'deepfm_criteo_c2(26 sparse+13 dense, shared table V=%d x emb16 fp32, batch %d/GPU, %s ids)' % (
      args.vocab, B, 'uniform' if args.uniform_ids else 'zipf1.05')
  config = {'workload': workload, 'optimizer': 'adagrad(sparse rows fused in backward)+adagrad(dense)',
            'l2_flush': 'none: table+accumulator %.1f GB >> 126 MB L2, ids rotate over 16 distinct batches'
            % ((args.vocab + 13) * 17 * 4 * 2 / 1e9), 'parallelism': 'dp%d' % world}

  if args.impl == 'reference':
    if rank != 0:
      return 0
    # bounded sample: each "step" is one full batch-8192 training step on the CPU oracle port
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    v, threads, dt = run_cpu(args, steps, warm, args.vocab)
    line = {'metric': METRIC, 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warm, 'ms_per_step': 1000.0 * dt / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'impl': 'reference', 'config': config,
            'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                             'sample': '%d full steps of batch %d (CPU oracle: C sparse path with OpenMP + '
                                       'numpy/BLAS dense, thread count picked as the fastest of all/half/quarter/16/8 '
                                       'host threads); TensorFlow is not installable here so the TF graph itself is '
                                       'not what runs' % (steps, B)},
            'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))
    return 0

  import torch
  import torch.distributed as dist
  from easyrec_b200 import _lib, kernels as K, workloads
  from easyrec_b200.trainer import Trainer
  torch.backends.cuda.matmul.allow_tf32 = False
  torch.backends.cudnn.allow_tf32 = False
  torch.cuda.set_device(local_rank)
  dev = 'cuda:%d' % local_rank
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(dev))
  lib = _lib.load()

  il, model = workloads.build_deepfm_criteo(B, args.vocab, dev, seed=20240)
  trainer = Trainer(model, il, 'adagrad', lr=0.01, use_cuda_graph=not args.no_graph, world_size=world)
  n_rot = 16
  host = [workloads.criteo_batch(B, 20240 + rank * 1000 + i, uniform=args.uniform_ids) for i in range(n_rot)]
  pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory(), torch.from_numpy(c).pin_memory())
            for a, b, c in host]
  devb = [({'sparse_fea': a.to(dev), 'dense_fea': b.to(dev)}, c.to(dev)) for a, b, c in pinned]

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident throughput -----------------------------------------------------
  sampler = ClockSampler(local_rank)   # started before the warm-up so nvidia-smi is already sampling
  if rank == 0:
    sampler.start()
  for i in range(max(args.warmup, 3)):
    trainer.train_step(*devb[i % n_rot])
  barrier()
  n0 = lib.er_launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  sampler.mark()
  ev0.record()
  for i in range(args.steps):
    loss, _ = trainer.train_step(*devb[i % n_rot])
  ev1.record()
  barrier()
  ms = ev0.elapsed_time(ev1)
  clocks = sampler.finish() if rank == 0 else None
  if world > 1:
    tms = torch.tensor([ms], device=dev)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
  launches_host = lib.er_launch_count() - n0
  per_step_launches = getattr(trainer, 'launches_per_step', None)
  if per_step_launches is None:
    per_step_launches = launches_host // max(args.steps, 1)
  value = world * B * args.steps / (ms / 1000.0)

  # ---- end to end: pinned host batch -> H2D -> step -> loss D2H -------------------------
  e2e_steps = args.steps
  sfeat = {'sparse_fea': torch.empty_like(devb[0][0]['sparse_fea']),
           'dense_fea': torch.empty_like(devb[0][0]['dense_fea'])}
  slab = torch.empty_like(devb[0][1])
  for i in range(3):
    a, b, c = pinned[i % n_rot]
    sfeat['sparse_fea'].copy_(a, non_blocking=True)
    sfeat['dense_fea'].copy_(b, non_blocking=True)
    slab.copy_(c, non_blocking=True)
    float(trainer.train_step(sfeat, slab)[0])
  barrier()
  ev0.record()
  for i in range(e2e_steps):
    a, b, c = pinned[i % n_rot]
    sfeat['sparse_fea'].copy_(a, non_blocking=True)
    sfeat['dense_fea'].copy_(b, non_blocking=True)
    slab.copy_(c, non_blocking=True)
    loss, _ = trainer.train_step(sfeat, slab)
    lv = float(loss)  # device -> host read of the step's loss
  ev1.record()
  barrier()
  e2e_ms = ev0.elapsed_time(ev1)
  if world > 1:
    tms = torch.tensor([e2e_ms], device=dev)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_ms = float(tms.item())
  e2e_value = world * B * e2e_steps / (e2e_ms / 1000.0)
  h2d = pinned[0][0].numel() * 8 + pinned[0][1].numel() * 4 + pinned[0][2].numel() * 4

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return 0

  # ---- per-kernel roofline (own kernels, CUDA events on the launching stream) -----------
  peak, peak_src = peaks()
  call = il.calls[DIM]
  arena = il.arenas[DIM]
  F = N_SPARSE + N_DENSE
  L = S = F * B
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
  gout = torch.randn(B, call.out_strides[0], device=dev) * 1e-3
  outs = call.alloc_outputs()
  rows_list, w_list, uniq = [], [], []
  for i in range(4):
    feats, _ = devb[i]
    dn = il.normalize_dense(feats['dense_fea'])
    cids, w = il._gather_inputs(DIM, feats['sparse_fea'], dn)
    rows = K.bucketize(cids, call.slots_dev, call.n_slots, call.n_seg)
    rows_list.append(rows.clone())
    w_list.append(w.clone())
    uniq.append(int(torch.unique(rows).numel()))
  opt = K.make_opt(_lib.OPT_ADAGRAD, 0.01)
  st = torch.cuda.current_stream()

  def time_kernel(fn, iters):
    tot = 0.0
    for it in range(iters):
      flush.fill_(float(it))  # evict L2 between timed launches
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(st)
      fn(it)
      e1.record(st)
      e1.synchronize()
      tot += e0.elapsed_time(e1)
    return tot / iters

  def run_fwd(it):
    K.embedding_fwd(arena.weight, DIM, rows_list[it % 4], call.slots_dev, call.n_slots, call.n_seg, outs,
                    weights=w_list[it % 4])

  def run_bwd(it):
    K.embedding_bwd(arena.weight, arena.state0, None, DIM, rows_list[it % 4], call.slots_dev, call.n_slots,
                    call.n_seg, [gout], opt, call.ws, weights=w_list[it % 4])

  def run_hash(it):
    K.bucketize(devb[it % 4][0]['sparse_fea'], il.calls[DIM].slots_dev, 1, N_SPARSE * B)

  for it in range(3):
    run_fwd(it)
    run_bwd(it)
  torch.cuda.synchronize()
  fwd_ms = time_kernel(run_fwd, args.kernel_iters)
  bwd_ms = time_kernel(run_bwd, args.kernel_iters)

  # the two stages of K7 on their own: the dedup sort (L2-resident, latency-bound) and the HBM stage
  def run_sort(it):
    K.embedding_bwd_presort(rows_list[it % 4], arena.n_rows, DIM, call.ws, call.slots_dev, call.n_slots)

  def run_after_sort(it):
    K.embedding_bwd(arena.weight, arena.state0, None, DIM, rows_list[it % 4], call.slots_dev, call.n_slots,
                    call.n_seg, [gout], opt, call.ws, weights=w_list[it % 4], sorted_from=(call.ws, DIM))

  sort_ms = time_kernel(run_sort, args.kernel_iters)
  tot = 0.0
  for it in range(args.kernel_iters):   # sort untimed, then flush, then the timed HBM stage
    run_sort(it)
    flush.fill_(float(it))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    run_after_sort(it)
    e1.record(st)
    e1.synchronize()
    tot += e0.elapsed_time(e1)
  upd_ms = tot / args.kernel_iters
  U = float(np.mean(uniq))
  fwd_bytes, bwd_bytes = algorithmic_bytes(L, S, U, DIM, 4)
  fwd_bytes += 4 * L  # per-lookup weights (13 raw slots carry values)
  bwd_bytes += 4 * L
  # dense-tower GEMM on the tensor cores: the largest layer of the step (forward 624 -> 256), timed alone
  gx_ = torch.randn(B, F * DIM, device=dev)
  gw_ = torch.randn(F * DIM, 256, device=dev) * 0.05
  gout_ = torch.empty(B, 256, device=dev)
  for it in range(3):
    K.gemm(gx_, gw_, out=gout_)
  gemm_ms = time_kernel(lambda it: K.gemm(gx_, gw_, out=gout_), args.kernel_iters)
  gemm_flop = 2.0 * B * F * DIM * 256
  pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
      os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
  bf16_peak = float(pk.get('bf16_tflops', 2250.0))
  k_gemm = {'kernel': 'er_gemm (gemm_tf32x3_kernel, [8192 x 624] x [624 x 256])', 'bound': 'tensor',
            'achieved': gemm_flop / (gemm_ms * 1e-3) / 1e12, 'peak': bf16_peak, 'unit': 'TFLOP/s', 'ms': gemm_ms,
            'algorithmic_flop': gemm_flop,
            'note': 'fp32-accurate product = 3 TF32 MMAs per k-step at half the bf16 rate: tensor-pipe '
                    'work is 6x the algorithmic flop count against this bf16 peak'}
  k_gemm['frac'] = k_gemm['achieved'] / bf16_peak
  k_gemm['tensor_pipe_frac'] = 6.0 * k_gemm['frac']
  k_fwd = {'kernel': 'er_embedding_fwd (fwd_single_kernel<4,4>)', 'bound': 'hbm',
           'achieved': fwd_bytes / (fwd_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
           'ms': fwd_ms, 'algorithmic_bytes': fwd_bytes}
  k_bwd = {'kernel': 'er_embedding_bwd (init_hist + 3 x scatter radix sort + bwd_scan_vec_kernel<4> + bwd_long_vec_kernel<4,1>)',
           'bound': 'hbm', 'achieved': bwd_bytes / (bwd_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
           'ms': bwd_ms, 'algorithmic_bytes': bwd_bytes, 'unique_rows': U}
  k_upd = {'kernel': 'er_embedding_bwd after the sort (bwd_scan_vec_kernel<4> + bwd_long_vec_kernel<4,1>: segment '
                     'sums + fused adagrad row update)', 'bound': 'hbm',
           'achieved': (bwd_bytes - 8 * L) / (upd_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s', 'ms': upd_ms,
           'algorithmic_bytes': bwd_bytes - 8 * L, 'sort_ms': sort_ms,
           'note': 'the radix sort (sort_ms) depends only on the rows and runs on a side stream under the dense '
                   'forward/backward inside the step'}
  for k in (k_fwd, k_bwd, k_upd):
    k['frac'] = k['achieved'] / peak
  dom = k_bwd if bwd_ms >= fwd_ms else k_fwd
  roofline = {'bound': 'hbm', 'achieved': dom['achieved'], 'peak': peak, 'unit': 'GB/s', 'frac': dom['frac'],
              # ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the pipeline's largest launch
              # (bwd_scan_vec_kernel<4>: 37.03 MB read + 0.07 MB written during the launch; the updated rows are
              # written back from L2 after it) -- profiles/r01_ncu_bwd_scan_vec.txt; fwd_single_kernel: 23.4 MB
              'traffic': 37.1e6 if dom is k_bwd else 23.5e6, 'traffic_source': 'profiles/r01_ncu_*.txt',
              'kernel': dom['kernel'], 'peak_source': peak_src,
              'kernels': [k_fwd, k_bwd, k_upd, k_gemm],
              'random_64B_row_ceiling_gbs': 1000.0,
              'ceiling_note': 'tools/microbench_gather.cu: independent random 64 B row reads reach 15.6 Grows/s '
                              '(1.0 TB/s of rows) at 320K lookups on this B200, not the 6.57 TB/s copy peak'}

  # ---- CPU baseline (oracle port) on a bounded sample ---------------------------------
  cpu = None
  if world == 1 and not args.no_cpu_baseline:
    v, threads, dt = run_cpu(args, 3, 1, args.vocab)
    cpu = {'value': v, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
           'sample': '3 full training steps of batch %d on the CPU oracle (C sparse path + numpy dense, fastest '
                     'thread count of all/half/quarter/16/8), %.1f s' % (B, dt)}

  line = {'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
          'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
          'clocks': clocks,
          'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                  'ms_per_step': e2e_ms / e2e_steps},
          'gpu_launches': int(per_step_launches * args.steps), 'gpu_launches_per_step': int(per_step_launches),
          'cuda_graph': not args.no_graph, 'roofline': roofline, 'cpu_baseline': cpu,
          'final_loss': float(lv)}
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()
  return 0


if __name__ == '__main__':
  sys.exit(main())
