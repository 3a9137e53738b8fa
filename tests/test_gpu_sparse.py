"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Bit-exact for integer work (hash/bucketize, CSR, sort, dedup row sets); fp32 tolerances are
written next to each comparison.
"""
import json
import os

import numpy as np
import pytest
import torch

from easyrec_b200 import _lib, kernels as K
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
KATS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_kats.json')))


def t(a, dtype=None):
  x = torch.from_numpy(np.ascontiguousarray(a))
  if dtype is not None:
    x = x.to(dtype)
  return x.to(DEV)


def edge_ids(rng, n):
  edges = [0, 1, -1, 9, 10, 11, 99, 100, 101, -9, -10, -99999, 2**31 - 1, 2**31, -2**31, 2**63 - 1,
           -2**63, -2**63 + 1, 10**18, 10**18 - 1, 10**18 + 1, -10**18, 12345678, 123456789012,
           99999999, 100000000, 9999999999999999, 10000000000000000, 2**53, 2**53 + 1]
  edges += [10**k for k in range(19)] + [10**k - 1 for k in range(1, 19)] + [-(10**k) for k in range(19)]
  r = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)
  small = rng.integers(-1000, 100000, n, dtype=np.int64)
  lens = rng.integers(1, 19, n)
  varlen = rng.integers(0, 2**62, n, dtype=np.int64) % (1 + 10**np.minimum(lens, 18))
  return np.concatenate([np.array(edges, np.int64), r, small, varlen])


def test_bucketize_bit_exact_all_modes():
  rng = np.random.default_rng(11)
  ids = edge_ids(rng, 20000)
  n = ids.size
  # four slots, one per mode, single-valued segments laid out back to back
  q = n // 4
  bounds = [0, q, 2 * q, 3 * q, n]
  modes = [_lib.BUCKET_FARM_DECIMAL, _lib.BUCKET_MOD, _lib.BUCKET_IDENTITY, _lib.BUCKET_NONE]
  nbs = [1000003, 977, 5000, 1 << 62]
  offs = [0, 1000003, 1000003 + 977, 1000003 + 977 + 5000]
  recs = [dict(num_buckets=nbs[i], row_offset=offs[i], seg_begin=bounds[i], n_seg=bounds[i + 1] - bounds[i],
               bucket_mode=modes[i]) for i in range(4)]
  slots = K.make_slots(recs)
  rows = K.bucketize(t(ids), K.slots_to_device(slots, DEV), 4, n).cpu().numpy()
  mode_l = np.concatenate([np.full(bounds[i + 1] - bounds[i], modes[i]) for i in range(4)])
  nb_l = np.concatenate([np.full(bounds[i + 1] - bounds[i], nbs[i]) for i in range(4)])
  off_l = np.concatenate([np.full(bounds[i + 1] - bounds[i], offs[i]) for i in range(4)])
  want, _ = O.bucketize(ids, mode_l, nb_l, off_l)
  assert np.array_equal(rows, want)


def test_bucketize_hash_every_decimal_length_and_sharding():
  rng = np.random.default_rng(5)
  ids = []
  for nd in range(1, 20):
    lo, hi = 10**(nd - 1), min(10**nd - 1, 2**63 - 1)
    v = rng.integers(lo, hi, 300, dtype=np.int64)
    ids += [v, -v]
  ids = np.concatenate(ids + [np.array([2**63 - 1, -2**63], np.int64)])
  n = ids.size
  for nb, shard in [(10_000_000, 1), (2**63 - 1, 1), (200_000_000, 8), (7, 2)]:
    slots = K.make_slots([dict(num_buckets=nb, row_offset=3, seg_begin=0, n_seg=n,
                               bucket_mode=_lib.BUCKET_FARM_DECIMAL, shard_n=shard)])
    owner = torch.empty(n, dtype=torch.int32, device=DEV)
    rows = K.bucketize(t(ids), K.slots_to_device(slots, DEV), 1, n, owner=owner).cpu().numpy()
    want, wown = O.bucketize(ids, 0, nb, 3, shard_n=shard)
    assert np.array_equal(rows, want), (nb, shard)
    assert np.array_equal(owner.cpu().numpy(), wown)
  # spot-check against the pure-python statement of the rule
  for v in (0, 7, -1, 1234567890123, -2**63):
    slots = K.make_slots([dict(num_buckets=1000, row_offset=0, seg_begin=0, n_seg=1, bucket_mode=0)])
    got = int(K.bucketize(t(np.array([v], np.int64)), K.slots_to_device(slots, DEV), 1, 1).item())
    assert got == O.fingerprint64(str(v)) % 1000


def test_csr_from_lens_matches_cumsum():
  rng = np.random.default_rng(2)
  for n_seg in (1, 5, 2047, 2048, 2049, 100_000, 655_360):
    lens = rng.integers(0, 6, n_seg).astype(np.int32)
    lens[rng.integers(0, n_seg, max(n_seg // 10, 1))] = 0
    total = int(lens.sum())
    row_ptr, seg_ids = K.csr_from_lens(t(lens), total)
    wp, ws = O.csr_from_lens(lens)
    assert np.array_equal(row_ptr.cpu().numpy(), wp)
    assert np.array_equal(seg_ids.cpu().numpy()[:total], ws)


def _pooled_from_bufs(bufs, slots, dim):
  out = []
  for s in slots:
    b = bufs[s['out_buf']]
    out.append(b[:s['n_seg'], s['out_col']:s['out_col'] + dim])
  return np.concatenate(out, 0)


@pytest.mark.parametrize('dim', [1, 3, 4, 8, 16, 32, 64, 128])
def test_embedding_fwd_single_valued_parity(dim):
  rng = np.random.default_rng(dim)
  V, B, F = 5000, 300, 7
  table = rng.normal(size=(V, dim)).astype(np.float32)
  rows = rng.integers(0, V, B * F).astype(np.int64)
  rows[rng.integers(0, B * F, 50)] = -1
  stride = ((F * dim + 3) // 4) * 4
  recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=3, combiner=0,
               out_buf=0, out_stride=stride, out_col=f * dim) for f in range(F)]
  slots = K.make_slots(recs)
  out = torch.full((B, stride), 7.0, device=DEV)
  K.embedding_fwd(t(table), dim, t(rows), K.slots_to_device(slots, DEV), F, B * F, [out])
  want, _ = O.embedding_fwd(table, rows, np.arange(B * F + 1, dtype=np.int32), 0)
  got = _pooled_from_bufs([out.cpu().numpy()], slots, dim)
  assert np.array_equal(got, want)  # single id, no arithmetic: bit exact


@pytest.mark.parametrize('dim', [1, 4, 16, 32, 6])
def test_embedding_fwd_csr_weighted_combiners_parity(dim):
  rng = np.random.default_rng(100 + dim)
  V, B = 3000, 257
  combs = [0, 1, 2, 0, 1]
  F = len(combs)
  table = rng.normal(size=(V, dim)).astype(np.float32)
  lens = rng.integers(0, 9, B * F).astype(np.int32)
  lens[::13] = 0
  lens[5] = 70  # a long segment
  L = int(lens.sum())
  rows = rng.integers(0, V, L).astype(np.int64)
  rows[rng.integers(0, L, L // 20)] = -1
  w = rng.normal(1.0, 1.0, L).astype(np.float32)  # includes <= 0 weights
  row_ptr, seg_ids = K.csr_from_lens(t(lens), L)
  recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=3, combiner=combs[f],
               out_buf=f % 2, out_stride=F * dim + (4 - F * dim % 4) % 4, out_col=f * dim) for f in range(F)]
  slots = K.make_slots(recs)
  stride = recs[0]['out_stride']
  bufs = [torch.zeros(B, stride, device=DEV), torch.zeros(B, stride, device=DEV)]
  scale = torch.empty(B * F, device=DEV)
  K.embedding_fwd(t(table), dim, t(rows), K.slots_to_device(slots, DEV), F, B * F, bufs, weights=t(w),
                  row_ptr=row_ptr, seg_scale=scale)
  comb_seg = np.repeat(np.array(combs, np.int32), B)
  want, wscale = O.embedding_fwd(table, rows, O.csr_from_lens(lens)[0], comb_seg, weights=w)
  got = _pooled_from_bufs([b.cpu().numpy() for b in bufs], slots, dim)
  # same operation order as the oracle (sequential, separate mul/add): bit exact
  assert np.array_equal(got, want)
  assert np.array_equal(scale.cpu().numpy(), wscale)


def test_embed_test_known_answers_on_gpu():
  for name, comb in (('embed_test_raw', 0), ('embed_test_seq_multi', 1)):
    k = KATS[name]
    table = np.array(k['table'], np.float32)
    lens = np.array(k['lens'], np.int32)
    ids = np.array(k['ids'], np.int64)
    n_seg = lens.size
    row_ptr, _ = K.csr_from_lens(t(lens), ids.size)
    slots = K.make_slots([dict(num_buckets=5, row_offset=0, seg_begin=0, n_seg=n_seg,
                               bucket_mode=_lib.BUCKET_IDENTITY, combiner=comb, out_buf=0, out_stride=2,
                               out_col=0)])
    sd = K.slots_to_device(slots, DEV)
    rows = K.bucketize(t(ids), sd, 1, n_seg, seg_ids=K.csr_from_lens(t(lens), ids.size)[1], row_ptr=row_ptr)
    out = torch.zeros(n_seg, 2, device=DEV)
    w = t(np.array(k['weights'], np.float32)) if 'weights' in k else None
    K.embedding_fwd(t(table), 2, rows, sd, 1, n_seg, [out], weights=w, row_ptr=row_ptr)
    got = out.cpu().numpy()
    if name == 'embed_test_raw':
      assert np.abs(got - np.array(k['expected'], np.float32)).max() < k['tolerance']
    else:
      for seg, want in list(k['expected_asserted'].items()) + list(k['expected_derived'].items()):
        assert np.abs(got[int(seg)] - np.array(want, np.float32)).max() < k['tolerance']


def test_sort_rows_is_stable_and_complete():
  rng = np.random.default_rng(9)
  for n, max_row in [(1, 10), (1023, 100), (1024, 70000), (1025, 2**24 + 5), (300_000, 10_000_013),
                     (212_992, 2**31 + 7)]:
    rows = rng.integers(0, max_row, n).astype(np.int64)
    rows[rng.integers(0, n, n // 50 + 1)] = -1
    if n > 5000:
      rows[rng.integers(0, n, n // 3)] = 17  # a hot row
    keys, vals = K.sort_rows(t(rows), max_row)
    keys = keys.cpu().numpy().view(np.uint32).astype(np.int64)
    vals = vals.cpu().numpy().view(np.uint32).astype(np.int64)
    k_in = np.where(rows < 0, max_row, rows)
    order = np.argsort(k_in, kind='stable')
    assert np.array_equal(vals, order)
    assert np.array_equal(keys, k_in[order])


def _run_bwd(kind, dim, rng, V, B, F, with_csr, hot):
  combs = [0, 1, 2][:F] if with_csr else [0] * F
  table = rng.normal(size=(V, dim)).astype(np.float32)
  s0 = np.full((V, dim), 0.1, np.float32) if kind == _lib.OPT_ADAGRAD else np.zeros((V, dim), np.float32)
  s1 = np.zeros((V, dim), np.float32)
  if with_csr:
    lens = rng.integers(0, 5, B * F).astype(np.int32)
    L = int(lens.sum())
  else:
    lens = np.ones(B * F, np.int32)
    L = B * F
  rows = rng.integers(0, V, L).astype(np.int64)
  if hot:
    rows[rng.integers(0, L, L // 3)] = 5  # > kLongRun duplicates -> CTA-wide path
    rows[rng.integers(0, L, 70)] = 9
  rows[rng.integers(0, L, L // 25 + 1)] = -1
  w = rng.uniform(0.1, 2.0, L).astype(np.float32) if with_csr else None
  stride = F * dim + (4 - F * dim % 4) % 4
  gout = rng.normal(size=(B, stride)).astype(np.float32)
  recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=3, combiner=combs[f],
               out_buf=0, out_stride=stride, out_col=f * dim) for f in range(F)]
  slots = K.make_slots(recs)
  sd = K.slots_to_device(slots, DEV)
  d_table, d_s0, d_s1 = t(table), t(s0), t(s1)
  d_rows = t(rows)
  row_ptr = seg_ids = scale = None
  if with_csr:
    row_ptr, seg_ids = K.csr_from_lens(t(lens), L)
    scale = torch.empty(B * F, device=DEV)
    out = torch.empty(B, stride, device=DEV)
    K.embedding_fwd(d_table, dim, d_rows, sd, F, B * F, [out], weights=t(w), row_ptr=row_ptr, seg_scale=scale)
  opt = K.make_opt(kind, 0.05, beta1_power=0.9**4, beta2_power=0.999**4, grad_scale=0.5)
  ws = K.bwd_workspace(L, DEV, dim)
  ur = torch.empty(L, dtype=torch.int64, device=DEV)
  ug = torch.empty(L, dim, device=DEV)
  nu = torch.zeros(1, dtype=torch.int32, device=DEV)
  K.embedding_bwd(d_table, d_s0 if kind != _lib.OPT_SGD else None,
                  d_s1 if kind in (_lib.OPT_LAZY_ADAM,) else None, dim, d_rows, sd, F, B * F, [t(gout)], opt,
                  ws, weights=None if w is None else t(w), seg_ids=seg_ids, row_ptr=row_ptr, seg_scale=scale,
                  uniq_rows=ur, uniq_grads=ug, n_uniq=nu)
  torch.cuda.synchronize()
  # oracle
  gseg = np.concatenate([gout[:, f * dim:(f + 1) * dim] for f in range(F)], 0)
  _, seg_of = O.csr_from_lens(lens)
  comb_seg = np.repeat(np.array(combs, np.int32), B)
  _, oscale = O.embedding_fwd(table, rows, O.csr_from_lens(lens)[0], comb_seg, weights=w)
  okind = {_lib.OPT_SGD: O.OPT_SGD, _lib.OPT_ADAGRAD: O.OPT_ADAGRAD, _lib.OPT_LAZY_ADAM: O.OPT_LAZY_ADAM}[kind]
  n, our, oug = O.embedding_bwd(table, s0, s1, rows, seg_of, gseg, okind, 0.05, weights=w,
                                seg_scale=oscale if with_csr else None, beta1_power=0.9**4,
                                beta2_power=0.999**4, grad_scale=0.5, want_uniq=True)
  return dict(n=n, our=our, oug=oug, table=table, s0=s0, s1=s1, nu=int(nu.item()), ur=ur.cpu().numpy(),
              ug=ug.cpu().numpy(), d_table=d_table.cpu().numpy(), d_s0=d_s0.cpu().numpy(), d_s1=d_s1.cpu().numpy())


@pytest.mark.parametrize('kind', [_lib.OPT_SGD, _lib.OPT_ADAGRAD, _lib.OPT_LAZY_ADAM])
@pytest.mark.parametrize('dim,with_csr,hot', [(16, False, False), (16, True, True), (1, False, True), (32, True, False),
                                              (6, True, True), (4, False, True)])
def test_embedding_bwd_dedup_update_parity(kind, dim, with_csr, hot):
  rng = np.random.default_rng(kind * 100 + dim)
  r = _run_bwd(kind, dim, rng, V=4000, B=500, F=3, with_csr=with_csr, hot=hot)
  # dedup: the set of touched rows and its order are integer work -> bit exact
  assert r['nu'] == r['n']
  assert np.array_equal(r['ur'][:r['n']], r['our'])
  # summed gradients: runs <= 64 are summed in the oracle's order (bit exact); hot rows use a
  # fixed tree -> fp32 reassociation, tolerance 1e-5 relative to the row's gradient scale
  np.testing.assert_allclose(r['ug'][:r['n']], r['oug'], rtol=2e-5, atol=2e-5)
  short = np.ones(r['n'], bool)
  short[np.isin(r['our'], [5, 9])] = False
  # (dim <= 32 sums a run with a fixed shuffle-scan tree: last-ulp differences vs sequential order)
  np.testing.assert_allclose(r['ug'][:r['n']][short], r['oug'][short], rtol=2e-6, atol=2e-6)
  # post-step rows and optimizer state: <= 1e-6 abs (BASELINE.md parity gate) for every row whose
  # gradient was summed in the oracle's order; the two hot rows (fixed-tree sum of ~500 N(0,1)
  # gradients, |G| ~ 30) get the same bound relative to their magnitude
  hot_rows = np.array([5, 9])
  cold = np.ones(r['table'].shape[0], bool)
  cold[hot_rows] = False
  np.testing.assert_allclose(r['d_table'][cold], r['table'][cold], rtol=0, atol=1e-6)
  np.testing.assert_allclose(r['d_s0'][cold], r['s0'][cold], rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(r['d_s1'][cold], r['s1'][cold], rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(r['d_table'][hot_rows], r['table'][hot_rows], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(r['d_s0'][hot_rows], r['s0'][hot_rows], rtol=1e-4, atol=1e-5)
  np.testing.assert_allclose(r['d_s1'][hot_rows], r['s1'][hot_rows], rtol=1e-4, atol=1e-5)


def test_bwd_ten_steps_adagrad_tracks_oracle():
  rng = np.random.default_rng(77)
  V, B, F, dim = 2000, 256, 4, 16
  table = rng.normal(0, 0.01, (V, dim)).astype(np.float32)
  acc = np.full((V, dim), 0.1, np.float32)
  d_table, d_acc = t(table), t(acc)
  stride = F * dim
  recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=3, combiner=0, out_buf=0,
               out_stride=stride, out_col=f * dim) for f in range(F)]
  sd = K.slots_to_device(K.make_slots(recs), DEV)
  ws = K.bwd_workspace(B * F, DEV, dim)
  for step in range(10):
    rows = (rng.zipf(1.3, B * F) % V).astype(np.int64)
    gout = rng.normal(0, 0.1, (B, stride)).astype(np.float32)
    K.embedding_bwd(d_table, d_acc, None, dim, t(rows), sd, F, B * F, [t(gout)],
                    K.make_opt(_lib.OPT_ADAGRAD, 0.01), ws)
    gseg = np.concatenate([gout[:, f * dim:(f + 1) * dim] for f in range(F)], 0)
    O.embedding_bwd(table, acc, None, rows, None, gseg, O.OPT_ADAGRAD, 0.01)
  np.testing.assert_allclose(d_table.cpu().numpy(), table, rtol=0, atol=1e-6)
  np.testing.assert_allclose(d_acc.cpu().numpy(), acc, rtol=1e-6, atol=1e-6)


def test_fm_and_sigmoid_ce_parity():
  rng = np.random.default_rng(4)
  for B, F, D in [(100, 39, 16), (33, 5, 6), (8192, 39, 16)]:
    x = rng.normal(size=(B, F * D)).astype(np.float32)
    gy = rng.normal(size=(B, D)).astype(np.float32)
    y = K.fm_fwd(t(x), F, D).cpu().numpy()
    want = O.fm_fwd(x, F, D)
    assert np.array_equal(y, want)  # same sequential order -> bit exact
    gx = K.fm_bwd(t(x), t(gy), F, D).cpu().numpy()
    np.testing.assert_allclose(gx, O.fm_bwd(x, gy, F, D), rtol=1e-6, atol=1e-6)
  logits = rng.normal(0, 3, 8192).astype(np.float32)
  labels = (rng.uniform(size=8192) < 0.25).astype(np.float32)
  loss, probs, g = K.sigmoid_ce(t(logits), t(labels))
  wl, wp, wg = O.sigmoid_ce(logits, labels)
  assert abs(float(loss.item()) - wl) < 1e-5
  np.testing.assert_allclose(probs.cpu().numpy(), wp, atol=1e-6)
  np.testing.assert_allclose(g.cpu().numpy(), wg, atol=1e-8, rtol=1e-5)
  # sample weights (data_config.sample_weight -> tf.losses.sigmoid_cross_entropy(weights), SUM_BY_NONZERO_WEIGHTS):
  # RankModel.weighted_ce hands the kernel w * B / count_nonzero(w); zero weights are real zeros in the gradient
  from easyrec_b200.model.rank_model import RankModel
  w = rng.uniform(0, 3, 8192).astype(np.float32)
  w[rng.uniform(size=8192) < 0.2] = 0.0
  tl = t(logits).requires_grad_(True)
  loss, probs = RankModel.weighted_ce(tl, t(labels), t(w))
  loss.backward()
  wl, wp, wg = O.sigmoid_ce(logits, labels, weights=w)
  assert abs(float(loss) - wl) < 2e-5 * max(1.0, abs(wl))
  np.testing.assert_allclose(probs.cpu().numpy(), wp, atol=1e-6)
  np.testing.assert_allclose(tl.grad.cpu().numpy(), wg, atol=1e-8, rtol=2e-5)
  assert np.all(tl.grad.cpu().numpy()[w == 0] == 0)


def test_fm_block_one_pass_matches_oracle():
  """fused FM + sum-of-squares forward, merged (tower + FM + regulariser) gradient backward."""
  rng = np.random.default_rng(5)
  for B, F, D in [(100, 39, 16), (37, 5, 8), (8192, 39, 16), (64, 26, 32), (19, 3, 4)]:
    x = rng.normal(size=(B, F * D)).astype(np.float32)
    gy = rng.normal(size=(B, D)).astype(np.float32)
    gp = rng.normal(size=(B, F * D)).astype(np.float32)
    coef = np.float32(0.37)
    y, sumsq = K.fm_block_fwd(t(x), F, D)
    # shuffle-tree vs sequential summation order: bound each element by the size of the two terms whose
    # difference it is (0.5*(S^2 - Q) cancels), not by the result
    x3 = x.astype(np.float64).reshape(B, F, D)
    s2, qq = x3.sum(1) ** 2, (x3 ** 2).sum(1)
    assert (np.abs(y.cpu().numpy() - 0.5 * (s2 - qq)) <= 2e-6 * (s2 + qq) + 1e-7).all()
    assert np.abs(y.cpu().numpy() - O.fm_fwd(x, F, D)).max() <= 1e-5 * max(1.0, float(np.abs(s2 + qq).max()))
    ref_sq = float((x.astype(np.float64) ** 2).sum())
    assert abs(float(sumsq.item()) - ref_sq) <= 2e-6 * ref_sq
    y2, sumsq2 = K.fm_block_fwd(t(x), F, D)
    assert torch.equal(y, y2) and torch.equal(sumsq, sumsq2)   # deterministic, workspace self-resets
    gx = K.fm_block_bwd(t(x), t(gy), t(gp), t(np.array([coef])), 2.0, F, D).cpu().numpy()
    want_gx = O.fm_bwd(x, gy, F, D).astype(np.float64) + gp + 2.0 * float(coef) * x.astype(np.float64)
    np.testing.assert_allclose(gx, want_gx, rtol=1e-5, atol=1e-5)
    gx0 = K.fm_block_bwd(t(x), t(gy), None, None, 0.0, F, D).cpu().numpy()
    np.testing.assert_allclose(gx0, O.fm_bwd(x, gy, F, D), rtol=1e-5, atol=1e-5)


def test_bwd_reuse_sort_is_identical_to_a_fresh_sort():
  """A second table looked up with the same rows (DeepFM's wide dim-1 table) reuses the first K7's sort."""
  rng = np.random.default_rng(21)
  V, B, F = 3000, 400, 5
  rows = (rng.zipf(1.2, B * F) % V).astype(np.int64)
  rows[rng.integers(0, B * F, 30)] = -1
  rows[rng.integers(0, B * F, 200)] = 7   # hot row -> chunked path
  d_rows = t(rows)
  res = {}
  for mode in ('fresh', 'reuse'):
    out = []
    ws16 = K.bwd_workspace(B * F, DEV, 16)
    src = None
    for dim in (16, 1):
      r2 = np.random.default_rng(dim)
      table = t(r2.normal(size=(V, dim)).astype(np.float32))
      acc = t(np.full((V, dim), 0.1, np.float32))
      stride = (F * dim + 3) // 4 * 4
      gout = t(r2.normal(size=(B, stride)).astype(np.float32))
      recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=3, combiner=0, out_buf=0,
                   out_stride=stride, out_col=f * dim) for f in range(F)]
      sd = K.slots_to_device(K.make_slots(recs), DEV)
      ws = ws16 if dim == 16 else K.bwd_workspace(B * F, DEV, dim)
      K.embedding_bwd(table, acc, None, dim, d_rows, sd, F, B * F, [gout], K.make_opt(_lib.OPT_ADAGRAD, 0.05), ws,
                      sorted_from=src if (mode == 'reuse' and dim == 1) else None)
      src = (ws16, 16)
      out.append((table.cpu(), acc.cpu()))
    res[mode] = out
  for (ta, aa), (tb, ab) in zip(res['fresh'], res['reuse']):
    assert torch.equal(ta, tb) and torch.equal(aa, ab)


def test_full_size_c2_properties():
  """BASELINE config 2 sizes (B=8192, 26+13 slots, V=10M, D=16): size-independent properties."""
  B, F, D, V = 8192, 39, 16, 10_000_013
  g = torch.Generator(device=DEV).manual_seed(1)
  table = torch.randn(V, D, device=DEV, generator=g) * 0.01
  ids = torch.randint(0, 2**40, (B * F,), device=DEV, generator=g, dtype=torch.int64)
  recs = [dict(num_buckets=V, row_offset=0, seg_begin=f * B, n_seg=B, bucket_mode=_lib.BUCKET_FARM_DECIMAL,
               combiner=0, out_buf=0, out_stride=F * D, out_col=f * D) for f in range(F)]
  sd = K.slots_to_device(K.make_slots(recs), DEV)
  rows = K.bucketize(ids, sd, F, B * F)
  assert int(rows.min()) >= 0 and int(rows.max()) < V
  # oracle on a sample of the full-size batch (bit exact)
  idx = torch.randint(0, B * F, (4096,), device=DEV, generator=g)
  want, _ = O.bucketize(ids[idx].cpu().numpy(), 0, V, 0)
  assert np.array_equal(rows[idx].cpu().numpy(), want)
  out = torch.empty(B, F * D, device=DEV)
  K.embedding_fwd(table, D, rows, sd, F, B * F, [out])
  # gather == index_select, laid out [B, F*D]
  ref = table[rows].reshape(F, B, D).permute(1, 0, 2).reshape(B, F * D)
  assert torch.equal(out, ref)
  # linearity of the pooled output in the table
  out2 = torch.empty_like(out)
  K.embedding_fwd(table * 2, D, rows, sd, F, B * F, [out2])
  assert torch.equal(out2, out * 2)
  # sort: sorted, a permutation, stable
  keys, vals = K.sort_rows(rows, V)
  k64 = keys.to(torch.int64) & 0xffffffff
  v64 = vals.to(torch.int64) & 0xffffffff
  assert bool((k64[1:] >= k64[:-1]).all())
  assert torch.equal(torch.sort(v64).values, torch.arange(B * F, device=DEV))
  same = k64[1:] == k64[:-1]
  assert bool((v64[1:][same] > v64[:-1][same]).all())
  assert torch.equal(rows[v64], k64)
  # dedup checksum: sum of per-row summed grads == column sums of the upstream gradient
  gout = torch.randn(B, F * D, device=DEV, generator=g)
  ur = torch.empty(B * F, dtype=torch.int64, device=DEV)
  ug = torch.empty(B * F, D, device=DEV)
  nu = torch.zeros(1, dtype=torch.int32, device=DEV)
  ws = K.bwd_workspace(B * F, DEV, D)
  acc = torch.full((V, D), 0.1, device=DEV)
  t0 = table.clone()
  K.embedding_bwd(table, acc, None, D, rows, sd, F, B * F, [gout], K.make_opt(_lib.OPT_ADAGRAD, 0.01), ws,
                  uniq_rows=ur, uniq_grads=ug, n_uniq=nu)
  n = int(nu.item())
  assert n == int(torch.unique(rows).numel())
  assert torch.equal(ur[:n], torch.unique(rows))
  tot = ug[:n].double().sum(0)
  ref_tot = gout.double().reshape(B, F, D).sum((0, 1))
  assert float((tot - ref_tot).abs().max()) < 1e-3
  # update touched exactly the deduplicated rows, and matches the closed form
  changed = (table != t0).any(1).nonzero().flatten()
  assert set(changed.tolist()) <= set(ur[:n].tolist())
  G = ug[:n]
  want_acc = 0.1 + G * G
  want_w = t0[ur[:n]] - 0.01 * G / want_acc.sqrt()
  assert float((acc[ur[:n]] - want_acc).abs().max()) < 1e-6
  assert float((table[ur[:n]] - want_w).abs().max()) < 1e-6
  # idempotence of the lookup after zero-gradient step
  t1 = table.clone()
  K.embedding_bwd(table, acc, None, D, rows, sd, F, B * F, [torch.zeros_like(gout)],
                  K.make_opt(_lib.OPT_ADAGRAD, 0.01), ws)
  assert torch.equal(table, t1)
