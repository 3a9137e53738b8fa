"""CPU: pin the oracle against every known answer available for the hot path."""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_kats.json')))


def test_fingerprint64_tensorflow_frozen_vectors():
  for s, want in KATS['fingerprint64']['vectors'].items():
    assert O.fingerprint64(s) == int(want), s
  assert O.fingerprint64('') == 0x9ae16a3b2f90404f  # k2 for the empty string


def test_string_to_hash_bucket_fast_examples():
  from easyrec_b200 import _lib
  assert len(KATS['fingerprint64']['hash_bucket_fast']) >= 6
  for case in KATS['fingerprint64']['hash_bucket_fast']:
    got = [O.fingerprint64(s) % case['num_buckets'] for s in case['inputs']]
    assert got == case['expected']
    assert [_lib.fingerprint64(s) % case['num_buckets'] for s in case['inputs']] == case['expected']   # product host hash


def test_integer_ids_hash_like_tensorflows_hashed_column():
  """HashedCategoricalColumn over int64 values 101, 201, 301 with 10 buckets -> [3, 7, 5] (TF's frozen test): the
  oracle's decimal-text rule (the restatement of K1's FARM_DECIMAL mode) must give the same buckets."""
  from easyrec_b200 import _lib
  rows, _ = O.bucketize(np.array([101, 201, 301], np.int64), _lib.BUCKET_FARM_DECIMAL, 10, 0)
  assert rows.tolist() == [3, 7, 5]


def test_product_host_hash_equals_oracle_all_lengths():
  # two independently written implementations (oracle C bytes-wise, product C++ memcpy-wise)
  from easyrec_b200 import _lib
  rnd = random.Random(7)
  for n in list(range(0, 200)) + [255, 256, 257, 1000, 4097]:
    for _ in range(3):
      s = bytes(rnd.getrandbits(8) for _ in range(n))
      assert _lib.fingerprint64(s) == O.fingerprint64(s), n


def test_embed_test_raw_known_answer():
  k = KATS['embed_test_raw']
  row_ptr, _ = O.csr_from_lens(k['lens'])
  out, _ = O.embedding_fwd(np.array(k['table'], np.float32), k['ids'], row_ptr, 0,
                           weights=np.array(k['weights'], np.float32))
  assert np.abs(out - np.array(k['expected'], np.float32)).max() < k['tolerance']


def test_embed_test_seq_multi_known_answer():
  k = KATS['embed_test_seq_multi']
  row_ptr, _ = O.csr_from_lens(k['lens'])
  out, _ = O.embedding_fwd(np.array(k['table'], np.float32), k['ids'], row_ptr, 1)
  for seg, want in list(k['expected_asserted'].items()) + list(k['expected_derived'].items()):
    assert np.abs(out[int(seg)] - np.array(want, np.float32)).max() < k['tolerance'], seg


def test_bucketize_rules():
  ids = np.array([0, 1, -1, -7, 12, 5, 2**62, -2**63], np.int64)
  # floored mod (python semantics), input/parquet_input.py:221
  rows, _ = O.bucketize(ids, 1, 5, 100)
  assert rows.tolist() == [100 + (int(v) % 5) for v in ids]
  # identity: -1 dropped, out of range -> 0
  rows, _ = O.bucketize(ids, 2, 10, 0)
  assert rows.tolist() == [0, 1, -1, 0, 0, 5, 0, 0]
  # hash of decimal text, negative numbers keep their sign
  rows, _ = O.bucketize(ids, 0, 1000003, 7)
  want = [7 + O.fingerprint64(str(int(v))) % 1000003 for v in ids]
  assert rows.tolist() == want
  # mod-sharding: owner = r % N, local = r // N (feature_column.py:296,317)
  rows, owner = O.bucketize(np.arange(20), 1, 1000, 0, shard_n=8)
  assert owner.tolist() == [i % 8 for i in range(20)]
  assert rows.tolist() == [i // 8 for i in range(20)]


def test_safe_lookup_pruning_and_combiners():
  table = np.arange(20, dtype=np.float32).reshape(10, 2) + 1
  ids = np.array([1, -1, 2, 3, 4, 5, 6], np.int64)
  w = np.array([1.0, 1.0, 0.0, 2.0, -1.0, 3.0, 4.0], np.float32)
  lens = [3, 0, 2, 2]
  row_ptr, _ = O.csr_from_lens(lens)
  s, _ = O.embedding_fwd(table, ids, row_ptr, 0, weights=w)
  np.testing.assert_allclose(s[0], 1 * table[1] + 0 * table[2])  # id<0 dropped, w=0 kept for sum
  np.testing.assert_allclose(s[1], 0)
  np.testing.assert_allclose(s[2], 2 * table[3] - table[4])
  m, sc = O.embedding_fwd(table, ids, row_ptr, 1, weights=w)
  np.testing.assert_allclose(m[0], table[1])  # w<=0 pruned for mean
  np.testing.assert_allclose(m[2], table[3])  # negative weight pruned
  np.testing.assert_allclose(m[3], (3 * table[5] + 4 * table[6]) / 7, rtol=1e-6)
  q, _ = O.embedding_fwd(table, ids, row_ptr, 2, weights=w)
  np.testing.assert_allclose(q[3], (3 * table[5] + 4 * table[6]) / 5, rtol=1e-6)


@pytest.mark.parametrize('kind', [O.OPT_SGD, O.OPT_ADAGRAD, O.OPT_LAZY_ADAM])
def test_optimizer_rules_against_formulas(kind):
  rng = np.random.default_rng(3)
  V, D, L = 50, 4, 200
  table = rng.normal(size=(V, D)).astype(np.float32)
  s0 = np.full((V, D), 0.1, np.float32) if kind == O.OPT_ADAGRAD else np.zeros((V, D), np.float32)
  s1 = np.zeros((V, D), np.float32)
  rows = rng.integers(0, V, L)
  rows[::17] = -1
  gseg = rng.normal(size=(L, D)).astype(np.float32)
  t0, a0, b0 = table.copy(), s0.copy(), s1.copy()
  n, ur, ug = O.embedding_bwd(table, s0, s1, rows, None, gseg, kind, 0.05, beta1_power=0.9**3,
                              beta2_power=0.999**3, want_uniq=True)
  G = np.zeros((V, D), np.float64)
  for l, r in enumerate(rows):
    if r >= 0:
      G[r] += gseg[l]
  touched = np.unique(rows[rows >= 0])
  assert n == touched.size and ur.tolist() == touched.tolist()
  np.testing.assert_allclose(ug, G[touched], rtol=1e-5, atol=1e-6)
  lr = 0.05
  if kind == O.OPT_SGD:
    want = t0 - lr * G
  elif kind == O.OPT_ADAGRAD:
    acc = a0 + G**2
    want = np.where(G != 0, t0 - lr * G / np.sqrt(acc), t0)
    np.testing.assert_allclose(s0[touched], acc[touched], rtol=1e-5)
  else:
    lr_t = lr * np.sqrt(1 - 0.999**3) / (1 - 0.9**3)
    # (1 - beta) is evaluated in fp32 by the TF graph (adam_s.py:196,202)
    m = float(np.float32(1) - np.float32(0.9)) * G
    v = float(np.float32(1) - np.float32(0.999)) * G**2
    want = t0 - lr_t * m / (np.sqrt(v) + 1e-8)
    np.testing.assert_allclose(s0[touched], m[touched], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(s1[touched], v[touched], rtol=1e-5, atol=1e-9)
  np.testing.assert_allclose(table[touched], want[touched], rtol=1e-5, atol=1e-6)
  untouched = np.setdiff1d(np.arange(V), touched)
  assert np.array_equal(table[untouched], t0[untouched])


def test_dense_oracle_matches_torch_autograd_cpu():
  """The numpy DNN/DeepFM backward is hand written: verify it with torch CPU autograd."""
  import torch
  rng = np.random.default_rng(0)
  B, F, D = 64, 5, 4

  def mk(i, o, bn=True):
    L = {'W': rng.normal(0, 0.3, (i, o)).astype(np.float32), 'b': rng.normal(0, 0.1, o).astype(np.float32)}
    if bn:
      L.update(gamma=rng.uniform(0.5, 1.5, o).astype(np.float32), beta=rng.normal(0, 0.1, o).astype(np.float32))
    return L

  params = {'dnn': [mk(F * D, 16), mk(16, 8)], 'final': [mk(1 + D + 8, 8), mk(8, 4)],
            'out_W': rng.normal(0, 0.3, (4, 1)).astype(np.float32), 'out_b': np.zeros(1, np.float32)}
  wide = rng.normal(size=(B, F)).astype(np.float32)
  deep = rng.normal(size=(B, F * D)).astype(np.float32)
  labels = (rng.uniform(size=B) < 0.3).astype(np.float32)
  logits, cache = O.deepfm_forward(wide, deep, F, D, params)
  loss, _, g_logits = O.sigmoid_ce(logits, labels)
  g_wide, g_deep, grads = O.deepfm_backward(g_logits, wide, deep, F, D, params, cache)

  tw = torch.tensor(wide, requires_grad=True)
  td = torch.tensor(deep, requires_grad=True)
  tp = {}

  def tdnn(x, layers, tag):
    for i, L in enumerate(layers):
      W = torch.tensor(L['W'], requires_grad=True)
      b = torch.tensor(L['b'], requires_grad=True)
      ga = torch.tensor(L['gamma'], requires_grad=True)
      be = torch.tensor(L['beta'], requires_grad=True)
      tp[(tag, i)] = (W, b, ga, be)
      z = x @ W + b
      mu = z.mean(0)
      var = ((z - mu)**2).mean(0)
      x = torch.relu((z - mu) / torch.sqrt(var + O.BN_EPS) * ga + be)
    return x

  v = td.reshape(B, F, D)
  fm = 0.5 * (v.sum(1)**2 - (v**2).sum(1))
  deep_fea = tdnn(td, params['dnn'], 'dnn')
  allf = torch.cat([tw.sum(1, keepdim=True), fm, deep_fea], 1)
  fin = tdnn(allf, params['final'], 'final')
  oW = torch.tensor(params['out_W'], requires_grad=True)
  tl = (fin @ oW)[:, 0]
  tloss = torch.nn.functional.binary_cross_entropy_with_logits(tl, torch.tensor(labels))
  tloss.backward()
  np.testing.assert_allclose(logits, tl.detach().numpy(), rtol=1e-4, atol=1e-5)
  assert abs(loss - float(tloss)) < 1e-5
  np.testing.assert_allclose(g_wide, tw.grad.numpy(), rtol=1e-3, atol=1e-6)
  np.testing.assert_allclose(g_deep, td.grad.numpy(), rtol=1e-3, atol=1e-6)
  for tag in ('dnn', 'final'):
    for i in range(2):
      W, b, ga, be = tp[(tag, i)]
      np.testing.assert_allclose(grads[tag][i]['W'], W.grad.numpy(), rtol=1e-3, atol=1e-6)
      np.testing.assert_allclose(grads[tag][i]['gamma'], ga.grad.numpy(), rtol=1e-3, atol=1e-6)
      np.testing.assert_allclose(grads[tag][i]['beta'], be.grad.numpy(), rtol=1e-3, atol=1e-6)
  np.testing.assert_allclose(grads['out_W'], oW.grad.numpy(), rtol=1e-3, atol=1e-6)


FORMULAS = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_formulas.json')))


def test_fm_oracle_matches_the_reference_code_output():
  """golden = layers/fm.py FM.__call__ executed on a numpy shim of its tf ops (make_formula_golden.py)."""
  c = FORMULAS['cases']['fm']
  x = np.asarray(c['x'], np.float32)          # [B, F, D]
  B, F, D = x.shape
  got = O.fm_fwd(x.reshape(B, F * D), F, D)
  np.testing.assert_allclose(got, np.asarray(c['y'], np.float32), rtol=1e-6, atol=1e-6)


def test_formula_golden_file_matches_its_generator_when_the_reference_is_mounted():
  if not os.path.isdir('/root/reference/easy_rec/python'):
    pytest.skip('reference checkout not mounted')
  import subprocess
  import sys
  import tempfile
  gen = os.path.join(os.path.dirname(__file__), 'golden', 'make_formula_golden.py')
  src = open(gen).read()
  with tempfile.TemporaryDirectory() as d:
    alt = os.path.join(d, 'gen.py')
    open(alt, 'w').write(src.replace("OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_formulas.json')",
                                     "OUT = %r" % os.path.join(d, 'out.json')))
    subprocess.check_call([sys.executable, alt], stdout=subprocess.DEVNULL)
    fresh = json.load(open(os.path.join(d, 'out.json')))
  assert fresh['cases'] == FORMULAS['cases']


def test_lazy_adam_oracle_matches_adam_s_sparse_apply():
  """golden = compat/adam_s.py _apply_sparse_shared executed for three steps (unique rows, summed grads)."""
  c = FORMULAS['cases']['lazy_adam_sparse']
  w = np.array(c['w0'], np.float32)
  m, v = np.zeros_like(w), np.zeros_like(w)
  p1, p2 = c['beta1'], c['beta2']
  for st in c['steps']:
    rows = np.array(st['indices'], np.int64)
    g = np.array(st['grad'], np.float32)
    O.embedding_bwd(w, m, v, rows, np.arange(rows.size, dtype=np.int32), g, O.OPT_LAZY_ADAM, c['lr'],
                    beta1=c['beta1'], beta2=c['beta2'], eps=c['epsilon'], beta1_power=p1, beta2_power=p2)
    p1, p2 = p1 * c['beta1'], p2 * c['beta2']
    np.testing.assert_allclose(m, np.array(st['m'], np.float32), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(v, np.array(st['v'], np.float32), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(w, np.array(st['w'], np.float32), rtol=2e-6, atol=1e-8)
  untouched = np.setdiff1d(np.arange(w.shape[0]), np.concatenate([s['indices'] for s in c['steps']]))
  assert untouched.size and np.array_equal(w[untouched], np.array(c['w0'], np.float32)[untouched])


def test_learning_rate_schedule_matches_exponential_decay_with_burnin():
  """golden = core/learning_schedules.py exponential_decay_with_burnin executed in fp32 (burn-in ramp, decay
  clock starting after the burn-in, staircase, floor)."""
  from easyrec_b200 import builder
  from easyrec_b200.config import config_util
  tmpl = ('train_config { optimizer_config { adagrad_optimizer { learning_rate { exponential_decay_learning_rate { '
          'initial_learning_rate: %r decay_steps: %d decay_factor: %r burnin_learning_rate: %r burnin_steps: %d '
          'min_learning_rate: %r staircase: %s } } } } }')
  for s in FORMULAS['cases']['lr_exponential_decay_with_burnin']['schedules']:
    cfg = config_util.get_configs_from_pipeline_file((tmpl % (
        s['initial_learning_rate'], s['decay_steps'], s['decay_factor'], s['burnin_learning_rate'], s['burnin_steps'],
        s['min_learning_rate'], 'true' if s['staircase'] else 'false')).encode())
    lr_fn = builder.optimizer_settings(cfg)['lr_fn']
    got = [lr_fn(st) for st in s['steps']]
    np.testing.assert_allclose(got, s['lr'], rtol=3e-6, atol=0)
  with pytest.raises(ValueError):
    builder.optimizer_settings(config_util.get_configs_from_pipeline_file(
        b'train_config { optimizer_config { adagrad_optimizer { } } }'))


LOOKUP = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_lookup.json')))['cases']


@pytest.mark.parametrize('k', [0, 1, 2])
def test_pooling_oracle_matches_embedding_lookup_ragged(k):
  """golden = compat/feature_column/feature_column.py embedding_lookup_ragged executed (unique -> gather ->
  sparse segment sum / mean / sqrtn; empty bags -> zeros)."""
  c = LOOKUP['embedding_lookup_ragged']
  case = c['outputs'][k]
  comb = {'sum': 0, 'mean': 1, 'sqrtn': 2}[case['combiner']]
  row_ptr, _ = O.csr_from_lens(np.array(c['lens'], np.int32))
  got, _ = O.embedding_fwd(np.array(c['table'], np.float32), np.array(c['ids'], np.int64), row_ptr, comb)
  np.testing.assert_allclose(got, np.array(case['y'], np.float32), rtol=1e-6, atol=1e-6)
  assert 0 in c['lens'] and not got[c['lens'].index(0)].any()


def test_shard_rule_oracle_matches_embedding_parallel_lookup_on_two_ranks():
  """golden = embedding_parallel_lookup executed by two threads whose hvd.alltoall really exchanges buffers
  (owner = id % N, local row = int64(id / N), sum combiner, [B, n_feat * D] output).  The oracle's bucketize
  with shard_n = N must address the same shard rows; gathering from the shards reproduces each rank's output."""
  from easyrec_b200 import _lib
  c = LOOKUP['embedding_parallel_lookup']
  N, B, F = c['world'], c['batch_size'], c['n_feature']
  full = np.array(c['table'], np.float32)
  V, D = full.shape
  assert c['shard_rows'] == (V + N - 1) // N
  shards = [full[r::N] for r in range(N)]
  for rank in c['ranks']:
    ids, lens = np.array(rank['ids'], np.int64), np.array(rank['lens'], np.int32)
    local, owner = O.bucketize(ids, _lib.BUCKET_IDENTITY, V, 0, shard_n=N)
    assert np.array_equal(owner, ids % N) and np.array_equal(local, (ids / N).astype(np.int64))
    rows = np.stack([shards[o][l] for o, l in zip(owner, local)]) if ids.size else np.zeros((0, D), np.float32)
    seg = np.repeat(np.arange(F * B), lens)
    pooled = np.zeros((F * B, D), np.float32)
    np.add.at(pooled, seg, rows)
    got = pooled.reshape(F, B, D).transpose(1, 0, 2).reshape(B, F * D)
    np.testing.assert_allclose(got, np.array(rank['y'], np.float32), rtol=1e-6, atol=1e-6)


def test_lookup_golden_file_matches_its_generator_when_the_reference_is_mounted(tmp_path):
  if not os.path.isdir('/root/reference/easy_rec/python'):
    pytest.skip('reference checkout not mounted')
  import importlib.util
  spec = importlib.util.spec_from_file_location(
      'make_lookup_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_lookup_golden.py'))
  gen = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(gen)
  gen.OUT = str(tmp_path / 'out.json')
  gen.main()
  assert json.load(open(gen.OUT))['cases'] == LOOKUP


def _layers(js):
  return [dict(W=np.array(l['w'], np.float32), b=np.array(l['b'], np.float32)) for l in js]


def test_deepfm_head_oracle_matches_build_predict_graph():
  """golden = model/deepfm.py DeepFM.build_predict_graph executed (reference FM class, dense stacks without BN):
  logits = dense(final_dnn(concat[sum(wide), fm, dnn(deep)]))."""
  c = FORMULAS['cases']['deepfm_head_final']
  params = dict(dnn=_layers(c['dnn']), final=_layers(c['final_dnn']),
                out_W=np.array(c['head']['output/kernel'], np.float32), out_b=np.array(c['head']['output/bias'], np.float32))
  logits, _ = O.deepfm_forward(np.array(c['wide'], np.float32), np.array(c['deep'], np.float32), c['n_field'], c['dim'], params)
  np.testing.assert_allclose(logits, np.array(c['logits'], np.float32)[:, 0], rtol=1e-5, atol=1e-5)
  # without final_dnn: wide + sum(fm) + dense(dnn(deep))  (model/deepfm.py:92-105)
  c = FORMULAS['cases']['deepfm_head_plain']
  wide, deep = np.array(c['wide'], np.float32), np.array(c['deep'], np.float32)
  h, _ = O.dnn_forward(deep, _layers(c['dnn']))
  want = wide.sum(1, keepdims=True) + O.fm_fwd(deep, c['n_field'], c['dim']).sum(1, keepdims=True) + \
      h @ np.array(c['head']['deep_logits/kernel'], np.float32) + np.array(c['head']['deep_logits/bias'], np.float32)
  np.testing.assert_allclose(want, np.array(c['logits'], np.float32), rtol=1e-5, atol=1e-5)


def test_adagrad_row_rule_reproduces_tensorflows_adagrad_test_constants():
  k = KATS['adagrad_sparse']
  for c in k['cases']:
    w = np.array(c['table'], np.float32)
    acc = np.full_like(w, k['initial_accumulator_value'])
    for _ in range(k['steps']):
      O.embedding_bwd(w, acc, None, np.array([c['row']], np.int64), np.zeros(1, np.int32),
                      np.array([[c['grad']]], np.float32), O.OPT_ADAGRAD, k['lr'])
    np.testing.assert_allclose(w, np.array(c['expected'], np.float32), rtol=k['tolerance'], atol=k['tolerance'])


@pytest.mark.parametrize('weighted', [True, False])
def test_pooling_oracle_follows_tensorflows_safe_lookup_test_cases(weighted):
  """ids < 0 pruned, weights <= 0 pruned for the mean combiner, empty rows -> zeros, weighted mean otherwise."""
  k = KATS['safe_embedding_lookup_sparse']
  e = np.random.default_rng(3).normal(size=(5, 4)).astype(np.float32)
  n_rows = k['dense_shape'][0]
  lens = np.bincount([i[0] for i in k['indices']], minlength=n_rows).astype(np.int32)
  row_ptr, _ = O.csr_from_lens(lens)
  got, _ = O.embedding_fwd(e, np.array(k['ids'], np.int64), row_ptr, 1,
                           weights=np.array(k['weights'], np.float32) if weighted else None)
  for r, spec in enumerate(k['expected_weighted' if weighted else 'expected_no_weights']):
    want = np.zeros(4, np.float32) if spec is None else sum(w * e[i] for i, w in spec['terms']) / spec['div']
    np.testing.assert_allclose(got[r], want, rtol=1e-6, atol=1e-6)


def test_interaction_oracles_match_the_reference_code_outputs():
  """DIN target attention, DCN cross v1, keras Cross (full / low rank), dot interaction, MMoE, list-wise match loss:
  the numpy restatements in oracle/oracle.py against the golden vectors produced by executing the reference."""
  c = FORMULAS['cases']
  f32 = np.float32

  def A(x):
    return np.array(x, f32)
  d = c['din_target_attention']
  att = O.din_attention(A(d['key']), A(d['hist']), np.array(d['lens']), _layers(d['mlp']))
  np.testing.assert_allclose(np.concatenate([att, A(d['key'])], axis=1), A(d['y']), rtol=1e-5, atol=1e-6)
  d = c['dcn_cross']
  np.testing.assert_allclose(O.cross_v1(A(d['x']), d['w'], d['b']), A(d['y']), rtol=1e-5, atol=1e-6)
  d = c['keras_cross_full']
  np.testing.assert_allclose(O.cross_v2(A(d['x0']), A(d['x']), d['w'], d['b'], d['diag_scale']), A(d['y']), rtol=1e-5, atol=1e-6)
  d = c['keras_cross_lowrank']
  np.testing.assert_allclose(O.cross_v2(A(d['x0']), A(d['x']), d['v'], d['b'], u=d['u']), A(d['y']), rtol=1e-5, atol=1e-6)
  for key in ('dot_interaction_self0', 'dot_interaction_self1'):
    d = c[key]
    np.testing.assert_allclose(O.dot_interaction(A(d['x']), d['self_interaction']), A(d['y']), rtol=1e-5, atol=1e-5)
  d = c['mmoe']
  got = O.mmoe(A(d['x']), [_layers(e) for e in d['experts']], [(g['w'], g['b']) for g in d['gates']])
  for a, b in zip(got, d['y']):
    np.testing.assert_allclose(a, A(b), rtol=1e-5, atol=1e-6)
  d = c['match_listwise']
  user, item = O.l2_normalize(A(d['user'])), O.l2_normalize(A(d['item']))
  np.testing.assert_allclose(user, A(d['user']), rtol=1e-6, atol=1e-6)       # already unit rows
  sim = (user @ item.T / f32(d['temperature'])).astype(f32)
  np.testing.assert_allclose(sim, A(d['sim']), rtol=1e-5, atol=1e-5)
  loss, probs = O.inbatch_softmax_ce(A(d['sim']), d['item_ids'], d['sample_weight'])
  np.testing.assert_allclose(probs, A(d['probs']), rtol=1e-5, atol=1e-7)
  assert abs(float(loss) - d['cross_entropy_loss']) < 1e-5
  reg_pos = np.mean(np.maximum(-(user * item).sum(1), 0) * A(d['sample_weight'])) / np.mean(A(d['sample_weight']))
  assert abs(float(reg_pos) - d['reg_pos_loss']) < 1e-6


def test_sigmoid_cross_entropy_oracle_reproduces_tensorflows_loss_test_values():
  for c in KATS['sigmoid_cross_entropy']['cases']:
    loss, _, _ = O.sigmoid_ce(np.array(c['logits'], np.float32), np.array(c['labels'], np.float32))
    assert round(abs(loss - c['expected']), c['places']) == 0
  # weights: the sum is divided by the number of NON-ZERO weights, not by their sum or by the batch size
  x, z, w = np.array([2.0, -1.0, 0.5, 3.0], np.float32), np.array([0, 1, 1, 0], np.float32), np.array([3.0, 0.0, 0.5, 2.0], np.float32)
  per = np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
  loss, _, _ = O.sigmoid_ce(x, z, w)
  assert abs(loss - float((per * w).sum() / 3)) < 1e-6


def test_bucketize_and_mean_pooling_reproduce_tensorflows_feature_column_tests():
  import types
  from easyrec_b200.input import readers
  k = KATS['bucketized_column']
  fc = types.SimpleNamespace(boundaries=k['boundaries'], num_buckets=0, min_val=0.0, max_val=0.0)
  got = readers.bucketize_raw(k['x'], fc)
  assert got.tolist() == k['expected_buckets']
  n = len(k['boundaries']) + 1                         # a 2-wide column offsets the k-th value by (len + 1) * k
  assert [int(b) + n * (i % 2) for i, b in enumerate(got)] == k['expected_ids_dim2']
  k = KATS['embedding_column_mean']
  row_ptr, _ = O.csr_from_lens(np.array(k['lens'], np.int32))
  out, _ = O.embedding_fwd(np.array(k['table'], np.float32), np.array(k['ids'], np.int64), row_ptr, 1)
  np.testing.assert_allclose(out, np.array(k['expected'], np.float32), rtol=1e-6, atol=1e-6)
