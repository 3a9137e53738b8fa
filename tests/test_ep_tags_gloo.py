"""CPU, world 2 and 3 over gloo, kernel doubles: EmbeddingParallel with MULTI-VALUED slots - TagFeatures (kv weights, mean /
sum combiners) and a multi-valued SequenceFeature next to single-valued ids - the ragged forms of
embedding_parallel_lookup (compat/feature_column/feature_column.py:248-357 `ragged_ids / ragged_lens / ragged_wgts`).
Each launch kind has its own exchange: K1 (owner, local row) over the CSR -> K8 -> all-to-alls -> the received rows
pooled by the same CSR; backward: local duplicate sums -> owners -> fused update.  The row-sharded model must stay the
same model as replicated data parallel (which gathers the lookups' segments, tests/test_dp_tags_gloo.py) on the same
per-rank batches."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))

CFG = b'''
train_config { %s
  optimizer_config { adagrad_optimizer { learning_rate { constant_learning_rate { learning_rate: 0.05 } } } } }
data_config { batch_size: 12 input_type: CSVInput separator: "," label_fields: "label"
  input_fields { input_name: "label" input_type: FLOAT } input_fields { input_name: "u" input_type: INT64 }
  input_fields { input_name: "t1" input_type: STRING } input_fields { input_name: "t2" input_type: STRING }
  input_fields { input_name: "x" input_type: FLOAT } input_fields { input_name: "key" input_type: INT64 }
  input_fields { input_name: "clk" input_type: STRING } }
feature_config {
  features { input_names: "u" feature_type: IdFeature embedding_dim: 4 hash_bucket_size: 23 }
  features { input_names: "t1" feature_type: TagFeature embedding_dim: 4 num_buckets: 19 separator: "|" kv_separator: ":"
             combiner: "mean" }
  features { input_names: "t2" feature_type: TagFeature embedding_dim: 4 hash_bucket_size: 13 separator: "|" combiner: "sum" }
  features { input_names: "x" feature_type: RawFeature embedding_dim: 4 min_val: 0.0 max_val: 2.0 }
  features { input_names: "key" feature_type: IdFeature embedding_dim: 4 num_buckets: 11 }
  features { input_names: "clk" feature_type: SequenceFeature embedding_dim: 4 num_buckets: 11 separator: "|" seq_multi_sep: "#"
             combiner: "mean" max_seq_len: 3 } }
model_config { model_class: "MultiTowerDIN"
  seq_att_groups { group_name: "din" seq_att_map { key: "key" hist_seq: "clk" } }
  feature_groups { group_name: "g" feature_names: ["u", "t1", "t2", "x"] wide_deep: DEEP }
  multi_tower { towers { input: "g" dnn { hidden_units: [8] } } din_towers { input: "din" dnn { hidden_units: [4, 1] } }
                final_dnn { hidden_units: [4] } l2_regularization: 1e-5 }
  embedding_regularization: 1e-5 }
'''


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _lines(B, seed):
  rng = np.random.default_rng(seed)
  out = []
  for _ in range(B):
    t1 = '|'.join('%d:%.2f' % (rng.integers(0, 19), rng.uniform(0.2, 2.0)) for _ in range(rng.integers(0, 4)))
    t2 = '|'.join('w%d' % rng.integers(0, 30) for _ in range(rng.integers(0, 5)))
    clk = '|'.join('#'.join(str(rng.integers(0, 11)) for _ in range(rng.integers(1, 4))) for _ in range(rng.integers(1, 4)))
    out.append('%d,%d,%s,%s,%.3f,%d,%s' % (rng.integers(0, 2), rng.integers(0, 1000), t1, t2, rng.uniform(0, 2),
                                          rng.integers(0, 11), clk))
  return out


def _worker(rank, port, ret, world, tmp, cuda=False):
  sys.path.insert(0, HERE)
  from test_dp_clip_gloo import _setup
  dev = _setup(rank, port, world, cuda)
  import ep_helpers
  from easyrec_b200.config import config_util
  from easyrec_b200.estimator import EasyRecEstimator
  from easyrec_b200.input import readers

  def make(text, ep):
    return EasyRecEstimator(text, device=dev, seed=5, world_size=world, rank=rank, embedding_parallel=ep)
  dp = make(CFG % b'', False)
  ep = make(CFG % b'train_distribute: EmbeddingParallelStrategy', None)
  assert ep.embedding_parallel and ep.input_layer.ep
  kinds = sorted(sc.kind for subs in ep.input_layer.subcalls.values() for sc in subs.values())
  assert kinds == ['mseq', 'single', 'tag'], kinds
  ep_helpers.copy_tables(dp.input_layer, ep.input_layer, rank, world)
  ep.model.load_state_dict(dp.model.state_dict())
  ep.trainer.dense_opt.flat_p.copy_(dp.trainer.dense_opt.flat_p)
  cfg = config_util.get_configs_from_pipeline_file(CFG % b'')
  losses = []
  for step in range(4):
    path = os.path.join(tmp, 'r%d_s%d.csv' % (rank, step))
    open(path, 'w').write('\n'.join(_lines(12, 1000 * rank + step)) + '\n')
    (feats, labels), = list(readers.CSVInput(cfg, dp.input_layer, path))
    feats, labels = readers.to_device(feats, labels, dev)
    l_dp, _ = dp.trainer.train_step(feats, labels)
    l_ep, _ = ep.trainer.train_step(feats, labels)
    losses.append((float(l_dp), float(l_ep)))
    assert abs(float(l_dp) - float(l_ep)) < 1e-5, losses
  worst = ep_helpers.compare(dp.input_layer, ep.input_layer, rank, world, 2e-6)
  d = float((dp.trainer.dense_opt.flat_p - ep.trainer.dense_opt.flat_p).abs().max())
  assert d < 1e-5, d
  assert losses[-1][0] != losses[0][0]
  ep.input_layer.check_exchange()
  ret[rank] = worst
  if cuda:
    dist.barrier()
    os._exit(0)
  dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('world', [2, 3])
def test_embedding_parallel_over_tag_and_multi_valued_sequence_slots_gloo(world, tmp_path):
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(_free_port(), ret, world, str(tmp_path)), nprocs=world, join=True)
  assert len(ret) == world
