"""CPU, world_size 2 over gloo, kernel doubles: data parallel over replicated tables with MULTI-VALUED slots (TagFeatures:
CSR lookups, kv weights, mean / sum combiners) - the segment of every lookup is all-gathered with the rows and the
gradient matrices, rank r's segments following rank r-1's in the gathered plan.  Without batch norm the data-parallel
step over two batches of B equals ONE process training on the concatenated batch of 2B (mean loss, averaged gradients):
that run is the reference here; replicas stay bit-identical.  (No embedding regulariser in this config: it is a SUM over the
batch's looked-up rows inside each worker's loss, so averaging the workers' gradients halves it against the concatenated
batch - in the reference under Horovod as well.)"""
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
train_config { optimizer_config { adagrad_optimizer { learning_rate { constant_learning_rate { learning_rate: 0.1 } } } } }
data_config { batch_size: %d input_type: CSVInput separator: "," label_fields: "label"
  input_fields { input_name: "label" input_type: FLOAT } input_fields { input_name: "u" input_type: INT64 }
  input_fields { input_name: "t1" input_type: STRING } input_fields { input_name: "t2" input_type: STRING }
  input_fields { input_name: "x" input_type: FLOAT } }
feature_config {
  features { input_names: "u" feature_type: IdFeature embedding_dim: 4 num_buckets: 9 embedding_name: "e" }
  features { input_names: "t1" feature_type: TagFeature embedding_dim: 4 num_buckets: 9 embedding_name: "e" separator: "|"
             kv_separator: ":" combiner: "mean" }
  features { input_names: "t2" feature_type: TagFeature embedding_dim: 4 hash_bucket_size: 13 separator: "|" combiner: "sum" }
  features { input_names: "x" feature_type: RawFeature embedding_dim: 4 min_val: 0.0 max_val: 2.0 } }
model_config { model_class: "DeepFM"
  feature_groups { group_name: "deep" feature_names: ["u", "t1", "t2", "x"] wide_deep: DEEP }
  feature_groups { group_name: "wide" feature_names: ["u", "t2"] wide_deep: WIDE }
  deepfm { dnn { hidden_units: [8] use_bn: false } final_dnn { hidden_units: [4] use_bn: false } l2_regularization: 1e-3 }
  }
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
    t1 = '|'.join('%d:%.2f' % (rng.integers(0, 9), rng.uniform(0.2, 2.0)) for _ in range(rng.integers(0, 4)))
    t2 = '|'.join('w%d' % rng.integers(0, 30) for _ in range(rng.integers(0, 5)))
    out.append('%d,%d,%s,%s,%.3f' % (rng.integers(0, 2), rng.integers(0, 9), t1, t2, rng.uniform(0, 2)))
  return out


def _batch(cfg_bytes, B, lines, tmp):
  from easyrec_b200 import builder
  from easyrec_b200.config import config_util
  from easyrec_b200.input import readers
  cfg = config_util.get_configs_from_pipeline_file(cfg_bytes)
  il, _, _ = builder.build_model(cfg, B, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  path = os.path.join(tmp, 'b%d_%d.csv' % (B, os.getpid()))
  open(path, 'w').write('\n'.join(lines) + '\n')
  (feats, labels), = list(readers.CSVInput(cfg, il, path))
  return feats, labels


def _worker(rank, port, ret, world, tmp, cuda=False):
  sys.path.insert(0, HERE)
  from test_dp_clip_gloo import _setup
  dev = _setup(rank, port, world, cuda)
  from easyrec_b200.estimator import EasyRecEstimator
  from easyrec_b200.input import readers
  B = 12
  parts = [_lines(B, 50 + r) for r in range(world)]
  dp = EasyRecEstimator(CFG % B, device=dev, seed=3, world_size=world, rank=rank, embedding_parallel=False)
  one = EasyRecEstimator(CFG % (B * world), device=dev, seed=3)          # the same model on the concatenated batch
  for d, a in dp.input_layer.arenas.items():
    one.input_layer.arenas[d].storage.copy_(a.storage)
  one.model.load_state_dict(dp.model.state_dict())
  one.trainer.dense_opt.flat_p.copy_(dp.trainer.dense_opt.flat_p)
  mine = readers.to_device(*_batch(CFG % B, B, parts[rank], tmp), dev)
  whole = readers.to_device(*_batch(CFG % (B * world), B * world, sum(parts, []), tmp), dev)
  for step in range(3):
    dp.trainer.train_step(*mine)
    one.trainer.train_step(*whole)
  worst = 0.0
  for d, a in dp.input_layer.arenas.items():
    worst = max(worst, float((a.storage - one.input_layer.arenas[d].storage).abs().max()))
  dworst = float((dp.trainer.dense_opt.flat_p - one.trainer.dense_opt.flat_p).abs().max())
  ret[rank] = (worst, dworst, float(sum(a.storage.double().sum() for a in dp.input_layer.arenas.values())),
               float(dp.trainer.dense_opt.flat_p.double().sum()))
  if cuda:
    dist.barrier()
    os._exit(0)
  dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_data_parallel_over_tag_slots_equals_one_process_on_the_concatenated_batch(tmp_path):
  world = 2
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(_free_port(), ret, world, str(tmp_path)), nprocs=world, join=True)
  assert len(ret) == world
  for worst, dworst, _, _ in ret.values():
    assert worst < 2e-6 and dworst < 2e-6, dict(ret)          # tables (weights and accumulators) and dense parameters
  assert len(set(v[2:] for v in ret.values())) == 1, dict(ret)   # replicas bit-identical
