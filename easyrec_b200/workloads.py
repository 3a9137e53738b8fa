"""Synthetic workloads of BASELINE.json `configs` (shapes from SURVEY.md section 8d).

C2 = DeepFM Criteo-shape: 13 RawFeature (embedding_dim 16, min/max of
examples/configs/deepfm_on_criteo.config:241-330) + 26 IdFeature hashed into ONE shared
table of V rows (embedding_name 'embedding', as samples/model_config/
dlrm_on_criteo_parquet_ep.config:319-324), emb 16, batch 8192, DNN [256,128,64] +
final_dnn [256,128,64], wide_output_dim 1.
"""
import collections

import numpy as np
import torch

from easyrec_b200 import _lib
from easyrec_b200 import input_layer as IL
from easyrec_b200.model.deepfm import DeepFM

CRITEO_MAX = [5775.0, 257675.0, 65535.0, 969.0, 23159456.0, 431037.0, 56311.0, 6047.0, 29019.0,
              46.0, 231.0, 4008.0, 7393.0]
CRITEO_MIN = [0.0, -3.0] + [0.0] * 11


def criteo_features(vocab, emb_dim=16, shared_table=True):
  feats = []
  for i in range(13):
    feats.append(IL.raw_feature('F%d' % (i + 1), emb_dim, CRITEO_MIN[i], CRITEO_MAX[i]))
  per = vocab if shared_table else max(vocab // 26, 1)
  for i in range(26):
    feats.append(IL.id_feature('C%d' % (i + 1), emb_dim, hash_bucket_size=per,
                               embedding_name='embedding' if shared_table else ''))
  names = [f.name for f in feats]
  groups = collections.OrderedDict([('deep', dict(features=names, wide=False)),
                                    ('wide', dict(features=names, wide=True))])
  return feats, groups


def build_deepfm_criteo(batch_size, vocab, device, emb_dim=16, shared_table=True, seed=20240,
                        emb_opt=_lib.OPT_ADAGRAD, dnn=(256, 128, 64), final=(256, 128, 64),
                        l2_reg=1e-5, emb_reg=1e-5):
  feats, groups = criteo_features(vocab, emb_dim, shared_table)
  gen = torch.Generator(device=device).manual_seed(seed)
  il = IL.InputLayer(feats, groups, batch_size, device, wide_output_dim=1,
                     embedding_optimizer=emb_opt, generator=gen)
  cpu_gen = torch.Generator().manual_seed(seed)
  model = DeepFM(il, list(dnn), list(final), l2_reg=l2_reg, embedding_reg=emb_reg,
                 generator=cpu_gen).to(device)
  return il, model


def criteo_batch(batch_size, seed, zipf_alpha=1.05, uniform=False):
  """Host-side synthetic batch: ids int64 [26*B] feature-major, dense fp32 [B,13], labels [B].

  ids ~ Zipf(alpha) over [0, 2^40) (uniform variant: worst case, U ~= L); dense ~ lognormal
  clipped to the config's [min, max]; labels ~ Bernoulli(0.25)."""
  rng = np.random.default_rng(seed)
  n = 26 * batch_size
  if uniform:
    ids = rng.integers(0, 2**40, n, dtype=np.int64)
  else:
    ids = (rng.zipf(zipf_alpha, n).astype(np.int64) - 1) % (2**40)
    # decorrelate slots: each slot sees its own permutation of the id space
    ids = ids * 26 + np.repeat(np.arange(26, dtype=np.int64), batch_size)
  dense = rng.lognormal(1.0, 2.0, (batch_size, 13)).astype(np.float32)
  dense = np.minimum(dense, np.array(CRITEO_MAX, np.float32))
  labels = (rng.uniform(size=batch_size) < 0.25).astype(np.float32)
  return ids, dense, labels
