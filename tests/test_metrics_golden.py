"""CPU: grouped AUC (eval_config.metrics_set gauc / session_auc) against values produced by executing the
reference's `_separated_auc_impl` closures (tests/golden/make_metrics_golden.py), and AUC against sklearn."""
import json
import os

import numpy as np
import pytest

from easyrec_b200 import metrics

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_metrics.json')))


@pytest.mark.parametrize('reduction', ['mean', 'mean_by_sample_num', 'mean_by_positive_num'])
def test_gauc_matches_the_reference_implementation(reduction):
  got = metrics.gauc(GOLD['labels'], GOLD['predictions'], GOLD['keys'], reduction)
  assert got.dtype == np.float32
  assert abs(float(got) - GOLD['gauc'][reduction]) < 1e-6
  assert metrics.gauc(np.ones(10), GOLD['predictions'][:10], GOLD['keys'][:10]) == GOLD['gauc_all_single_class'] == 0.0


def test_auc_is_the_tie_aware_mann_whitney_statistic():
  from sklearn.metrics import roc_auc_score
  rng = np.random.default_rng(0)
  for n, decimals in [(1000, 1), (5000, 3), (50, 8)]:
    labels = (rng.uniform(size=n) < 0.3).astype(np.int64)
    scores = np.round(rng.uniform(size=n) * 0.5 + labels * 0.2, decimals)
    assert abs(metrics.auc(labels, scores) - roc_auc_score(labels, scores)) < 1e-12
  assert np.isnan(metrics.auc(np.zeros(5), np.arange(5)))


def test_evaluate_resolves_group_metrics_from_eval_config():
  """eval_config { metrics_set { gauc { uid_field } } }: the key is the packed id column of that input field."""
  import sys
  import torch
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  from test_config import MINI
  from easyrec_b200 import builder
  from easyrec_b200.config import config_util
  from easyrec_b200.estimator import EasyRecEstimator
  cfg = config_util.get_configs_from_pipeline_file(
      MINI + b'eval_config { metrics_set { auc {} } metrics_set { gauc { uid_field: "C1" reduction: "mean_by_sample_num" } } '
      b'metrics_set { session_auc { session_id_field: "C1" } } }')
  est = EasyRecEstimator.__new__(EasyRecEstimator)     # no device: only the config plumbing is under test
  est._pipeline_config = cfg
  est.input_layer, _, _ = builder.build_model(cfg, 32, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  assert est._group_fields() == [('gauc', 0, 'mean_by_sample_num'), ('session_auc', 0, 'mean')]
  bad = config_util.get_configs_from_pipeline_file(MINI + b'eval_config { metrics_set { gauc { uid_field: "F1" } } }')
  est._pipeline_config = bad
  with pytest.raises(ValueError, match='not a single-valued id feature'):
    est._group_fields()
