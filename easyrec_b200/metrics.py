"""Evaluation metrics of the rank models (model/rank_model.py:334-470): AUC and the grouped AUCs
(core/metrics.py:59-106 `_separated_auc_impl` behind `gauc` / `session_auc`, :260-298).

Exact, on the host, over the predictions an `evaluate()` pass collected: `auc` is the Mann-Whitney statistic
with average ranks over ties (= sklearn.metrics.roc_auc_score, which the reference calls per group; its global
`tf.metrics.auc` is a 200-threshold approximation of the same quantity)."""
import numpy as np


def _tie_averaged_ranks(scores):
  """1-based ranks, equal scores sharing the mean of their positions."""
  order = np.argsort(scores, kind='mergesort')
  s = scores[order]
  first = np.concatenate([[True], s[1:] != s[:-1]])
  start = np.flatnonzero(first)
  end = np.concatenate([start[1:], [len(s)]])
  mean_rank = 0.5 * (start + end - 1) + 1
  ranks = np.empty(len(s), np.float64)
  ranks[order] = np.repeat(mean_rank, end - start)
  return ranks


def auc(labels, scores):
  labels = np.asarray(labels).reshape(-1).astype(np.float64)
  scores = np.asarray(scores).reshape(-1).astype(np.float64)
  n_pos = labels.sum()
  n_neg = len(labels) - n_pos
  if n_pos == 0 or n_neg == 0:
    return float('nan')
  ranks = _tie_averaged_ranks(scores)
  return float((ranks[labels > 0].sum() - n_pos * (n_pos + 1) / 2) / (n_pos * n_neg))


def gauc(labels, scores, keys, reduction='mean'):
  """AUC per key (user / session), groups with a single class skipped, averaged with weight 1 ('mean'), the
  group's sample count ('mean_by_sample_num') or its positive count ('mean_by_positive_num'); 0.0 when no group
  qualifies.  float32 like the reference's value op."""
  assert reduction in ['mean', 'mean_by_sample_num', 'mean_by_positive_num'], \
      'reduction method must in mean | mean_by_sample_num | mean_by_positive_num'
  labels = np.asarray(labels).reshape(-1).astype(np.float64)
  scores = np.asarray(scores).reshape(-1).astype(np.float64)
  keys = np.asarray(keys).reshape(-1)
  order = np.argsort(keys, kind='mergesort')
  k = keys[order]
  bounds = np.flatnonzero(np.concatenate([[True], k[1:] != k[:-1], [True]]))
  metrics, weights = [], []
  for a, b in zip(bounds[:-1], bounds[1:]):
    idx = order[a:b]
    lab = labels[idx]
    pos = lab.sum()
    if pos == 0 or pos == len(lab):
      continue
    metrics.append(auc(lab, scores[idx]))
    weights.append({'mean': 1, 'mean_by_sample_num': len(lab), 'mean_by_positive_num': pos}[reduction])
  if not metrics:
    return np.float32(0.0)
  return np.average(metrics, weights=weights).astype(np.float32)
