"""Shared pieces of the rank models (reference: model/rank_model.py:57-151, 213-335;
builders/loss_builder.py:36-55; layers/input_layer.py:369-375 embedding regularisation)."""
import torch
from torch import nn

from easyrec_b200 import embedding as E


class RankModel(nn.Module):
  """num_class == 1 CLASSIFICATION head: logits -> sigmoid, tf.losses.sigmoid_cross_entropy."""

  l2_reg = 0.0
  embedding_reg = 0.0

  @staticmethod
  def wide_output_dim(model_config):
    return 1

  def l2_of(self, name, param):
    """kernel_regularizer = l2_regularizer(l2_regularization) on dense kernels (layers/dnn.py:57-62);
    biases, batch-norm and cross-layer vectors are not regularised."""
    return self.l2_reg if name.endswith('kernel') else 0.0

  def embedding_reg_loss(self, tensors):
    """l2_regularizer(embedding_regularization) over the LOOKED-UP group outputs
    (layers/input_layer.py:369-375, compat/regularizers.py): scale * sum(x^2) / 2."""
    if self.embedding_reg <= 0 or not tensors:
      return 0.0
    # (a group that carries in-group target attention names the looked-up tensors it is made of: `_er_reg`)
    flat = [r for t in tensors for r in getattr(t, '_er_reg', [t])]
    return self.embedding_reg * 0.5 * sum((t * t).sum() for t in flat)

  accepts_sample_weight = True

  @staticmethod
  def weighted_ce(logits, labels, sample_weight=None):
    """tf.losses.sigmoid_cross_entropy(labels, logits, weights=sample_weight) (model/rank_model.py:213-269,
    builders/loss_builder.py:36-39), reduction SUM_BY_NONZERO_WEIGHTS: sum(w * ce) / count_nonzero(w).  The kernel
    divides by the batch size, so the weights carry the ratio B / count_nonzero(w) (two tiny device ops)."""
    w = None
    if sample_weight is not None:
      sw = sample_weight.to(torch.float32).reshape(-1)
      nnz = (sw != 0).sum().clamp(min=1).to(torch.float32)
      w = (sw * (float(sw.numel()) / nnz)).contiguous()
    return E.sigmoid_cross_entropy(logits, labels, weights=w)

  loss_type = 'CLASSIFICATION'   # model_config.loss_type (set by builder.build_model)

  def data_loss(self, logits, labels, sample_weight=None, loss_type=None):
    """(loss, predictions) of one binary head by its loss_type - model_config.loss_type, or the task tower's
    (model/rank_model.py:75-129, 213-269, builders/loss_builder.py:36-55): CLASSIFICATION = sigmoid cross entropy,
    predictions `probs`; L2_LOSS / SIGMOID_L2_LOSS = tf.losses.mean_squared_error(labels, y, weights) with
    y = logits / sigmoid(logits), predictions `y`.  Both reduce by SUM_BY_NONZERO_WEIGHTS."""
    lt = loss_type or getattr(self, 'loss_type', 'CLASSIFICATION')
    if lt == 'CLASSIFICATION':
      return RankModel.weighted_ce(logits, labels, sample_weight)
    if lt not in ('L2_LOSS', 'SIGMOID_L2_LOSS'):
      raise NotImplementedError('loss_type %s' % lt)
    y = torch.sigmoid(logits) if lt == 'SIGMOID_L2_LOSS' else logits
    d = y - labels.to(y.dtype)
    if sample_weight is None:
      return (d * d).mean(), y.detach()
    sw = sample_weight.to(torch.float32).reshape(-1)
    nnz = (sw != 0).sum().clamp(min=1).to(torch.float32)
    return (sw * d * d).sum() / nnz, y.detach()

  def loss(self, logits, labels, sample_weight=None):
    ce, probs = self.data_loss(logits, labels, sample_weight)
    return ce + self.embedding_reg_loss(getattr(self, '_emb_outputs', ())), probs
