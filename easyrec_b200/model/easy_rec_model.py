"""The reference's model contract on top of the fused path (model/easy_rec_model.py:51-183, registry
utils/load_class.py:203-222 / `EasyRecModel.create_class`, main.py:137):

    cls = EasyRecModel.create_class(model_config.model_class)
    model = cls(model_config, feature_configs, features, labels, is_training, input_layer=il)
    predictions = model.build_predict_graph()     # {'logits', 'probs', 'y'?} (+ '_<tower>' for multi-task models)
    losses = model.build_loss_graph()             # {'cross_entropy_loss'[...], 'regularization_loss'}
    metrics = model.build_metric_graph(eval_config)
    names = model.get_outputs()

The reference builds a TF graph once and feeds it; here the same four calls evaluate the torch model eagerly on the
`features` / `labels` bound at construction (rebind with `set_inputs`).  Prediction and loss keys follow
model/rank_model.py:75-129,213-269 and model/multi_task_model.py:124-141,201-280.  The arenas / tables live in the
InputLayer, which the caller passes in (the estimator owns it); without one it is built from the feature configs."""
import torch

from easyrec_b200 import metrics as M
from easyrec_b200 import model as registry


class EasyRecModel(object):
  """mixin of every registered model class: adds the reference constructor form and the graph-building calls"""

  @staticmethod
  def create_class(name):
    """model_class name -> class whose constructor takes the reference's arguments."""
    base = registry.get_model_class(name)
    cached = _WRAPPED.get(name)
    if cached is None:
      cached = type(name, (_Bound, base), {'_base_cls': base, '__doc__': base.__doc__})
      _WRAPPED[name] = cached
    return cached


_WRAPPED = {}


class _Bound(EasyRecModel):

  def __new__(cls, model_config, feature_configs=None, features=None, labels=None, is_training=False, input_layer=None,
              generator=None, **kw):
    if input_layer is None:
      raise ValueError('pass input_layer= (the estimator owns the embedding arenas: builder.build_model / '
                       'EasyRecEstimator.input_layer)')
    self = cls._base_cls.from_config(model_config, input_layer, generator=generator)
    self.__class__ = cls
    return self

  def __init__(self, model_config, feature_configs=None, features=None, labels=None, is_training=False, input_layer=None,
               generator=None, **kw):
    self._model_config = model_config
    self._feature_configs = feature_configs
    self._prediction_dict = {}
    self._loss_dict = {}
    self._metric_dict = {}
    self.set_inputs(features, labels, is_training)

  def set_inputs(self, features, labels=None, is_training=False):
    self._feature_dict = features
    self._labels = labels
    self._is_training = bool(is_training)
    self._is_predicting = labels is None
    self._logits = None
    return self

  # -- names of the towers / outputs -------------------------------------------------------------------------
  def _towers(self):
    return list(getattr(self, 'tower_names', []) or [])

  def build_predict_graph(self):
    self.train(self._is_training)
    logits = self(self._feature_dict)
    self._logits = logits
    d = {}
    towers = self._towers()
    if logits.dim() == 2 and towers and logits.shape[1] == len(towers):      # multi-task: one column per tower
      for t, name in enumerate(towers):
        d['logits_' + name] = logits[:, t]
        d['probs_' + name] = torch.sigmoid(logits[:, t])
    elif logits.dim() == 2:      # list-wise match model: the [B, B] similarity matrix
      d['logits'] = logits
      d['probs'] = torch.softmax(logits, dim=1)
    else:
      d['logits'] = logits
      d['probs'] = torch.sigmoid(logits)
    self._prediction_dict = d
    return d

  def build_loss_graph(self):
    if self._logits is None:
      self.build_predict_graph()
    if self._labels is None:
      raise ValueError('build_loss_graph needs labels')
    total, _ = self.loss(self._logits, self._labels)
    reg = None
    if hasattr(self, 'regularization_loss'):
      reg = self.regularization_loss()
    elif hasattr(self, 'embedding_reg_loss'):
      reg = self.embedding_reg_loss(getattr(self, '_emb_outputs', ()))
    d = {}
    if reg is not None and (torch.is_tensor(reg) or reg != 0.0):
      d['regularization_loss'] = reg
      d['cross_entropy_loss'] = total - reg
    else:
      d['cross_entropy_loss'] = total
    self._loss_dict = d
    self._total_loss = total
    return d

  def build_metric_graph(self, eval_config):
    """eval_config.metrics_set on the bound batch, by the reference's streaming definitions (metrics.MetricSet: auc =
    tf.metrics.auc at AUC.num_thresholds thresholds, max_f1, the mean / root-mean errors; '<metric>_<tower_name>' per
    task tower).  The estimator keeps one MetricSet over a whole evaluate() pass; this call evaluates it on one batch."""
    if self._logits is None:
      self.build_predict_graph()
    out = {}
    heads = M.heads_of(self)
    if self._labels is not None and heads:
      mset = M.MetricSet(eval_config.metrics_set if eval_config is not None else [], heads, self._logits.device)
      mset.update(self._logits, self._labels)
      out = mset.result()
    self._metric_dict = out
    return out

  def get_outputs(self):
    towers = self._towers()
    if towers and getattr(self, '_logits', None) is not None and self._logits.dim() == 2 and \
        self._logits.shape[1] == len(towers):
      return [k + '_' + n for n in towers for k in ('probs', 'logits')]
    if towers and self._logits is None and len(towers) > 1:
      return [k + '_' + n for n in towers for k in ('probs', 'logits')]
    return ['probs', 'logits']
