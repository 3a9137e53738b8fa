"""MultiTaskModel over a configurable backbone (reference: model/multi_task_model.py:28-141, 201-280):
the backbone yields one tensor per task (e.g. the MMoE block) or a single shared tensor; each
`model_params.task_towers` entry adds its tower DNN + dense(num_class=1); loss = sum_t weight_t * sigmoid CE."""
import torch
from torch import nn

from easyrec_b200 import embedding as E
from easyrec_b200 import layers as L
from easyrec_b200 import model as registry
from easyrec_b200.backbone import Backbone, regularised_groups
from easyrec_b200.model.rank_model import RankModel


@registry.register('MultiTaskModel')
class MultiTaskBackboneModel(RankModel):

  @classmethod
  def from_config(cls, model_config, input_layer, generator=None):
    if not model_config.HasField('backbone'):
      raise NotImplementedError('model_class MultiTaskModel needs a `backbone`')
    return cls(model_config, input_layer, generator=generator)

  def __init__(self, model_config, input_layer, generator=None):
    super().__init__()
    self.input_layer = input_layer
    self.backbone = Backbone(model_config.backbone, input_layer, input_layer.batch_size, generator)
    mp = model_config.model_params
    towers = list(mp.task_towers)
    assert towers, 'model_params.task_towers is empty'
    for t in towers:
      if t.relation_tower_names or t.num_class != 1:
        raise NotImplementedError('task tower %s: relation towers / multi-class' % t.tower_name)
    self.l2_reg = mp.l2_regularization
    self.embedding_reg = model_config.embedding_regularization
    self.tower_names = [t.tower_name for t in towers]
    self.label_names = [t.label_name if t.HasField('label_name') else None for t in towers]
    self.task_weights = [float(t.weight) for t in towers]
    self.per_task = self.backbone.n_outputs == len(towers) and self.backbone.n_outputs > 1
    d = self.backbone.out_dims[0] if self.per_task else self.backbone.out_dim
    self.tower_dnn = nn.ModuleList()
    self.tower_out = nn.ModuleList()
    for t in towers:
      units = L.units_of(t.dnn) if t.HasField('dnn') else []
      self.tower_dnn.append(L.DNN(d, units, generator=generator) if units else nn.Identity())
      self.tower_out.append(L.Dense(units[-1] if units else d, 1, generator))
    self.groups = regularised_groups(model_config.backbone)

  def forward(self, features):
    g = self.input_layer.lookup(features)
    self._emb_outputs = tuple(g[name][0] for name in self.groups)
    out = self.backbone(g)
    if self.per_task:
      xs = list(out)
    else:
      x = torch.cat(list(out), dim=-1) if isinstance(out, (list, tuple)) else out
      xs = [x] * len(self.tower_out)
    logits = [o(d(x))[:, 0] for x, d, o in zip(xs, self.tower_dnn, self.tower_out)]
    return torch.stack(logits, dim=1)   # [B, n_task]

  def loss(self, logits, labels, sample_weight=None):
    total = 0.0
    probs = []
    cols = getattr(self, 'label_cols', None) or list(range(len(self.task_weights)))
    for t, w in enumerate(self.task_weights):
      lt = (getattr(self, 'task_loss_types', None) or ['CLASSIFICATION'] * len(self.task_weights))[t]
      ce, p = self.data_loss(logits[:, t].contiguous(), labels[:, cols[t]].contiguous(), sample_weight, loss_type=lt)
      total = total + w * ce
      probs.append(p)
    return total + self.embedding_reg_loss(self._emb_outputs), torch.stack(probs, dim=1)
