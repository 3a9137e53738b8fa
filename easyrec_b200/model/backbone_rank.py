"""RankModel over a configurable backbone (reference: model/rank_model.py:36-56 `build_predict_graph` with
`self.backbone`, model/easy_rec_model.py:99-118 backbone construction): model_class "RankModel" +
`backbone { blocks ... }` + `model_params { l2_regularization }`; a dense(num_class) head is added when the
backbone output is wider than one logit (rank_model.py:50-54)."""
import torch

from easyrec_b200 import layers as L
from easyrec_b200 import model as registry
from easyrec_b200.backbone import Backbone, regularised_groups
from easyrec_b200.model.rank_model import RankModel


@registry.register('RankModel')
class BackboneRankModel(RankModel):

  @classmethod
  def from_config(cls, model_config, input_layer, generator=None):
    if not model_config.HasField('backbone'):
      raise NotImplementedError('model_class RankModel needs a `backbone` (rank_model.py:40-43)')
    return cls(model_config, input_layer, generator=generator)

  def __init__(self, model_config, input_layer, generator=None):
    super().__init__()
    self.input_layer = input_layer
    self.backbone = Backbone(model_config.backbone, input_layer, input_layer.batch_size, generator)
    self.l2_reg = model_config.model_params.l2_regularization if model_config.HasField('model_params') else 0.0
    self.embedding_reg = model_config.embedding_regularization
    self.output = L.Dense(self.backbone.out_dim, 1, generator) if self.backbone.out_dim != 1 else None
    self.groups = regularised_groups(model_config.backbone)

  def forward(self, features):
    g = self.input_layer.lookup(features)
    self._emb_outputs = tuple(g[name][0] for name in self.groups)
    out = self.backbone(g)
    if isinstance(out, (list, tuple)):
      out = torch.cat(list(out), dim=-1)
    if self.output is not None:
      out = self.output(out)
    return out[:, 0]
