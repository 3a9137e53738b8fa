"""MatchModel over a configurable backbone (reference: model/match_model.py:28-139 with `self.backbone`,
samples/model_config/dssm_on_taobao_backbone.config): the backbone's output_blocks are the user and item
tower outputs (`model_params.user_tower_idx_in_output / item_tower_idx_in_output`); similarity, in-batch
softmax / point-wise loss and the positive-similarity regulariser are DSSM's (model/dssm.py)."""
import torch
from torch import nn

from easyrec_b200 import interactions as I
from easyrec_b200 import model as registry
from easyrec_b200.backbone import Backbone, regularised_groups
from easyrec_b200.model.dssm import DSSM


@registry.register('MatchModel')
class MatchBackboneModel(DSSM):

  @classmethod
  def from_config(cls, model_config, input_layer, generator=None):
    if not model_config.HasField('backbone'):
      raise NotImplementedError('model_class MatchModel needs a `backbone`')
    return cls(model_config, input_layer, generator=generator)

  def __init__(self, model_config, input_layer, generator=None):
    nn.Module.__init__(self)
    self.input_layer = input_layer
    self.backbone = Backbone(model_config.backbone, input_layer, input_layer.batch_size, generator)
    assert self.backbone.n_outputs >= 2, 'a match backbone outputs the user and the item tower'
    mp = model_config.model_params
    self.u_idx, self.i_idx = mp.user_tower_idx_in_output, mp.item_tower_idx_in_output
    simi = mp.DESCRIPTOR.fields_by_name['simi_func'].enum_type.values_by_number[mp.simi_func].name
    loss_name = model_config.DESCRIPTOR.fields_by_name['loss_type'].enum_type.values_by_number[
        model_config.loss_type].name
    self.cosine, self.temperature = simi == 'COSINE', float(mp.temperature)
    self.listwise = loss_name == 'SOFTMAX_CROSS_ENTROPY'
    self.scale_simi = mp.scale_simi
    if self.scale_simi:
      self.sim_w = nn.Parameter(torch.ones(1))
      self.sim_b = nn.Parameter(torch.zeros(1))
    self.item_id = None
    self.l2_reg = mp.l2_regularization
    self.embedding_reg = model_config.embedding_regularization
    self.groups = regularised_groups(model_config.backbone)

  def towers(self, features):
    g = self.input_layer.lookup(features)
    self._emb_outputs = tuple(g[name][0] for name in self.groups)
    outs = self.backbone(g)
    u, i = outs[self.u_idx], outs[self.i_idx]
    if self.cosine:
      u, i = I.l2_normalize(u.contiguous()), I.l2_normalize(i.contiguous())
    return u, i
