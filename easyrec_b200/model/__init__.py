"""Model registry: model_class name -> class (utils/load_class.py:203-222, main.py:137)."""

_REGISTRY = {}


def register(name):
  def deco(cls):
    _REGISTRY[name] = cls
    return cls
  return deco


def get_model_class(name):
  from easyrec_b200.model import (backbone_rank, dbmtl, dcn, deepfm, dlrm, dssm, match_backbone, mmoe, multi_task_backbone,  # noqa: F401
                                  multi_tower_din, wide_and_deep)
  if name not in _REGISTRY:
    raise KeyError('model_class %r is outside the hot-path scope (have: %s)' % (name, sorted(_REGISTRY)))
  return _REGISTRY[name]
