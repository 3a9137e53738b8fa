"""Pipeline-config loading, mirroring easy_rec/python/utils/config_util.py:46-79,192-340.

`get_configs_from_pipeline_file` returns an `EasyRecConfig` message parsed with
google.protobuf.text_format from an UNMODIFIED EasyRec `.config` file.  With the built-in subset
schema, fields the hot path does not consume are skipped (and listed by `unknown_fields`); with
EASYREC_PROTO_DIR set the reference's complete schema is loaded and parsing is strict.
"""
import json
import os
import re

from google.protobuf import text_format

from easyrec_b200.config import proto_loader


def _is_full_schema(schema):
  return schema.has('ExportConfig')


def get_configs_from_pipeline_file(pipeline_config_path, schema=None):
  """utils/config_util.py:46-79."""
  if isinstance(pipeline_config_path, bytes):
    text, path = pipeline_config_path, '<bytes>'
  else:
    path = pipeline_config_path
    assert os.path.exists(path), 'pipeline_config_path [%s] not exists' % path
    with open(path, 'rb') as f:  # byte-transparent: configs may hold raw \x02/\x03 separators
      text = f.read()
  schema = schema or proto_loader.default_schema()
  cfg = schema.EasyRecConfig()
  text_format.Merge(text, cfg, allow_unknown_field=not _is_full_schema(schema))
  return auto_expand_share_feature_configs(cfg)


def auto_expand_names(input_name):
  """field[1-3] -> [field1, field2, field3]  (utils/config_util.py:116-135)."""
  m = re.match(r'([a-zA-Z_]+)\[([0-9]+)-([0-9]+)\]', input_name)
  if m:
    return ['%s%d' % (m.group(1), t) for t in range(int(m.group(2)), int(m.group(3)) + 1)]
  return [input_name]


def auto_expand_share_feature_configs(pipeline_config):
  """utils/config_util.py:81-113: a FeatureConfig with `shared_names` stands for one more feature per shared name - a
  copy of the config (without its own input_names and shared_names) reading that input; the copies are appended to the
  feature list, the original keeps its own input.  Patterns in shared_names expand when
  data_config.auto_expand_input_fields is set."""
  feats = pipeline_config.feature_configs if len(pipeline_config.feature_configs) > 0 else pipeline_config.feature_config.features
  expand = bool(getattr(pipeline_config.data_config, 'auto_expand_input_fields', False))
  new = []
  for fc in list(feats):
    if len(fc.shared_names) == 0:
      continue
    names = []
    for n in fc.shared_names:
      names.extend(auto_expand_names(n) if expand else [n])
    del fc.shared_names[:]
    for n in names:
      c = type(fc)()
      c.CopyFrom(fc)
      del c.input_names[:]
      c.input_names.append(n)
      new.append(c)
  for c in new:
    feats.add().CopyFrom(c)
  # data_config.input_fields named by a pattern stand for one field per expanded name (input/input.py:62-75)
  dc = pipeline_config.data_config
  if expand and any(len(auto_expand_names(f.input_name)) > 1 for f in dc.input_fields):
    fields = [type(f)() for f in dc.input_fields]
    for dst, src in zip(fields, dc.input_fields):
      dst.CopyFrom(src)
    del dc.input_fields[:]
    for f in fields:
      for n in auto_expand_names(f.input_name):
        g = dc.input_fields.add()
        g.CopyFrom(f)
        g.input_name = n
  # feature_groups name their features with the same patterns: FeatureGroup._auto_expand_feature_name always expands
  # them (feature_column/feature_group.py:46-60)
  for g in pipeline_config.model_config.feature_groups:
    names = [x for n in g.feature_names for x in auto_expand_names(n)]
    if names != list(g.feature_names):
      del g.feature_names[:]
      g.feature_names.extend(names)
  return pipeline_config


def unknown_fields(pipeline_config_path, schema=None):
  """Top-level field names present in the file but absent from the loaded schema."""
  schema = schema or proto_loader.default_schema()
  if _is_full_schema(schema):
    return []
  with open(pipeline_config_path, 'rb') as f:
    text = f.read().decode('utf-8', errors='replace')
  names = set(re.findall(r'^\s*([A-Za-z_][\w]*)\s*[:{]', text, re.M))
  known = set()

  def walk(desc, seen):
    if desc.full_name in seen:
      return
    seen.add(desc.full_name)
    for f in desc.fields:
      known.add(f.name)
      if f.message_type is not None:
        walk(f.message_type, seen)

  walk(schema.EasyRecConfig.DESCRIPTOR, set())
  return sorted(n for n in names if n not in known)


def get_feature_configs(pipeline_config):
  """feature_configs (old) or feature_config.features (v2): utils/config_util.py `get_compatible_feature_configs`."""
  if len(pipeline_config.feature_configs) > 0:
    return list(pipeline_config.feature_configs)
  return list(pipeline_config.feature_config.features)


def edit_config(pipeline_config, edit_config_json):
  """Dotted-path overrides, the common subset of utils/config_util.py:192-340:
  {"train_config.num_steps": 10, "data_config.batch_size": 64,
   "model_config.deepfm.dnn.hidden_units[0]": 32}."""
  if isinstance(edit_config_json, str):
    edit_config_json = json.loads(edit_config_json)
  for path, value in edit_config_json.items():
    obj = pipeline_config
    parts = path.split('.')
    for i, part in enumerate(parts):
      m = re.match(r'^(\w+)(?:\[(\d+)\])?$', part)
      assert m, 'bad edit path: %s' % path
      name, idx = m.group(1), m.group(2)
      last = i == len(parts) - 1
      if idx is not None:
        seq = getattr(obj, name)
        if last:
          seq[int(idx)] = type(seq[int(idx)])(value)
        else:
          obj = seq[int(idx)]
      elif last:
        field = obj.DESCRIPTOR.fields_by_name[name]
        if field.enum_type is not None and isinstance(value, str):
          value = field.enum_type.values_by_name[value].number
        elif field.type in (field.TYPE_FLOAT, field.TYPE_DOUBLE):
          value = float(value)
        elif field.type == field.TYPE_BOOL:
          value = value in (True, 'true', 'True', 1)
        elif field.type == field.TYPE_STRING:
          value = str(value)
        else:
          value = int(value)
        setattr(obj, name, value)
      else:
        obj = getattr(obj, name)
  return pipeline_config
