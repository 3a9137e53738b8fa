"""For the reference sample configs that build here: fields set in the config that the subset schema dropped
(parsed with the full reference schema vs the subset), most frequent first.  Needs /root/reference.
python tools/ignored_config_fields.py"""
import glob, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyrec_b200 import builder
from easyrec_b200.config import config_util, proto_loader
REF='/root/reference'
full = proto_loader.load_schema(sorted(glob.glob(os.path.join(REF, 'easy_rec/python/protos/*.proto'))), virtual_name='full_ref.proto')
paths = sorted(glob.glob(os.path.join(REF, 'samples/model_config/*.config'))) + sorted(glob.glob(os.path.join(REF, 'examples/configs/*.config')))
def walk(msg, prefix, out):
  for fd, v in msg.ListFields():
    p = prefix + '.' + fd.name
    if fd.type == fd.TYPE_MESSAGE and not fd.message_type.GetOptions().map_entry:
      items = list(v) if getattr(fd,'is_repeated',False) else [v]
      for it in items:
        out.add(p); walk(it, p, out)
    else:
      out.add(p)
ign = collections.Counter()
for p in paths:
  try:
    cfg = config_util.get_configs_from_pipeline_file(p)
    builder.build_model(cfg, 8, 'cpu', cpu_generator=torch.Generator().manual_seed(0))
  except Exception:
    continue
  a, b = set(), set()
  walk(config_util.get_configs_from_pipeline_file(p, schema=full), '', a)
  walk(cfg, '', b)
  for f in a - b:
    ign[f] += 1
for f, n in ign.most_common(60):
  print(n, f)
