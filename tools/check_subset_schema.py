"""Cross-check easyrec_b200/config/easyrec_subset.proto against the reference's full schema:
every message/field/enum value of the subset must exist in alibaba/EasyRec's protos with the same
number, type, label and default.  Needs /root/reference (or EASYREC_PROTO_DIR)."""
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyrec_b200.config import proto_loader as PL  # noqa: E402


def check(ref_dir):
  here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'easyrec_b200', 'config')
  sub = PL.load_schema([os.path.join(here, 'easyrec_subset.proto')], virtual_name='sub.proto')
  ref = PL.load_schema(sorted(glob.glob(os.path.join(ref_dir, '*.proto'))), virtual_name='ref.proto')
  problems = []

  def cmp_enum(se, re_, where):
    rv = {v.name: v.number for v in re_.values}
    for v in se.values:
      if rv.get(v.name) != v.number:
        problems.append('%s: enum value %s=%d, reference %r' % (where, v.name, v.number, rv.get(v.name)))

  def _attr(fd, attr):
    if attr != 'label':
      return getattr(fd, attr)
    if hasattr(fd, 'is_repeated'):     # protobuf >= 5.29 deprecates FieldDescriptor.label
      rep, req = fd.is_repeated, fd.is_required
      rep, req = (rep() if callable(rep) else rep), (req() if callable(req) else req)
      return 'repeated' if rep else 'required' if req else 'optional'
    return {fd.LABEL_REPEATED: 'repeated', fd.LABEL_REQUIRED: 'required'}.get(fd.label, 'optional')

  def cmp_msg(sd, rd):
    rf = {f.name: f for f in rd.fields}
    for f in sd.fields:
      r = rf.get(f.name)
      if r is None:
        problems.append('%s.%s: not in reference' % (sd.full_name, f.name))
        continue
      for attr in ('number', 'type', 'label'):
        a, b = _attr(f, attr), _attr(r, attr)
        if a != b:
          problems.append('%s.%s: %s %r != reference %r' % (sd.full_name, f.name, attr, a, b))
      if f.has_default_value != r.has_default_value or (f.has_default_value and f.default_value != r.default_value):
        problems.append('%s.%s: default %r != reference %r' % (sd.full_name, f.name, f.default_value, r.default_value))
      if f.message_type is not None and r.message_type is not None and f.message_type.name != r.message_type.name:
        problems.append('%s.%s: message type %s != %s' % (sd.full_name, f.name, f.message_type.name, r.message_type.name))
      so = f.containing_oneof.name if f.containing_oneof else None
      ro = r.containing_oneof.name if r.containing_oneof else None
      if so != ro:
        problems.append('%s.%s: oneof %r != %r' % (sd.full_name, f.name, so, ro))
    for n in sd.nested_types:
      rn = {x.name: x for x in rd.nested_types}.get(n.name)
      if rn is None:
        problems.append('%s: nested message missing in reference' % n.full_name)
      else:
        cmp_msg(n, rn)
    for e in sd.enum_types:
      re_ = {x.name: x for x in rd.enum_types}.get(e.name)
      if re_ is None:
        problems.append('%s: nested enum missing in reference' % e.full_name)
      else:
        cmp_enum(e, re_, e.full_name)

  fd = sub._pool.FindFileByName('sub.proto')
  for name, sd in fd.message_types_by_name.items():
    try:
      rd = ref._pool.FindMessageTypeByName('protos.' + name)
    except KeyError:
      problems.append('message %s missing in reference' % name)
      continue
    cmp_msg(sd, rd)
  for name, se in fd.enum_types_by_name.items():
    try:
      cmp_enum(se, ref._pool.FindEnumTypeByName('protos.' + name), name)
    except KeyError:
      problems.append('enum %s missing in reference' % name)
  return problems


if __name__ == '__main__':
  d = os.environ.get('EASYREC_PROTO_DIR', '/root/reference/easy_rec/python/protos')
  if not os.path.isdir(d):
    print('reference protos not available')
    sys.exit(0)
  ps = check(d)
  print('\n'.join(ps) if ps else 'subset schema is consistent with the reference schema')
  sys.exit(1 if ps else 0)
