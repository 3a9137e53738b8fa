"""Runtime proto2 schema loader (there is no protoc in the image and the reference ships no
*_pb2.py: scripts/gen_proto.sh:34 generates them with a downloaded protoc).

`load_schema(paths)` parses `.proto` text -- this repo's own subset schema
(easyrec_subset.proto) or, when EASYREC_PROTO_DIR points at an EasyRec checkout, the
reference's full `easy_rec/python/protos/*.proto` -- into ONE FileDescriptorProto, registers it
in a private DescriptorPool and returns message classes.  Pipeline configs
(`samples/model_config/*.config`) are then read with google.protobuf.text_format exactly as
`utils/config_util.py:46-79` does.

Supported grammar: syntax/package/import/option statements, message (nested), enum, oneof,
optional/required/repeated fields, map<k,v> fields, [default = x, packed = x, deprecated = x]
options, reserved/extensions ranges (ignored), // and /* */ comments.
"""
import os
import re

from google.protobuf import descriptor_pb2
from google.protobuf import descriptor_pool
from google.protobuf import message_factory

FD = descriptor_pb2.FieldDescriptorProto

_SCALARS = {
    'double': FD.TYPE_DOUBLE, 'float': FD.TYPE_FLOAT, 'int32': FD.TYPE_INT32,
    'int64': FD.TYPE_INT64, 'uint32': FD.TYPE_UINT32, 'uint64': FD.TYPE_UINT64,
    'sint32': FD.TYPE_SINT32, 'sint64': FD.TYPE_SINT64, 'fixed32': FD.TYPE_FIXED32,
    'fixed64': FD.TYPE_FIXED64, 'sfixed32': FD.TYPE_SFIXED32, 'sfixed64': FD.TYPE_SFIXED64,
    'bool': FD.TYPE_BOOL, 'string': FD.TYPE_STRING, 'bytes': FD.TYPE_BYTES,
}
_LABELS = {'optional': FD.LABEL_OPTIONAL, 'required': FD.LABEL_REQUIRED,
           'repeated': FD.LABEL_REPEATED}

_TOKEN = re.compile(
    r'''\s+|//[^\n]*|/\*.*?\*/|("(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*')|([A-Za-z_][\w.]*)|'''
    r'''([-+]?(?:0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|inf|nan))|(.)''', re.S)


def _tokenize(text):
  out = []
  for m in _TOKEN.finditer(text):
    s, ident, num, punct = m.groups()
    if s is not None:
      out.append(('str', s))
    elif ident is not None:
      out.append(('id', ident))
    elif num is not None:
      out.append(('num', num))
    elif punct is not None:
      out.append(('p', punct))
  return out


def _unquote(s):
  body = s[1:-1]
  return bytes(body, 'utf-8').decode('unicode_escape').encode('latin-1').decode('utf-8') \
      if '\\' in body else body


class _Parser(object):

  def __init__(self, text, fname):
    self.toks = _tokenize(text)
    self.i = 0
    self.fname = fname

  def peek(self):
    return self.toks[self.i] if self.i < len(self.toks) else ('eof', '')

  def next(self):
    t = self.peek()
    self.i += 1
    return t

  def expect(self, val):
    t = self.next()
    if t[1] != val:
      raise ValueError('%s: expected %r, got %r (token %d)' % (self.fname, val, t[1], self.i))
    return t

  def skip_statement(self):
    depth = 0
    while True:
      t = self.next()
      if t[0] == 'eof':
        return
      if t[1] == '{':
        depth += 1
      elif t[1] == '}':
        depth -= 1
        if depth == 0:
          if self.peek()[1] == ';':
            self.next()
          return
      elif t[1] == ';' and depth == 0:
        return

  def parse_file(self, fdp):
    while self.peek()[0] != 'eof':
      t = self.peek()
      if t[1] in ('syntax', 'package', 'import', 'option'):
        self.skip_statement()
      elif t[1] == 'message':
        self.next()
        self.parse_message(fdp.message_type.add())
      elif t[1] == 'enum':
        self.next()
        self.parse_enum(fdp.enum_type.add())
      elif t[1] == ';':
        self.next()
      else:
        raise ValueError('%s: unexpected top-level token %r' % (self.fname, t[1]))

  def parse_enum(self, ep):
    ep.name = self.next()[1]
    self.expect('{')
    while self.peek()[1] != '}':
      t = self.next()
      if t[1] in ('option', 'reserved'):
        self.i -= 1
        self.skip_statement()
        continue
      if t[1] == ';':
        continue
      v = ep.value.add()
      v.name = t[1]
      self.expect('=')
      v.number = int(self.next()[1], 0)
      if self.peek()[1] == '[':
        while self.next()[1] != ']':
          pass
      self.expect(';')
    self.expect('}')
    if self.peek()[1] == ';':
      self.next()

  def parse_options(self, field):
    # [default = x, packed = true, ...]
    self.expect('[')
    while True:
      name = self.next()[1]
      self.expect('=')
      t = self.next()
      val = t[1]
      if val == '-' or val == '+':  # sign split from identifier such as -inf
        val = val + self.next()[1]
      if name == 'default':
        field.default_value = _unquote(val) if t[0] == 'str' else val
      elif name == 'packed':
        field.options.packed = (val == 'true')
      elif name == 'deprecated':
        field.options.deprecated = (val == 'true')
      t = self.next()
      if t[1] == ']':
        return
      if t[1] != ',':
        raise ValueError('%s: bad field options near %r' % (self.fname, t[1]))

  def parse_field(self, mp, label, oneof_index=None):
    t = self.next()
    ftype = t[1]
    field = mp.field.add()
    if ftype == 'map':
      self.expect('<')
      ktype = self.next()[1]
      self.expect(',')
      vtype = self.next()[1]
      self.expect('>')
      field.name = self.next()[1]
      self.expect('=')
      field.number = int(self.next()[1], 0)
      entry = mp.nested_type.add()
      entry.name = ''.join(p.capitalize() for p in field.name.split('_')) + 'Entry'
      entry.options.map_entry = True
      for nm, num, ty in (('key', 1, ktype), ('value', 2, vtype)):
        f = entry.field.add()
        f.name, f.number, f.label = nm, num, FD.LABEL_OPTIONAL
        if ty in _SCALARS:
          f.type = _SCALARS[ty]
        else:
          f.type_name = ty
      field.label = FD.LABEL_REPEATED
      field.type_name = entry.name
    else:
      field.label = _LABELS[label]
      if ftype in _SCALARS:
        field.type = _SCALARS[ftype]
      else:
        field.type_name = ftype  # resolved later
      field.name = self.next()[1]
      self.expect('=')
      field.number = int(self.next()[1], 0)
    if oneof_index is not None:
      field.oneof_index = oneof_index
    if self.peek()[1] == '[':
      self.parse_options(field)
    self.expect(';')

  def parse_message(self, mp):
    mp.name = self.next()[1]
    self.expect('{')
    while self.peek()[1] != '}':
      t = self.next()
      v = t[1]
      if v == ';':
        continue
      if v == 'message':
        self.parse_message(mp.nested_type.add())
      elif v == 'enum':
        self.parse_enum(mp.enum_type.add())
      elif v == 'oneof':
        od = mp.oneof_decl.add()
        od.name = self.next()[1]
        idx = len(mp.oneof_decl) - 1
        self.expect('{')
        while self.peek()[1] != '}':
          if self.peek()[1] == 'option':
            self.skip_statement()
            continue
          self.parse_field(mp, 'optional', idx)
        self.expect('}')
      elif v in ('option', 'reserved', 'extensions'):
        self.i -= 1
        self.skip_statement()
      elif v in _LABELS:
        self.parse_field(mp, v)
      elif t[0] == 'id':  # proto3-style field without a label (predict.proto, tf_predict.proto)
        self.i -= 1
        self.parse_field(mp, 'optional')
      else:
        raise ValueError('%s: unexpected token %r in message %s' % (self.fname, v, mp.name))
    self.expect('}')
    if self.peek()[1] == ';':
      self.next()


def _collect(prefix, msgs, enums, table):
  for e in enums:
    table[prefix + e.name] = 'enum'
  for m in msgs:
    table[prefix + m.name] = 'message'
    _collect(prefix + m.name + '.', m.nested_type, m.enum_type, table)


def _resolve(pkg, scope, msgs, table):
  for m in msgs:
    here = scope + [m.name]
    for f in m.field:
      if not f.type_name or f.type_name.startswith('.'):
        continue
      name = f.type_name
      if name.startswith('google.protobuf.'):  # well-known types (layer.proto uses Struct)
        f.type = FD.TYPE_MESSAGE
        f.type_name = '.' + name
        continue
      found = None
      for depth in range(len(here), -1, -1):
        cand = '.'.join(here[:depth] + [name])
        if cand in table:
          found = cand
          break
      if found is None:
        raise ValueError('unresolved type %s in message %s' % (name, '.'.join(here)))
      f.type = FD.TYPE_ENUM if table[found] == 'enum' else FD.TYPE_MESSAGE
      f.type_name = '.' + (pkg + '.' if pkg else '') + found
    _resolve(pkg, here, m.nested_type, table)


class Schema(object):
  """Message classes of a loaded schema: schema.EasyRecConfig(), schema['FeatureConfig']."""

  def __init__(self, pool, package, names):
    self._pool = pool
    self._package = package
    self._names = names
    self._cache = {}

  def __getitem__(self, name):
    if name not in self._cache:
      full = (self._package + '.' if self._package else '') + name
      self._cache[name] = message_factory.GetMessageClass(self._pool.FindMessageTypeByName(full))
    return self._cache[name]

  def __getattr__(self, name):
    if name.startswith('_'):
      raise AttributeError(name)
    try:
      return self[name]
    except KeyError:
      raise AttributeError(name)

  def enum(self, name):
    full = (self._package + '.' if self._package else '') + name
    return self._pool.FindEnumTypeByName(full)

  def has(self, name):
    return name in self._names


def load_schema(paths, package='protos', virtual_name='easyrec_b200_schema.proto'):
  """Parse the given .proto files into one descriptor and return a Schema."""
  fdp = descriptor_pb2.FileDescriptorProto()
  fdp.name = virtual_name
  fdp.package = package
  fdp.syntax = 'proto2'
  for p in paths:
    with open(p, 'rb') as f:
      text = f.read().decode('utf-8', errors='surrogateescape')
    m = re.search(r'^\s*package\s+([\w.]+)\s*;', text, re.M)
    if m and m.group(1) != package:
      continue  # serving protos of another package (predict.proto, tf_predict.proto)
    _Parser(text, os.path.basename(p)).parse_file(fdp)
  # de-duplicate top-level names (the reference defines a few messages twice across files)
  seen = set()
  for lst in (fdp.message_type, fdp.enum_type):
    keep = []
    for m in lst:
      if m.name not in seen:
        seen.add(m.name)
        keep.append(m)
    del lst[len(keep):]
    # rebuild in order (protobuf repeated fields cannot be reassigned directly)
    tmp = [descriptor_pb2.DescriptorProto.FromString(k.SerializeToString())
           if isinstance(k, descriptor_pb2.DescriptorProto)
           else descriptor_pb2.EnumDescriptorProto.FromString(k.SerializeToString()) for k in keep]
    del lst[:]
    lst.extend(tmp)
  table = {}
  _collect('', fdp.message_type, fdp.enum_type, table)
  _resolve(package, [], fdp.message_type, table)
  pool = descriptor_pool.DescriptorPool()
  from google.protobuf import struct_pb2
  wkt = descriptor_pb2.FileDescriptorProto.FromString(struct_pb2.DESCRIPTOR.serialized_pb)
  pool.Add(wkt)
  fdp.dependency.append(wkt.name)
  pool.Add(fdp)
  return Schema(pool, package, set(table))


_DEFAULT = None


def default_schema():
  """EASYREC_PROTO_DIR (an EasyRec checkout's easy_rec/python/protos) when set, else this
  repo's own subset schema."""
  global _DEFAULT
  if _DEFAULT is None:
    d = os.environ.get('EASYREC_PROTO_DIR')
    if d and os.path.isdir(d):
      paths = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith('.proto'))
    else:
      paths = [os.path.join(os.path.dirname(os.path.abspath(__file__)), 'easyrec_subset.proto')]
    _DEFAULT = load_schema(paths)
  return _DEFAULT
