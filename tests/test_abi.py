"""CPU: the C-ABI library loads and exports exactly what include/er_b200.h declares."""
import ctypes
import os
import re

from easyrec_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, 'include', 'er_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(er_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_all_exported_and_bound():
  names = _declared()
  assert len(names) >= 15
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for n in names:
    assert hasattr(lib, n), 'missing export: ' + n
  assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_struct_layout():
  lib = _lib.load()
  assert lib.er_abi_version() == _lib.ABI_VERSION == 3
  assert _lib.SLOT_DTYPE.itemsize == 48
  assert ctypes.sizeof(_lib.ErOpt) == 40 and _lib.ErOpt.hyper_dev.offset == 32


def test_invalid_arguments_fail_loudly_without_gpu():
  lib = _lib.load()
  # null pointers are rejected on the host before any CUDA call
  st = lib.er_fm_fwd(None, 4, 2, 4, 8, None, None)
  assert st == 1 and b'er_fm_fwd' in lib.er_last_error()


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, 'easyrec_b200')
  for dp, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
        text = open(os.path.join(dp, f)).read()
        assert 'oracle' not in text.lower() or f == '_lib.py' or 'never' in text.lower(), (dp, f)
