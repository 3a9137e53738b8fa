"""Re-check tests/golden/reference_kats.json against the reference's test source.

The reference cannot be imported here (TensorFlow is absent), so the fixture was
transcribed by hand.  When /root/reference is present this script verifies that the
transcribed constants still appear in easy_rec/python/test/embed_test.py (table
constants, inputs with their \\x03/\\x04 separators, asserted outputs).
Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

REF = '/root/reference/easy_rec/python/test/embed_test.py'
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
  kats = json.load(open(os.path.join(HERE, 'reference_kats.json')))
  if not os.path.exists(REF):
    print('reference not mounted; nothing to verify')
    return 0
  src = open(REF, 'rb').read()
  need = [
      b'consts: [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]', b"'0.1,0.2,0.3,0.4,0.5'",
      b'fea_val[0][0] - 9.5', b'fea_val[0][1] - 11.0',
      b"'0\x041\x031\x042', '1\x043\x032\x044\x043\x030'", b'fea_val[0][0][0] - 2',
      b'fea_val[0][0][1] - 3', b'fea_val[0][1][0] - 4', b'fea_val[0][1][1] - 5',
      b'num_buckets: 5', b"combiner: 'mean'", b'raw_input_dim: 5'
  ]
  missing = [n for n in need if n not in src]
  if missing:
    print('MISSING in reference test:', missing)
    return 1
  # the transcription of the sequence sample into (ids, lens)
  samples = ['0\x041\x031\x042', '1\x043\x032\x044\x043\x030']
  ids, lens, T = [], [], 3
  for s in samples:
    pos = s.split('\x03')
    for t in range(T):
      toks = pos[t].split('\x04') if t < len(pos) else []
      ids += [int(x) for x in toks]
      lens.append(len(toks))
  assert ids == kats['embed_test_seq_multi']['ids'], ids
  assert lens == kats['embed_test_seq_multi']['lens'], lens
  print('golden fixture matches the reference test source')
  return 0


if __name__ == '__main__':
  sys.exit(main())
