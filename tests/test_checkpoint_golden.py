"""CPU: sharded-table checkpoints against data produced by the reference saver's own python
closures (`_save_embed` / `_load_embed`, compat/embedding_parallel_saver.py:99-173; the C++ LoadEmbed op
ops/src/load_dense_embed.cc:52-136 has the same rule) - tests/golden/make_checkpoint_golden.py.

  * python save: same file names, byte-identical part files, stale parts of a larger job removed by worker 0;
  * native restore (er_load_embed through the C ABI): every (old workers -> new workers, rank) re-shard is
    bit-identical to the reference's result;
  * an Arena with its Adagrad slot survives save on 2 workers -> restore on 3 workers / 1 worker."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from easyrec_b200 import _lib, checkpoint, embedding as E

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'reference_checkpoint.json')))
_spec = importlib.util.spec_from_file_location('make_checkpoint_golden', os.path.join(HERE, 'golden', 'make_checkpoint_golden.py'))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


def test_save_and_reshard_match_the_reference_saver(tmp_path):
  full = gen.table()
  V, D, var = GOLD['rows'], GOLD['dim'], GOLD['var_name']
  ckpt = str(tmp_path / 'model.ckpt-100')
  for save in GOLD['saves']:            # 3 workers, then 2 workers into the same directory
    world = save['world']
    for rank in range(world):
      checkpoint.save_embed(gen.shard(full, rank, world), ckpt, var, rank, world)
    files = sorted(os.listdir(ckpt + '-embedding'))
    assert ['model.ckpt-100-embedding/' + f for f in files] == save['files']
    assert [hashlib.sha256(open(os.path.join(ckpt + '-embedding', f), 'rb').read()).hexdigest() for f in files] == save['sha256']
    for load in save['loads']:
      part_size = (V + load['world'] - 1) // load['world']
      got = checkpoint.load_embed(ckpt, var, D, part_size, load['rank'], load['world'])
      assert np.array_equal(got, np.array(load['vals'], np.float32)), (world, load['world'], load['rank'])
      # and it is the shard of the original table (the padding row of an odd split reads as zeros)
      assert np.array_equal(got, gen.shard(full, load['rank'], load['world']))


def test_load_embed_reports_errors_like_the_op(tmp_path):
  ckpt = str(tmp_path / 'model.ckpt-1')
  with pytest.raises(_lib.ErError, match='cannot open'):
    checkpoint.load_embed(ckpt, 'v:0', 4, 10, 0, 1)
  checkpoint.save_embed(np.ones((5, 4), np.float32), ckpt, 'v:0', 0, 1)
  with pytest.raises(_lib.ErError, match='no part files'):
    checkpoint.load_embed(ckpt, 'w:0', 4, 5, 0, 1)
  with pytest.raises(_lib.ErError, match='should be equal to embed_part_size'):
    checkpoint.load_embed(ckpt, 'v:0', 4, 9, 0, 1)     # the files hold 5 rows, the variable wants 9
  assert np.array_equal(checkpoint.load_embed(ckpt, 'v:0', 4, 5, 0, 1), np.ones((5, 4), np.float32))


def _arena(world, rank, tables):
  a = E.Arena(4, 'cpu', shard_n=world, shard_rank=rank)
  for name, v in tables:
    a.add_table(name, v)
  a.materialize(_lib.OPT_ADAGRAD, generator=torch.Generator().manual_seed(rank))
  return a


def test_arena_with_optimizer_slot_survives_a_change_of_worker_count(tmp_path):
  tables = [('user_id', 41), ('item_id', 12)]
  full_w = {n: torch.randn(v, 4, generator=torch.Generator().manual_seed(v)) for n, v in tables}
  full_acc = {n: torch.rand(v, 4, generator=torch.Generator().manual_seed(v + 1)) + 0.1 for n, v in tables}
  ckpt = str(tmp_path / 'model.ckpt-7')
  for rank in range(2):
    a = _arena(2, rank, tables)
    for n, v in tables:
      off, local, _ = a.tables[n]
      for dst, src in ((a.weight, full_w[n]), (a.state0, full_acc[n])):
        dst[off:off + local].zero_()
        rows = src[rank::2]
        dst[off:off + rows.shape[0]].copy_(rows)
    written = checkpoint.save_arena(a, ckpt)
    assert len(written) == 4 and all('-part-%d.bin' % rank in w for w in written)
  assert 'embed-input_layer__user_id__embedding_weights__Adagrad:0-part-1.bin' in os.listdir(ckpt + '-embedding')
  for world in (3, 1):
    for rank in range(world):
      b = _arena(world, rank, tables)
      checkpoint.restore_arena(b, ckpt)
      for n, v in tables:
        off, local, _ = b.tables[n]
        for got, src in ((b.weight, full_w[n]), (b.state0, full_acc[n])):
          rows = src[rank::world]
          assert torch.equal(got[off:off + rows.shape[0]], rows)
          assert not got[off + rows.shape[0]:off + local].any()


def test_golden_file_matches_its_generator_when_the_reference_is_mounted(tmp_path):
  if not os.path.isdir('/root/reference/easy_rec/python'):
    pytest.skip('reference checkout not mounted')
  gen.OUT = str(tmp_path / 'out.json')
  gen.main()
  assert json.load(open(gen.OUT)) == GOLD
