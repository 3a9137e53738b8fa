"""InputLayer: feature groups -> dense tensors, on the fused sm_100a lookup path.

Mirrors the reference surface
  InputLayer(feature_configs, feature_groups, ..., wide_output_dim)      layers/input_layer.py:33-69
  input_layer(features, group_name) -> (concat [B, sum D], [per-feature])  layers/input_layer.py:245-278
with `FeatureColumnParser` (feature_column/feature_column.py:44-203, 259-656) collapsed into a
static *table plan*: which table each feature reads (shared `embedding_name` groups), its bucket
rule, combiner and output column -- fixed at construction, uploaded once as er_slot_t records.

Input contract (the reference's packed form, input/parquet_input.py:201-239):
  features['sparse_fea'] = ids int64 [n_sparse*B] feature-major (single-valued), or
                           (ids int64 [L], lens int32 [n_sparse*B]) for multi-valued features
  features['dense_fea']  = float32 [B, sum raw_input_dim] in raw-feature config order
Outputs keep feature_group CONFIG order (compat/feature_column/feature_column.py:388-414).
"""
import collections

import numpy as np
import torch

from easyrec_b200 import _lib
from easyrec_b200 import embedding as E
from easyrec_b200 import kernels as K

FeatureSpec = collections.namedtuple(
    'FeatureSpec',
    ['name', 'kind',            # 'id' | 'raw' | 'tag' | 'seq'
     'embedding_dim', 'bucket_mode', 'num_buckets', 'combiner', 'embedding_name',
     'min_val', 'max_val', 'raw_input_dim', 'seq_len'])


def id_feature(name, embedding_dim, hash_bucket_size=0, num_buckets=0, combiner='sum',
               embedding_name='', packed_mod=False):
  """IdFeature: hash_bucket_size -> Fingerprint64(as_string) % size; num_buckets -> identity
  (feature_column/feature_column.py:259-300).  packed_mod: the Parquet packed rule
  `vals % num_buckets` (input/parquet_input.py:221)."""
  if hash_bucket_size > 0:
    mode, nb = _lib.BUCKET_FARM_DECIMAL, hash_bucket_size
  elif packed_mod:
    mode, nb = _lib.BUCKET_MOD, num_buckets
  else:
    mode, nb = _lib.BUCKET_IDENTITY, num_buckets
  return FeatureSpec(name, 'id', embedding_dim, mode, nb, combiner, embedding_name, 0., 0., 1, 1)


def raw_feature(name, embedding_dim=0, min_val=0.0, max_val=0.0, raw_input_dim=1):
  """RawFeature: (x-min)/(max-min) when max>min (input/input.py:638-640); with embedding_dim>0
  it becomes ids 0..k-1 weighted by the values (input/input.py:648-673)."""
  return FeatureSpec(name, 'raw', embedding_dim, _lib.BUCKET_NONE, raw_input_dim, 'sum', '',
                     float(min_val), float(max_val), raw_input_dim, 1)


_COMBINER = {'sum': _lib.COMBINER_SUM, 'mean': _lib.COMBINER_MEAN, 'sqrtn': _lib.COMBINER_SQRTN}


class InputLayer(object):
  """Builds arenas + fused calls for a set of feature groups and evaluates them.

  groups: OrderedDict group_name -> dict(features=[names...], wide=bool)
  A group marked wide uses `wide_output_dim` columns per feature with combiner sum
  (feature_column/feature_column.py:616-622)."""

  def __init__(self, features, groups, batch_size, device, wide_output_dim=1,
               embedding_optimizer=_lib.OPT_ADAGRAD, shard_n=1, shard_rank=0, generator=None,
               adagrad_init=0.1):
    self.features = collections.OrderedDict((f.name, f) for f in features)
    self.groups = groups
    self.batch_size = batch_size
    self.device = device
    self.wide_output_dim = wide_output_dim
    self.sparse_names = [f.name for f in features if f.kind != 'raw']
    self.raw_names = [f.name for f in features if f.kind == 'raw']
    self.raw_cols = {}
    c = 0
    for n in self.raw_names:
      self.raw_cols[n] = (c, c + self.features[n].raw_input_dim)
      c += self.features[n].raw_input_dim
    self.n_dense = c
    # ---- table plan: one arena per embedding dim --------------------------------------
    self.arenas = collections.OrderedDict()
    plan = collections.OrderedDict()  # dim -> list of (group, feature, Slot)
    self.group_layout = {}            # group -> list of (feature, kind, dim, arena_dim, buf, col)
    self.group_bufs = collections.OrderedDict()   # (dim, group) -> buf index within arena call
    for gname, g in groups.items():
      layout = []
      for fname in g['features']:
        f = self.features[fname]
        wide = bool(g.get('wide'))
        dim = wide_output_dim if wide else f.embedding_dim
        if f.kind == 'raw' and dim == 0:
          layout.append((fname, 'dense', f.raw_input_dim, None, None, None))
          continue
        table = (f.embedding_name or fname + '_embedding') + ('_wide' if wide else '')
        arena = self.arenas.setdefault(dim, E.Arena(dim, device, shard_n, shard_rank))
        arena.add_table(table, f.num_buckets)
        key = (dim, gname)
        if key not in self.group_bufs:
          self.group_bufs[key] = sum(1 for k in self.group_bufs if k[0] == dim)
        comb = _lib.COMBINER_SUM if (wide or f.kind == 'raw') else _COMBINER[f.combiner]
        slot = E.Slot(gname + '/' + fname, table, f.bucket_mode, f.num_buckets, comb,
                      out_buf=self.group_bufs[key])
        plan.setdefault(dim, []).append((gname, fname, slot))
        layout.append((fname, 'emb', dim, dim, self.group_bufs[key], None))
      self.group_layout[gname] = layout
    for a in self.arenas.values():
      a.materialize(embedding_optimizer, generator=generator, adagrad_init=adagrad_init)
    # ---- fused calls (all slots single-valued here; CSR inputs go through lookup_csr) ----
    self.calls = collections.OrderedDict()
    self.call_feature_idx = {}
    self.static_ids = {}
    self.static_w = {}
    B = batch_size
    for dim, items in plan.items():
      n_bufs = sum(1 for k in self.group_bufs if k[0] == dim)
      widths = [0] * n_bufs
      for gname, fname, slot in items:
        widths[slot.out_buf] += dim
      call = E.ArenaCall(self.arenas[dim], [s for _, _, s in items], B, widths, single_valued=True)
      self.calls[dim] = call
      for i, (gname, fname, slot) in enumerate(items):
        lay = self.group_layout[gname]
        for j, e in enumerate(lay):
          if e[0] == fname and e[1] == 'emb' and e[5] is None:
            lay[j] = e[:5] + (call.slot_cols[i],)
            break
      # where each slot's ids / weights come from
      sparse_idx = {n: i for i, n in enumerate(self.sparse_names)}
      src = []
      for gname, fname, slot in items:
        f = self.features[fname]
        src.append(('raw', self.raw_cols[fname][0]) if f.kind == 'raw' else ('id', sparse_idx[fname]))
      self.call_feature_idx[dim] = src
      self.static_ids[dim] = torch.zeros(call.n_seg, dtype=torch.int64, device=device)
      has_raw = any(k == 'raw' for k, _ in src)
      self.static_w[dim] = (torch.ones(call.n_seg, dtype=torch.float32, device=device)
                            if has_raw else None)
      identity = [k for k, _ in src] == ['id'] * len(src) and [i for _, i in src] == list(
          range(len(self.sparse_names)))
      call.identity_ids = identity
    mn = [self.features[n].min_val for n in self.raw_names for _ in range(self.features[n].raw_input_dim)]
    mx = [self.features[n].max_val for n in self.raw_names for _ in range(self.features[n].raw_input_dim)]
    self.raw_min = torch.tensor(mn, dtype=torch.float32, device=device)
    rng = np.array(mx, np.float32) - np.array(mn, np.float32)
    self.raw_has_range = bool((rng > 0).any())
    self.raw_range = torch.tensor(np.where(rng > 0, rng, 1.0), dtype=torch.float32, device=device)
    self.raw_sub = torch.tensor(np.where(rng > 0, np.array(mn, np.float32), 0.0),
                                dtype=torch.float32, device=device)
    self.opt_holder = {'opt': K.make_opt(embedding_optimizer, 0.01)}
    self._pending = []
    self._rows_cache = {}
    self._gather_plan = {}

  # ------------------------------------------------------------------
  def set_optimizer_step(self, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """Per-step hyper-parameters of the fused row update (host-side schedule,
    core/learning_schedules.py:30-75; beta powers as compat/adam_s.py:233-245)."""
    kind = next(iter(self.arenas.values())).opt_kind
    self.opt_holder['opt'] = K.make_opt(kind, lr, beta1, beta2, eps, beta1**(step + 1),
                                        beta2**(step + 1), grad_scale)

  def backward_update(self):
    """After loss.backward(): K7 for every arena looked up since the last call (dedup, segment
    sum and the fused optimizer row update).  The reference's counterpart is
    opt.apply_gradients on the tables' IndexedSlices (compat/optimizers.py:413-416)."""
    for call, rows, w, outs in self._pending:
      E.fused_backward_update(call, rows, outs, self.opt_holder['opt'], weights=w)
    self._pending = []

  def normalize_dense(self, dense):
    if not self.raw_has_range:
      return dense
    return (dense - self.raw_sub) / self.raw_range  # (x - min) / (max - min), input/input.py:638-640

  def _gather_inputs(self, dim, ids, dense_norm):
    """ids int64 [n_sparse*B] feature-major -> (ids, weights) in this call's slot order.

    Two strided copies per arena (ids of the id slots, normalised values of the raw slots);
    slots that are already in packed order are used in place."""
    call = self.calls[dim]
    if call.identity_ids:
      return ids, None
    B = self.batch_size
    S = call.n_slots
    out_ids = self.static_ids[dim]
    out_w = self.static_w[dim]
    plan = self._gather_plan.get(dim)
    if plan is None:
      src = self.call_feature_idx[dim]
      id_pos = [i for i, (k, _) in enumerate(src) if k == 'id']
      id_src = [j for k, j in src if k == 'id']
      raw_pos = [i for i, (k, _) in enumerate(src) if k == 'raw']
      raw_src = [j for k, j in src if k == 'raw']

      def as_slice(v):
        return (v[0], v[-1] + 1) if v and v == list(range(v[0], v[0] + len(v))) else None

      plan = dict(id_pos=as_slice(id_pos), id_src=as_slice(id_src), raw_pos=as_slice(raw_pos),
                  raw_src=as_slice(raw_src),
                  id_pos_t=torch.tensor(id_pos, dtype=torch.int64, device=self.device),
                  id_src_t=torch.tensor(id_src, dtype=torch.int64, device=self.device),
                  raw_pos_t=torch.tensor(raw_pos, dtype=torch.int64, device=self.device),
                  raw_src_t=torch.tensor(raw_src, dtype=torch.int64, device=self.device))
      self._gather_plan[dim] = plan
    ids2 = ids.view(len(self.sparse_names), B)
    o2 = out_ids.view(S, B)
    if plan['id_pos'] and plan['id_src']:
      o2[plan['id_pos'][0]:plan['id_pos'][1]].copy_(ids2[plan['id_src'][0]:plan['id_src'][1]])
    elif plan['id_pos_t'].numel():
      o2.index_copy_(0, plan['id_pos_t'], ids2.index_select(0, plan['id_src_t']))
    if out_w is not None and plan['raw_pos_t'].numel():
      w2 = out_w.view(S, B)
      dt = dense_norm.t()
      if plan['raw_pos'] and plan['raw_src']:
        w2[plan['raw_pos'][0]:plan['raw_pos'][1]].copy_(dt[plan['raw_src'][0]:plan['raw_src'][1]])
      else:
        w2.index_copy_(0, plan['raw_pos_t'], dt.index_select(0, plan['raw_src_t']))
    return out_ids, out_w

  def lookup(self, features):
    """Runs K1 + K2 for every arena; returns {group: (concat, [per-feature views])}."""
    ids = features['sparse_fea']
    dense = features.get('dense_fea')
    dense_norm = self.normalize_dense(dense) if dense is not None else None
    results = {}
    self._rows_cache = {}
    self._pending = []
    for dim, call in self.calls.items():
      # arenas whose slots read the same features with the same bucket rules and row offsets
      # (e.g. DeepFM's wide and deep groups) share one gather + one K1 launch
      key = (call.slots_np[['num_buckets', 'row_offset', 'seg_begin', 'bucket_mode']].tobytes(),
             tuple(self.call_feature_idx[dim]))
      hit = self._rows_cache.get(key)
      if hit is None:
        cids, w = self._gather_inputs(dim, ids, dense_norm)
        rows = K.bucketize(cids, call.slots_dev, call.n_slots, call.n_seg)
        self._rows_cache[key] = (rows, w)
      else:
        rows, w = hit
      outs = E.fused_lookup(call, rows, weights=w)
      results[dim] = outs
      self._pending.append((call, rows, w, outs))
    out = {}
    for gname, layout in self.group_layout.items():
      mats = {}
      per_feature = []
      pieces = []
      for (fname, kind, width, adim, buf, col) in layout:
        if kind == 'dense':
          c0, c1 = self.raw_cols[fname]
          v = dense_norm[:, c0:c1]
        else:
          m = results[adim][buf]
          mats[(adim, buf)] = m
          v = m[:, col:col + width]
        per_feature.append(v)
        pieces.append((kind, adim, buf))
      if len(mats) == 1 and all(k == 'emb' for k, _, _ in pieces):
        (adim, buf), m = next(iter(mats.items()))
        w = self.calls[adim].out_widths[buf]
        concat = m if m.shape[1] == w else m[:, :w]
      else:
        concat = torch.cat(per_feature, dim=1)
      out[gname] = (concat, per_feature)
    return out
