"""Cartesian-merge planner (MicroRec's data-structure side of FleetRec).

The reference ships its tables already merged (the "MicroRec-style Cartesian-merged tables" of the
north star) and carries no planner or builder, only the warning that the merged row count may
overflow `int` (FPGA/host/embedding_47_krnl/host.cpp:379-382).  This module supplies the missing
piece on the host side:

  plan_merges()   which table pairs to merge under an HBM byte budget
  apply_merges()  the merged catalogue (tables, concat segments) + the index remap
  build()         an Engine whose merged tables are built ON THE DEVICE from their sources
                  (fr_merge_tables), ready to be driven with remapped indices

Definition: M = A x B has rowsA*rowsB rows of dimA+dimB floats,
M[iA*rowsB + iB] = A[iA] || B[iB].  One lookup in M replaces one lookup in A and one in B; the
concat vector is unchanged bit for bit, which is what the parity tests assert.
"""
import copy
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .catalogue import Model, Segment, Table

INT32_MAX = 2 ** 31 - 1


@dataclass
class MergePlan:
    pairs: List[Tuple[int, int]]     # (table id A, table id B) in the ORIGINAL model
    extra_bytes: int                 # merged bytes minus the bytes of the sources they replace
    lookups_saved: int               # random accesses per item removed (= len(pairs))


def merged_cost(a: Table, b: Table) -> int:
    return a.rows * b.rows * (a.dim + b.dim) * 4


def plan_merges(model: Model, budget_bytes: int, max_rows: int = INT32_MAX,
                candidates: Sequence[int] = None) -> MergePlan:
    """Greedy, deterministic: every merge removes exactly one random access per item, so pairs are
    taken in order of increasing merged size (cheapest access removed first) among tables not yet
    used, as long as the extra bytes fit the budget and the merged row count stays addressable by
    the int32 index the lookup consumes (embedding_47_krnl.cpp:899-914 streams `int` indices).
    `candidates` restricts the tables considered (default: all)."""
    ids = list(range(model.n_tables)) if candidates is None else list(candidates)
    tabs = sorted((model.tables[i] for i in ids), key=lambda t: (t.rows * t.dim, t.rows, t.id))
    used, pairs, extra = set(), [], 0
    # smallest-with-next-smallest is optimal for a product cost under "one pair per table"
    i = 0
    while i + 1 < len(tabs):
        a, b = tabs[i], tabs[i + 1]
        if a.id in used or b.id in used:
            i += 1
            continue
        cost = merged_cost(a, b) - (a.rows * a.dim + b.rows * b.dim) * 4
        if a.rows * b.rows > max_rows or extra + cost > budget_bytes:
            break   # candidates are sorted: every later pair is at least as large
        pairs.append((a.id, b.id))
        used.update((a.id, b.id))
        extra += cost
        i += 2
    return MergePlan(pairs=pairs, extra_bytes=extra, lookups_saved=len(pairs))


@dataclass
class MergedModel:
    model: Model                      # catalogue the engine is created with
    pairs: List[Tuple[int, int]]
    kept: List[int]                   # original ids of unmerged tables, in new-id order
    merged_ids: List[int]             # new id of the merged table of pairs[k]
    source_ids: List[Tuple[int, int]]  # new ids of the (A, B) source tables kept for the device-side build
    rows: List[Tuple[int, int]]       # (rowsA, rowsB) of pairs[k]

    def remap(self, idx: np.ndarray) -> np.ndarray:
        """idx [B][T_original] -> [B][T_new]; merged columns carry iA*rowsB + iB (computed in int64,
        range-checked, host.cpp:379-382), source-table columns are unused and left 0."""
        idx = np.asarray(idx)
        out = np.zeros((idx.shape[0], self.model.n_tables), np.int32)
        for new, old in enumerate(self.kept):
            out[:, new] = idx[:, old]
        for k, (a, b) in enumerate(self.pairs):
            m = idx[:, a].astype(np.int64) * self.rows[k][1] + idx[:, b].astype(np.int64)
            if m.size and m.max() > INT32_MAX:
                raise OverflowError(f"merged index {int(m.max())} of pair {(a, b)} exceeds int32")
            out[:, self.merged_ids[k]] = m.astype(np.int32)
        return out


def apply_merges(model: Model, pairs: Sequence[Tuple[int, int]]) -> MergedModel:
    """New catalogue: [unmerged tables in original order] + [one merged table per pair] + [the
    sources of every pair] (the sources are only there so the device can build the merged image;
    no concat segment reads them).  Concat segments keep their dst, so the wire order is unchanged."""
    merged_of = {}
    for k, (a, b) in enumerate(pairs):
        if a in merged_of or b in merged_of or a == b:
            raise ValueError(f"table used twice in merge pairs: {(a, b)}")
        merged_of[a] = (k, 0)
        merged_of[b] = (k, model.tables[a].dim)
    kept = [t.id for t in model.tables if t.id not in merged_of]
    new_id = {old: new for new, old in enumerate(kept)}
    tables = []
    for old in kept:
        t = copy.copy(model.tables[old])
        t.id = new_id[old]
        tables.append(t)
    merged_ids, rows = [], []
    for a, b in pairs:
        ta, tb = model.tables[a], model.tables[b]
        if ta.rows * tb.rows > INT32_MAX:
            raise OverflowError(f"merged table of {(a, b)} has {ta.rows * tb.rows} rows (> int32)")
        merged_ids.append(len(tables))
        rows.append((ta.rows, tb.rows))
        tables.append(Table(id=len(tables), tier=ta.tier, tier_index=ta.tier_index, bank=ta.bank, round=ta.round,
                            rows=ta.rows * tb.rows, dim=ta.dim + tb.dim, axi_padded=(ta.dim + tb.dim) // 4, addr_axi=0))
    source_ids = []
    for a, b in pairs:
        pair_ids = []
        for old in (a, b):
            t = copy.copy(model.tables[old])
            t.id = len(tables)
            pair_ids.append(t.id)
            tables.append(t)
        source_ids.append(tuple(pair_ids))
    segs = []
    for s in model.segments:
        if s.table in merged_of:
            k, off = merged_of[s.table]
            segs.append(Segment(dst=s.dst, table=merged_ids[k], col=s.col + off, len=s.len, pad=s.pad))
        else:
            segs.append(Segment(dst=s.dst, table=new_id[s.table], col=s.col, len=s.len, pad=s.pad))
    m = Model(name=model.name + "+merged", tables=tables, segments=segs, concat_floats=model.concat_floats,
              data_floats=model.data_floats, hidden=list(model.hidden), fpga_batch=model.fpga_batch,
              source=model.source, concat_spec=model.concat_spec, extra=dict(model.extra))
    return MergedModel(model=m, pairs=list(pairs), kept=kept, merged_ids=merged_ids, source_ids=source_ids, rows=rows)


def lookups_per_item(mm: MergedModel) -> int:
    """Distinct tables an item's concat vector reads (random accesses per item)."""
    return len({s.table for s in mm.model.segments})


def build(mm: MergedModel, original_tables, **engine_kw):
    """Engine for the merged catalogue: unmerged tables and merge sources are uploaded, merged images
    are built on the device (fr_merge_tables).  original_tables[t] is the [rows][dim] image of the
    ORIGINAL table t."""
    from .engine import Engine
    eng = Engine(mm.model, **engine_kw)
    for new, old in enumerate(mm.kept):
        eng.load_table(new, original_tables[old])
    for k, (a, b) in enumerate(mm.pairs):
        sa, sb = mm.source_ids[k]
        eng.load_table(sa, original_tables[a])
        eng.load_table(sb, original_tables[b])
        eng.merge_tables(sa, sb, mm.merged_ids[k])
    return eng
