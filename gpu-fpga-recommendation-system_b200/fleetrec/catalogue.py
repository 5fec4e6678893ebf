"""Table catalogues of the three FleetRec models (+ the full 377-table model).

Data, not macros: generated once by tools/gen_catalogue.py from the reference's
FPGA/kernel/user_krnl/embedding_{47,98,377}_krnl/src/hls/constants.hpp and
committed under ../catalogue/*.json.  `segments` restates gather_embeddings()
(embedding_47_krnl.cpp:1097-1217, embedding_98_krnl.cpp:1331-1605,
embedding_377_krnl.cpp:1665-1873): float [dst, dst+len) of an item's concat
vector is table[idx[table]][col, col+len).
"""
import copy
import json
import os
from dataclasses import dataclass, field
from typing import List

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "catalogue")
TIERS = {"HBM": 0, "DDR": 1, "PLRAM": 2, "CPU": 3}
MODEL_NAMES = ("small", "medium", "large_half", "large")

# embedding_47_krnl.cpp:903-904 (load_access_idx)
IDX_RANDOM = (3, 99, 38, 72, 29, 57, 1, 72, 36, 76, 35, 50, 37, 57, 13, 66,
              26, 70, 41, 93, 48, 82, 44, 78, 25, 52, 3, 92, 36, 56, 46, 88)


@dataclass
class Table:
    id: int
    tier: str
    tier_index: int
    bank: int
    round: int
    rows: int
    dim: int
    axi_padded: int
    addr_axi: int
    half: str = ""


@dataclass
class Segment:
    dst: int
    table: int
    col: int
    len: int
    pad: bool = False


@dataclass
class Model:
    name: str
    tables: List[Table]
    segments: List[Segment]
    concat_floats: int
    data_floats: int
    hidden: List[int]
    fpga_batch: int = 32
    source: str = ""
    concat_spec: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def n_tables(self):
        return len(self.tables)

    @property
    def layer_dims(self):
        """[in, h1, h2, h3, 1] (cuda_server.c constant.h:21-27)."""
        return [self.concat_floats] + list(self.hidden)

    def table_bytes(self):
        return sum(t.rows * t.dim * 4 for t in self.tables)

    def gather_bytes_per_item(self, materialised=False):
        """sum dim_t*4 + T*4 (+ D_pad*4 when the concat is written)."""
        b = sum(t.dim for t in self.tables) * 4 + self.n_tables * 4
        return b + (self.concat_floats * 4 if materialised else 0)

    def mlp_flops_per_item(self):
        d = self.layer_dims
        return 2 * sum(d[k] * d[k + 1] for k in range(4))

    def with_row_cap(self, cap):
        """Same dims and concat order, every table truncated to <= cap rows (tests)."""
        m = copy.deepcopy(self)
        for t in m.tables:
            t.rows = min(t.rows, cap)
        m.extra["row_cap"] = cap
        return m


def load(name) -> Model:
    if name not in MODEL_NAMES:
        raise ValueError(f"unknown model {name!r}; one of {MODEL_NAMES}")
    with open(os.path.join(_DIR, name + ".json")) as f:
        j = json.load(f)
    tables = [Table(**t) for t in j["tables"]]
    segs = [Segment(**s) for s in j["segments"]]
    return Model(name=j["name"], tables=tables, segments=segs, concat_floats=j["concat_floats"],
                 data_floats=j["data_floats"], hidden=j["hidden"], fpga_batch=j["fpga_batch"],
                 source=j["source"], concat_spec=j["concat_spec"])


def synthetic(n_tables, rows, dim, hidden=(1024, 512, 256, 1), name="stress") -> Model:
    """BASELINE.json configs[4]: n_tables x rows x dim, concat in table order."""
    tables = [Table(id=i, tier="HBM", tier_index=i, bank=i % 28, round=i // 28, rows=rows, dim=dim,
                    axi_padded=dim // 4, addr_axi=0) for i in range(n_tables)]
    segs = [Segment(dst=i * dim, table=i, col=0, len=dim) for i in range(n_tables)]
    return Model(name=name, tables=tables, segments=segs, concat_floats=n_tables * dim,
                 data_floats=n_tables * dim, hidden=list(hidden))
