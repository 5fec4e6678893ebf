#!/usr/bin/env python3
"""Regenerate the FleetRec table catalogues from the reference's constants.hpp.

Runs only in the build container (needs /root/reference).  Its outputs are
committed, so nothing at test/bench time reads the reference tree:

  gpu-fpga-recommendation-system_b200/catalogue/{small,medium,large_half,large}.json
  gpu-fpga-recommendation-system_b200/csrc/fr_catalogue_data.inc   (same data for the C library)

What is parsed (numbers only, `^#define NAME VALUE`):
  FPGA/kernel/user_krnl/embedding_{47,98,377}_krnl/src/hls/constants.hpp
    TABLE_SIZE_<tier>_<i>, DATA_SIZE_<tier>_<i>, AXI_PADDED_SIZE_<tier>_<i>,
    ADDR_AXI_<tier>_<i>, <tier>_BANK_NUM, VECTOR_SIZE_<tier>_BANK_<b>, INPUT_SIZE
What is hand-stated here (and then machine-checked against the reference by
oracle/ref_harness, which executes the reference's own gather_embeddings()):
  the per-model emit order of bank vectors on the wire, see CONCAT_SPEC below
  (embedding_47_krnl.cpp:1140-1216, embedding_98_krnl.cpp:1487-1604,
   embedding_377_krnl.cpp:1633-1662,1780-1872).

Catalogue invariants asserted:
  (2) DATA_SIZE == PADDED_SIZE == 4*AXI_PADDED_SIZE for every table
  (3) table t lives in bank t mod NBANKS, ADDR_AXI_t == sum(rows*axi) of lower rounds
  (4) VECTOR_SIZE_<tier>_BANK_b == sum of dims of that bank's tables
"""
import json
import os
import re
import sys

REF = "/root/reference/FPGA/kernel/user_krnl"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-fpga-recommendation-system_b200")

MODELS = {
    "small": ("embedding_47_krnl", [1024, 512, 256, 1]),
    "medium": ("embedding_98_krnl", [1024, 512, 256, 1]),
    "large_half": ("embedding_377_krnl", [2048, 512, 256, 1]),
}

# Emit order of bank vectors per item.  Token = <tier letter><bank>[-<bank>]
# or "dup:<tier letter><bank>:<n>" = the first n floats of that bank vector
# again (the medium model's 876->880 pad, embedding_98_krnl.cpp:1078,1099).
CONCAT_SPEC = {
    "small": "P0-16 H0-27 D0-1",
    "medium": "P0-18 H27 dup:P16:4 H0-26 D0-1",
    "large_half": "P8-10 H0-7 D0-1 P0-7 H8-27",
}
TIER_OF = {"H": "HBM", "D": "DDR", "P": "PLRAM"}
TIER_ORDER = ["HBM", "DDR", "PLRAM"]   # global table id order used by idx[B][T]


def parse_defines(path):
    d = {}
    with open(path) as f:
        for line in f:
            m = re.match(r"^#define\s+(\w+)\s+(\d+)\b", line)
            if m:
                d[m.group(1)] = int(m.group(2))
    return d


def build(model):
    krnl, hidden = MODELS[model]
    d = parse_defines(f"{REF}/{krnl}/src/hls/constants.hpp")
    nbanks = {"HBM": d["HBM_BANK_NUM"], "DDR": d["DDR_BANK"], "PLRAM": d["PLRAM_BANK_NUM"]}
    tables = []
    first_of_tier = {}
    for tier in TIER_ORDER:
        n = d[f"TABLE_NUM_{tier}"]
        first_of_tier[tier] = len(tables)
        bank_fill = [0] * nbanks[tier]
        for i in range(n):
            dim = d[f"DATA_SIZE_{tier}_{i}"]
            axi = d[f"AXI_PADDED_SIZE_{tier}_{i}"]
            rows = d[f"TABLE_SIZE_{tier}_{i}"]
            assert dim == d[f"PADDED_SIZE_{tier}_{i}"] == 4 * axi, (model, tier, i)
            bank, rnd = i % nbanks[tier], i // nbanks[tier]
            addr = d[f"ADDR_AXI_{tier}_{i}"]
            assert addr == bank_fill[bank], (model, tier, i, addr, bank_fill[bank])
            bank_fill[bank] += rows * axi
            tables.append(dict(id=len(tables), tier=tier, tier_index=i, bank=bank, round=rnd,
                               rows=rows, dim=dim, axi_padded=axi, addr_axi=addr))
        for b in range(nbanks[tier]):
            vs = sum(t["dim"] for t in tables[first_of_tier[tier]:] if t["bank"] == b)
            assert vs == d[f"VECTOR_SIZE_{tier}_BANK_{b}"], (model, tier, b)

    def bank_tables(letter, b):
        tier = TIER_OF[letter]
        return [t for t in tables if t["tier"] == tier and t["bank"] == b]

    segs, off = [], 0
    for tok in CONCAT_SPEC[model].split():
        if tok.startswith("dup:"):
            _, bk, n = tok.split(":")
            n = int(n)
            for t in bank_tables(bk[0], int(bk[1:])):
                take = min(n, t["dim"])
                if take:
                    segs.append(dict(dst=off, table=t["id"], col=0, len=take, pad=True))
                    off += take
                    n -= take
            assert n == 0
            continue
        m = re.match(r"([HDP])(\d+)(?:-(\d+))?$", tok)
        lo = int(m.group(2))
        hi = int(m.group(3)) if m.group(3) else lo
        for b in range(lo, hi + 1):
            for t in bank_tables(m.group(1), b):
                segs.append(dict(dst=off, table=t["id"], col=0, len=t["dim"], pad=False))
                off += t["dim"]
    real = [s for s in segs if not s["pad"]]
    assert sorted(s["table"] for s in real) == list(range(len(tables))), "every table exactly once"
    assert off == d["INPUT_SIZE"], (off, d["INPUT_SIZE"])
    assert off % 16 == 0
    return dict(name=model, source=f"FPGA/kernel/user_krnl/{krnl}/src/hls/constants.hpp",
                concat_spec=CONCAT_SPEC[model], n_tables=len(tables), concat_floats=off,
                data_floats=sum(t["dim"] for t in tables), hidden=hidden,
                fpga_batch=d["BATCH_SIZE"], tables=tables, segments=segs)


def build_large(half):
    """Full 377-table model: CPU0 64-float block, then FPGA0 and FPGA1 halves
    (GPU/final_network_cublasLt_3_nodes_no_FIFO_scatter/constant.h:25-27 and the
    receive order cuda_server.c:513-587), as a true per-item concat."""
    tables, segs = [], []
    cpu = dict(id=0, tier="CPU", tier_index=0, bank=0, round=0, rows=1000000, dim=64,
               axi_padded=16, addr_axi=0, half="cpu")
    tables.append(cpu)
    segs.append(dict(dst=0, table=0, col=0, len=64, pad=False))
    off = 64
    for h in ("fpga0", "fpga1"):
        base = len(tables)
        for t in half["tables"]:
            t2 = dict(t)
            t2["id"] = base + t["id"]
            t2["half"] = h
            tables.append(t2)
        for s in half["segments"]:
            s2 = dict(s)
            s2["dst"] = off + s["dst"]
            s2["table"] = base + s["table"]
            segs.append(s2)
        off += half["concat_floats"]
    assert off == 3968 and len(tables) == 377
    return dict(name="large", source=half["source"] + " x2 + CPU0 64-float block",
                concat_spec="C0 | " + half["concat_spec"] + " | " + half["concat_spec"],
                n_tables=len(tables), concat_floats=off,
                data_floats=sum(t["dim"] for t in tables), hidden=[2048, 512, 256, 1],
                fpga_batch=half["fpga_batch"], tables=tables, segments=segs)


def emit_c(models, path):
    tiers = {"HBM": 0, "DDR": 1, "PLRAM": 2, "CPU": 3}
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_catalogue.py from the reference's constants.hpp -- do not edit.\n")
        f.write("// Rows: {tier, tier_index, bank, round, rows, dim}; segments: {dst, table, col, len}.\n")
        for m in models:
            n = m["name"]
            f.write(f"static const fr_table_desc k_{n}_tables[] = {{\n")
            for t in m["tables"]:
                f.write(f"  {{{tiers[t['tier']]},{t['tier_index']},{t['bank']},{t['round']},"
                        f"{t['rows']}LL,{t['dim']}}},\n")
            f.write("};\n")
            f.write(f"static const fr_segment_desc k_{n}_segments[] = {{\n")
            for s in m["segments"]:
                f.write(f"  {{{s['dst']},{s['table']},{s['col']},{s['len']}}},\n")
            f.write("};\n")
        f.write("static const fr_builtin_model k_builtin_models[] = {\n")
        for m in models:
            n = m["name"]
            h = m["hidden"]
            f.write(f"  {{\"{n}\", {m['n_tables']}, k_{n}_tables, {len(m['segments'])}, k_{n}_segments, "
                    f"{m['concat_floats']}, {{{h[0]},{h[1]},{h[2]},{h[3]}}}}},\n")
        f.write("};\n")


def main():
    models = [build(m) for m in MODELS]
    models.append(build_large(models[2]))
    os.makedirs(os.path.join(PKG, "catalogue"), exist_ok=True)
    for m in models:
        with open(os.path.join(PKG, "catalogue", m["name"] + ".json"), "w") as f:
            json.dump(m, f, indent=0, separators=(",", ":"))
            f.write("\n")
        nbytes = sum(t["rows"] * t["dim"] * 4 for t in m["tables"])
        print(f"{m['name']:10s} tables={m['n_tables']:3d} concat={m['concat_floats']:4d} "
              f"data={m['data_floats']:4d} rows={sum(t['rows'] for t in m['tables']):,} "
              f"bytes={nbytes / 1e9:.3f} GB segs={len(m['segments'])}")
    emit_c(models, os.path.join(PKG, "csrc", "fr_catalogue_data.inc"))


if __name__ == "__main__":
    sys.exit(main())
