#!/usr/bin/env python3
"""Generate tests/golden/*.npz by EXECUTING the reference's own lookup kernels.

Needs /root/reference (build container only).  `make -C oracle ref` compiles
FPGA/kernel/user_krnl/embedding_{47,98,377}_krnl/src/hls/*.cpp against the
oracle/shim headers; this script runs each kernel's top function
(embedding_<N>_krnl) for 3 FPGA batches with
  * HBM/DDR tables = position-encoding hash fill (seed 0x5EED), rows capped at 128
    (the kernel's fixed index list, embedding_47_krnl.cpp:903-904, reads rows < 100),
    laid out in bank images at the reference's ADDR_AXI_* offsets (host.cpp style),
  * PLRAM tables = the kernel's own init_plram_* fill (even rows 1.0, odd rows 0.0),
and stores the emitted TCP payload (wire order, item-major fp32) as golden vectors,
plus the stream-level output of the reference's gather_embeddings() fed with
tagged words.  The vectors are committed; tests compare the oracle (CPU) and the
CUDA path against them without touching /root/reference.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
from fleetrec import catalogue  # noqa: E402
from oracle import oracle, ref  # noqa: E402

SEED = 0x5EED
ROW_CAP = 128


def golden_tables(cat):
    return [oracle.fill_reference(t.rows, t.dim) if t.tier == "PLRAM"
            else oracle.fill_hash(SEED, t.id, t.rows, t.dim) for t in cat.tables]


def main():
    ref.build()
    for m in ref.REF_MODELS:
        cat = catalogue.load(m).with_row_cap(ROW_CAP)
        out, written = ref.run_top(m, cat, golden_tables(cat), batch_num=3)
        assert out.shape == (96, cat.concat_floats) and written == out.size
        tagged = ref.run_gather_tagged(m, cat, 32)
        path = os.path.join(ROOT, "tests", "golden", f"ref_{m}.npz")
        np.savez_compressed(path, wire_first32=out[:32], wire96_sha256=hashlib.sha256(out.tobytes()).hexdigest(),
                            tagged_item0=tagged[0], tagged_item31=tagged[31], seed=SEED, row_cap=ROW_CAP)
        print(m, out.shape, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
