"""Front-end of oracle/_ref/libref_<model>.so: the REFERENCE's own HLS lookup
kernel executed on the CPU (oracle/ref_harness.cpp).  TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_MODELS = ("small", "medium", "large_half")


def available(model):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{model}.so"))


def build():
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _lib(model):
    L = C.CDLL(os.path.join(_HERE, "_ref", f"libref_{model}.so"))
    L.ref_run_top.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
    L.ref_run_gather_tagged.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    return L


def info(model):
    L = _lib(model)
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    L.ref_info(C.byref(a), C.byref(b), C.byref(c))
    return dict(input_size=a.value, n_plram=b.value, fpga_batch=c.value)


def bank_images(cat, tables):
    """Lay HBM/DDR tables out as host.cpp does: one flat image per bank, table rows
    back to back at ADDR_AXI_* (axi units = 4 floats).  Order: HBM0..27, DDR0, DDR1."""
    imgs = []
    for tier, nb in (("HBM", 28), ("DDR", 2)):
        for b in range(nb):
            ts = [t for t in cat.tables if t.tier == tier and t.bank == b]
            size = max(t.addr_axi * 4 + t.rows * t.dim for t in ts)
            img = np.zeros(size, np.float32)
            for t in ts:
                img[t.addr_axi * 4: t.addr_axi * 4 + t.rows * t.dim] = tables[t.id].reshape(-1)
            imgs.append(img)
    return imgs


def run_top(model, cat, tables, batch_num=3, use_conn=4, pkg_word_count=16):
    """Run embedding_<N>_krnl() end to end; returns the tx stream as [items][INPUT_SIZE]
    (only complete items; sendData's credit accounting leaves the last packet unsent)."""
    L = _lib(model)
    imgs = bank_images(cat, tables)
    ptrs = (C.c_void_p * 30)(*[i.ctypes.data for i in imgs])
    n = batch_num * 32 * cat.concat_floats
    out = np.zeros(n, np.float32)
    written = C.c_long()
    rc = L.ref_run_top(ptrs, batch_num, use_conn, pkg_word_count, out.ctypes.data, n, C.byref(written))
    assert rc == 0
    items = written.value // cat.concat_floats
    return out[: items * cat.concat_floats].reshape(items, cat.concat_floats), written.value


def stream_words(cat):
    """axi words per item on each lookup stream, order HBM0..27, DDR0, DDR1, PLRAM0.."""
    n_pl = max(t.bank for t in cat.tables if t.tier == "PLRAM") + 1
    words = []
    for tier, nb in (("HBM", 28), ("DDR", 2), ("PLRAM", n_pl)):
        for b in range(nb):
            words.append(sum(t.dim for t in cat.tables if t.tier == tier and t.bank == b) // 4)
    return words


def run_gather_tagged(model, cat, n_items=32):
    L = _lib(model)
    words = stream_words(cat)
    w = (C.c_int * len(words))(*words)
    out = np.zeros((n_items, cat.concat_floats), np.float32)
    rc = L.ref_run_gather_tagged(n_items, w, out.ctypes.data)
    assert rc == 0, rc
    return out
