"""Host logic that needs no GPU: the C-ABI library loads and exports every symbol
include/fleetrec.h declares, built-in catalogues equal the JSON ones, descriptor
validation and the no-fallback failure."""
import ctypes as C
import os
import re

import pytest

import fleetrec
from fleetrec import _capi, catalogue, shard
from fleetrec.engine import model_desc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _capi.build()
    return _capi.lib()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "fleetrec.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    raw = C.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None


def test_builtin_catalogues_match_json(L):
    for name in catalogue.MODEL_NAMES:
        d = _capi.ModelDesc()
        assert L.fr_model_builtin(name.encode(), C.byref(d)) == _capi.FR_OK
        m = catalogue.load(name)
        assert (d.n_tables, d.n_segments, d.concat_floats) == (m.n_tables, len(m.segments), m.concat_floats)
        assert list(d.hidden) == m.hidden
        for i, t in enumerate(m.tables):
            assert (d.tables[i].rows, d.tables[i].dim, d.tables[i].bank, d.tables[i].round) == \
                   (t.rows, t.dim, t.bank, t.round)
        for i, s in enumerate(m.segments):
            assert (d.segments[i].dst, d.segments[i].table, d.segments[i].col, d.segments[i].len) == \
                   (s.dst, s.table, s.col, s.len)
    d = _capi.ModelDesc()
    assert L.fr_model_builtin(b"nope", C.byref(d)) == _capi.FR_ERR_INVALID
    assert b"unknown model" in L.fr_last_error(None)


def test_descriptor_validation_and_no_cpu_fallback(L):
    m = catalogue.load("small")
    d = model_desc(m)
    h = C.c_void_p()
    dev = (C.c_int * 1)(0)
    assert L.fr_create(C.byref(d), 2, dev, C.byref(h)) == _capi.FR_ERR_INVALID       # one engine = one GPU
    bad = model_desc(m)
    bad.concat_floats = 350
    assert L.fr_create(C.byref(bad), 1, dev, C.byref(h)) == _capi.FR_ERR_INVALID
    assert b"multiple of 16" in L.fr_last_error(None)
    bad = model_desc(m)
    bad.segments[3].len = 6
    assert L.fr_create(C.byref(bad), 1, dev, C.byref(h)) == _capi.FR_ERR_INVALID
    import torch
    if not torch.cuda.is_available():
        rc = L.fr_create(C.byref(d), 1, dev, C.byref(h))
        assert rc == _capi.FR_ERR_CUDA and not h.value
        assert b"no CPU fallback" in L.fr_last_error(None)
        with pytest.raises(fleetrec.FleetRecError):
            fleetrec.Engine(m)


def test_merge_index_is_int64(L):
    assert L.fr_merge_index(99_999_999, 9_999_999, 10_000_000) == 99_999_999 * 10_000_000 + 9_999_999


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpu-fpga-recommendation-system_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".inc")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle's", "").replace("the oracle", ""), (dp, f)


def test_header_enumerators_match_the_python_constants():
    """The option / hint / dtype enumerators of include/fleetrec.h and fleetrec._capi agree (the binding is hand-written)."""
    hdr = open(os.path.join(ROOT, "include", "fleetrec.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    vals = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", hdr, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            name, _, v = item.partition("=")
            nxt = int(v) if v.strip() else nxt
            vals[name.strip()] = nxt
            nxt += 1
    for name in ("FR_OPT_CUDA_GRAPHS", "FR_OPT_CHECK_INDICES", "FR_OPT_FUSE_LOOKUP", "FR_OPT_TILE_HINT", "FR_OPT_F16_OPERANDS",
                 "FR_OPT_INDEX_FORMAT", "FR_IDX_I32", "FR_IDX_PACKED", "FR_HINT_AUTO", "FR_HINT_LATENCY", "FR_HINT_THROUGHPUT", "FR_F16_OFF", "FR_F16_GUARDED", "FR_TABLE_F32",
                 "FR_TABLE_F16", "FR_TABLE_BF16", "FR_TABLE_FP8", "FR_PREC_TF32", "FR_PREC_FP32", "FR_MLP_LINEAR",
                 "FR_MLP_BIAS_RELU_SIGMOID", "FR_OK", "FR_ERR_INVALID", "FR_ERR_CUDA", "FR_ERR_OOM", "FR_ERR_STATE",
                 "FR_ERR_UNSUPPORTED", "FR_INGEST_CONCAT", "FR_INGEST_INDICES"):
        assert vals[name] == getattr(_capi, name), name


def test_release_library_has_no_experiments_and_reads_two_env_hooks(L):
    assert L.fr_build_has_experiments() == 0
    src = open(os.path.join(ROOT, "gpu-fpga-recommendation-system_b200", "csrc", "fr_api.cu")).read()
    csrc = os.path.join(ROOT, "gpu-fpga-recommendation-system_b200", "csrc")
    sites = [(f, i) for f in os.listdir(csrc) if f.endswith((".cu", ".h"))
             for i, line in enumerate(open(os.path.join(csrc, f))) if "getenv(" in line and not line.strip().startswith("//")]
    assert {f for f, _ in sites} == {"fr_api.cu"}, sites                 # one translation unit reads the environment
    release = src[src.index("static void fr_read_knobs"):src.index("#ifdef FR_EXPERIMENTS", src.index("static void fr_read_knobs"))]
    assert sorted(re.findall(r'getenv\("(\w+)"\)', release)) == ["FR_TC_MAX_CLUSTERS", "FR_TC_TILES"]


def _layout(L, rows, fmt):
    raw = C.CDLL(_capi.LIB_PATH)
    raw.frdbg_index_layout.argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int)]
    n = len(rows)
    off, words = (C.c_int32 * max(n, 1))(), C.c_int(0)
    assert raw.frdbg_index_layout((C.c_int64 * max(n, 1))(*rows), n, fmt, off, C.byref(words)) == 0
    return [int(off[i]) for i in range(n)], words.value


def test_packed_index_rows_layout_and_decode(L):
    """FR_IDX_PACKED on the host: the layout rule (int32 columns of the tables above 65536 rows first, then uint16
    columns, row padded to 4 bytes), fleetrec.pack_indices, and the lookup kernels' index decode (fr_index_at, the
    same inline function compiled for the host) agree for every column of ragged rows."""
    import numpy as np
    raw = C.CDLL(_capi.LIB_PATH)
    raw.frdbg_index_at.restype = C.c_int64
    raw.frdbg_index_at.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int]
    rng = np.random.default_rng(7)
    cases = [[1], [65536], [65537], [65536, 65537, 3, 10 ** 8, 65536], [5] * 7, [10 ** 7] * 3, [],
             [int(r) for r in rng.choice([4, 1000, 65535, 65536, 65537, 10 ** 6, 10 ** 8], size=47)]]
    cases.append([t.rows for t in catalogue.load("small").tables])
    for rows in cases:
        for fmt in (fleetrec.FR_IDX_I32, fleetrec.FR_IDX_PACKED):
            off, words = _layout(L, rows, fmt)
            wid = [2 if o < 0 else 4 for o in off]
            byte = [o & 0x7FFFFFFF for o in off]
            assert shard.index_layout(rows, packed=fmt == fleetrec.FR_IDX_PACKED) == (byte, wid, 4 * words)   # the host-side mirror
            if fmt == fleetrec.FR_IDX_I32:
                assert byte == [4 * i for i in range(len(rows))] and all(w == 4 for w in wid) and words == len(rows)
            else:
                assert wid == [4 if r > 65536 else 2 for r in rows]
                wide = [b for b, w in zip(byte, wid) if w == 4]
                narrow = [b for b, w in zip(byte, wid) if w == 2]
                assert wide == [4 * i for i in range(len(wide))]                       # table order, first
                assert narrow == [4 * len(wide) + 2 * i for i in range(len(narrow))]   # then the uint16 columns
                assert words == (4 * len(wide) + 2 * len(narrow) + 3) // 4
            B = 9
            idx = np.stack([rng.integers(0, r, size=B) for r in rows], axis=1).astype(np.int32) if rows else np.zeros((B, 0), np.int32)
            if rows:
                idx[0] = [r - 1 for r in rows]      # the largest index of every table
                idx[1] = 0
            packed = fleetrec.pack_indices(idx, (byte, wid, 4 * words))
            assert packed.dtype == np.int32 and packed.shape == (B, words)
            packed = np.ascontiguousarray(packed)
            for b in range(B):
                for c in range(len(rows)):
                    assert raw.frdbg_index_at(packed.ctypes.data, b, words, off[c]) == idx[b, c], (rows, fmt, b, c)
    small = [t.rows for t in catalogue.load("small").tables]
    assert 4 * _layout(L, small, fleetrec.FR_IDX_PACKED)[1] == 120 and 4 * _layout(L, small, fleetrec.FR_IDX_I32)[1] == 188


def test_lookup_descriptor_unpack_is_the_struct_layout(L):
    """The lookup kernels fetch a 32-byte piece descriptor as two 128-bit words and unpack the fields by hand
    (fr_gather.cu: unpack_chunk); the same function on the host reproduces any descriptor bit for bit."""
    import numpy as np
    raw = C.CDLL(_capi.LIB_PATH)
    raw.frdbg_chunk_roundtrip.argtypes = [C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(3)
    for _ in range(64):
        src = rng.integers(0, 2 ** 32, size=8, dtype=np.uint64).astype(np.uint32)
        src[7] = 0                                   # pad_
        out = np.full(8, 0xFFFFFFFF, np.uint32)
        raw.frdbg_chunk_roundtrip(src.ctypes.data, out.ctypes.data)
        assert np.array_equal(src, out)
