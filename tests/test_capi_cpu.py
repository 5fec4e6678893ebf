"""Host logic that needs no GPU: the C-ABI library loads and exports every symbol
include/fleetrec.h declares, built-in catalogues equal the JSON ones, descriptor
validation and the no-fallback failure."""
import ctypes as C
import os
import re

import pytest

import fleetrec
from fleetrec import _capi, catalogue
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
