"""numpy front-end of oracle/libfr_oracle.so (fr_oracle.c).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Every
function here restates a reference function; see fr_oracle.h for file:line.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Seg(C.Structure):
    _fields_ = [("dst", C.c_int), ("table", C.c_int), ("col", C.c_int), ("len", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "libfr_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.fro_fill_reference.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64]
        L.fro_hash_bits.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32]
        L.fro_hash_bits.restype = C.c_uint32
        L.fro_fill_hash.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int64, C.c_int]
        L.fro_idx_reference.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fro_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_int]
        L.fro_mlp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_int]
        L.fro_merge_index.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.fro_merge_index.restype = C.c_int64
        L.fro_merge_tables.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.fro_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads():
    return lib().fro_max_threads()


def fill_reference(rows, dim, debug_rows=0):
    """host.cpp:66-88 / embedding_47_krnl.cpp:871-897."""
    t = np.empty((rows, dim), np.float32)
    lib().fro_fill_reference(t.ctypes.data, rows, dim, debug_rows)
    return t


def fill_hash(seed, table_id, rows, dim):
    t = np.empty((rows, dim), np.float32)
    lib().fro_fill_hash(t.ctypes.data, seed, table_id, rows, dim)
    return t


def hash_bits(seed, table, row, col):
    return lib().fro_hash_bits(seed, table, row, col)


def idx_reference(B, T):
    """embedding_47_krnl.cpp:899-914: the same 32-entry list for every table."""
    idx = np.empty((B, T), np.int32)
    lib().fro_idx_reference(idx.ctypes.data, B, T)
    return idx


def _seg_array(model):
    segs = (_Seg * len(model.segments))()
    for i, s in enumerate(model.segments):
        segs[i] = _Seg(s.dst, s.table, s.col, s.len)
    return segs


def gather(model, tables, idx, threads=0):
    """Rows L1-L4: lookup + concat in reference wire order -> [B][concat_floats]."""
    idx = np.ascontiguousarray(idx, np.int32)
    B, T = idx.shape
    assert T == model.n_tables == len(tables)
    ptrs = (C.c_void_p * T)()
    dims = (C.c_int * T)()
    for i, t in enumerate(tables):
        assert t.dtype == np.float32 and t.flags.c_contiguous and t.shape[1] == model.tables[i].dim
        assert idx[:, i].max(initial=0) < t.shape[0] and idx[:, i].min(initial=0) >= 0
        ptrs[i] = t.ctypes.data
        dims[i] = t.shape[1]
    out = np.empty((B, model.concat_floats), np.float32)
    segs = _seg_array(model)
    lib().fro_gather(ptrs, dims, segs, len(model.segments), idx.ctypes.data, T, B, model.concat_floats,
                     out.ctypes.data, threads)
    return out


def mlp(x, dims, W, bias, mode, acc64=False, threads=0):
    """cuda_server.c:468-491.  W[k]: [in_k][out_k]; mode 0 LINEAR, 1 BIAS_RELU_SIGMOID."""
    x = np.ascontiguousarray(x, np.float32)
    B = x.shape[0]
    assert x.shape[1] == dims[0]
    Wp = (C.c_void_p * 4)()
    bp = (C.c_void_p * 4)()
    keep = []
    for k in range(4):
        w = np.ascontiguousarray(W[k], np.float32)
        assert w.shape == (dims[k], dims[k + 1])
        keep.append(w)
        Wp[k] = w.ctypes.data
        if bias is not None and bias[k] is not None:
            b = np.ascontiguousarray(bias[k], np.float32)
            keep.append(b)
            bp[k] = b.ctypes.data
        else:
            bp[k] = None
    d = (C.c_int * 5)(*dims)
    out = np.empty((B,), np.float32)
    lib().fro_mlp(x.ctypes.data, B, d, Wp, bp if bias is not None else None, mode, int(acc64), out.ctypes.data,
                  threads)
    return out


def merge_index(iA, iB, rowsB):
    return lib().fro_merge_index(iA, iB, rowsB)


def merge_tables(A, B):
    M = np.empty((A.shape[0] * B.shape[0], A.shape[1] + B.shape[1]), np.float32)
    lib().fro_merge_tables(A.ctypes.data, A.shape[0], A.shape[1], B.ctypes.data, B.shape[0], B.shape[1],
                           M.ctypes.data)
    return M


# ---- workload generators shared by tests and bench (host side, numpy) -------
def make_tables(model, fill="hash", seed=0x5EED):
    if fill == "hash":
        return [fill_hash(seed, t.id, t.rows, t.dim) for t in model.tables]
    if fill == "reference":
        return [fill_reference(t.rows, t.dim) for t in model.tables]
    raise ValueError(fill)


def make_weights(dims, seed=42, mode=1):
    """W ~ N(0, 1/in), b ~ N(0, 0.01)."""
    rng = np.random.default_rng(seed)
    W = [(rng.standard_normal((dims[k], dims[k + 1])) / np.sqrt(dims[k])).astype(np.float32) for k in range(4)]
    b = [(0.01 * rng.standard_normal(dims[k + 1])).astype(np.float32) for k in range(4)]
    return W, b


def zipf_indices(model, B, s=1.05, seed=1234):
    """Per-table Zipf(s) row indices over rows_t (bounded, inverse-CDF on a
    power-law approximation; rank r drawn with p ~ r^-s, then scattered by a
    fixed multiplicative permutation so hot rows are not physically adjacent)."""
    rng = np.random.default_rng(seed)
    idx = np.empty((B, model.n_tables), np.int32)
    for t in model.tables:
        n = t.rows
        u = rng.random(B)
        # inverse CDF of the continuous density x^-s on [1, n+1)
        a = 1.0 - s
        r = ((u * ((n + 1.0) ** a - 1.0) + 1.0) ** (1.0 / a)).astype(np.int64) - 1
        r = np.clip(r, 0, n - 1)
        idx[:, t.id] = (r * 2654435761 % n).astype(np.int32)
    return idx


def uniform_indices(model, B, seed=4321):
    rng = np.random.default_rng(seed)
    idx = np.empty((B, model.n_tables), np.int32)
    for t in model.tables:
        idx[:, t.id] = rng.integers(0, t.rows, B, dtype=np.int64).astype(np.int32)
    return idx


def hash_rows(seed, table_id, rows, dim):
    """Vectorised fro_hash_bits: the [len(rows)][dim] fp32 values a hash-filled table
    holds at the given row indices, computed without materialising the table
    (full-size parity checks: 12.8 GB tables never exist on the host)."""
    rows = np.asarray(rows, np.uint64).reshape(-1, 1)
    cols = np.arange(dim, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        z = rows * np.uint64(0x9E3779B97F4A7C15) + (np.uint64(table_id) << np.uint64(40)) \
            + (cols << np.uint64(28)) + np.uint64(seed)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    h = (z >> np.uint64(16)).astype(np.uint32)
    expo = np.uint32(118) + ((h >> np.uint32(23)) & np.uint32(0xFF)) % np.uint32(9)
    bits = (h & np.uint32(0x80000000)) | (expo << np.uint32(23)) | (h & np.uint32(0x007FFFFF))
    return bits.view(np.float32)


def gather_hashed(model, seed, idx):
    """oracle.gather() for hash-filled tables of any size (no table memory)."""
    idx = np.asarray(idx)
    out = np.empty((idx.shape[0], model.concat_floats), np.float32)
    cache = {}
    for s in model.segments:
        if s.table not in cache:
            cache[s.table] = hash_rows(seed, s.table, idx[:, s.table], model.tables[s.table].dim)
        out[:, s.dst:s.dst + s.len] = cache[s.table][:, s.col:s.col + s.len]
    return out


# ---- reduced-precision table storage: the STATED dequant --------------------
def quantize_dequantize(x, table_dtype):
    """What a table stored as f16 / bf16 returns for fp32 contents x: round to nearest even into
    the 2-byte (or 1-byte) type, widen exactly back to fp32.  table_dtype: 0 fp32 (identity), 1 f16, 2 bf16, 3 fp8 e4m3.
    Quantisation is element-wise, so it commutes with the lookup: gather(quantised tables) ==
    quantize_dequantize(gather(fp32 tables))."""
    x = np.ascontiguousarray(x, np.float32)
    if table_dtype == 0:
        return x
    if table_dtype == 1:
        with np.errstate(over="ignore"):
            return x.astype(np.float16).astype(np.float32)
    if table_dtype == 2:
        b = x.view(np.uint32).astype(np.uint64)
        r = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000            # RNE on the upper 16 bits
        return r.astype(np.uint32).view(np.float32).reshape(x.shape)
    if table_dtype == 3:
        return _e4m3_round(x)
    raise ValueError(table_dtype)


def _e4m3_round(x):
    """fp32 -> FP8 E4M3 (1-4-3, bias 7, no infinities, max 448) -> fp32: round to nearest, ties to the code with an even
    mantissa, saturating to +-448 (what __nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3) stores); NaN stays NaN."""
    codes = np.arange(127, dtype=np.int64)                       # 0x00 .. 0x7E: the non-negative finite values, ascending
    e, m = codes >> 3, codes & 7
    grid = np.where(e == 0, m * 2.0 ** -9, (1 + m / 8.0) * 2.0 ** (e - 7.0))
    a = np.abs(x.astype(np.float64))
    hi = np.clip(np.searchsorted(grid, a, side="left"), 0, 126)  # first grid value >= |x| (or the maximum)
    lo = np.clip(hi - 1, 0, 126)
    d_lo, d_hi = a - grid[lo], grid[hi] - a
    pick_hi = (d_hi < d_lo) | ((d_hi == d_lo) & (hi % 2 == 0))   # tie -> even code
    q = np.where(a >= grid[126], grid[126], np.where(pick_hi, grid[hi], grid[lo]))
    out = np.copysign(q, x).astype(np.float32)
    return np.where(np.isnan(x), np.float32(np.nan), out).astype(np.float32).reshape(x.shape)
