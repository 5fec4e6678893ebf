"""ctypes binding of libfleetrec.so -- the reference-side stub for Python hosts.

Every entry point of include/fleetrec.h is declared here (tests assert the list
is complete).  The library is built in-tree by ../Makefile (nvcc, sm_100a); there
is no fallback: a missing library raises, a missing GPU makes fr_create fail.
"""
import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# FLEETREC_LIB: another build of the same library (libfleetrec_exp.so, `make exp`: the experimental kernel variants)
LIB_PATH = os.environ.get("FLEETREC_LIB") or os.path.join(PKG_DIR, "libfleetrec.so")

FR_OK, FR_ERR_INVALID, FR_ERR_CUDA, FR_ERR_OOM, FR_ERR_STATE, FR_ERR_UNSUPPORTED = range(6)
FR_MLP_LINEAR, FR_MLP_BIAS_RELU_SIGMOID = 0, 1
FR_PREC_TF32, FR_PREC_FP32 = 0, 1
FR_TABLE_F32, FR_TABLE_F16, FR_TABLE_BF16, FR_TABLE_FP8 = 0, 1, 2, 3
FR_OPT_CUDA_GRAPHS, FR_OPT_CHECK_INDICES, FR_OPT_FUSE_LOOKUP, FR_OPT_TILE_HINT, FR_OPT_F16_OPERANDS, FR_OPT_INDEX_FORMAT = range(6)
FR_IDX_I32, FR_IDX_PACKED = 0, 1
FR_HINT_AUTO, FR_HINT_LATENCY, FR_HINT_THROUGHPUT = 0, 1, 2
FR_F16_OFF, FR_F16_GUARDED = 0, 1


class TableDesc(C.Structure):
    _fields_ = [("tier", C.c_int), ("tier_index", C.c_int), ("bank", C.c_int), ("round", C.c_int),
                ("rows", C.c_int64), ("dim", C.c_int)]


class SegmentDesc(C.Structure):
    _fields_ = [("dst", C.c_int), ("table", C.c_int), ("col", C.c_int), ("len", C.c_int)]


class ModelDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n_tables", C.c_int), ("tables", C.POINTER(TableDesc)),
                ("n_segments", C.c_int), ("segments", C.POINTER(SegmentDesc)), ("concat_floats", C.c_int),
                ("hidden", C.c_int * 4), ("mlp_mode", C.c_int), ("precision", C.c_int), ("max_batch", C.c_int),
                ("table_dtype", C.c_int)]


class BatcherConfig(C.Structure):
    _fields_ = [("max_batch", C.c_int), ("max_delay_us", C.c_int), ("n_workers", C.c_int)]


class BatcherStats(C.Structure):
    _fields_ = [("batches", C.c_int64), ("items", C.c_int64), ("requests", C.c_int64), ("closed_by_deadline", C.c_int64),
                ("latency_p50_us", C.c_float), ("latency_p99_us", C.c_float)]


class IngestConfig(C.Structure):
    _fields_ = [("base_port", C.c_int), ("n_conn", C.c_int), ("batch", C.c_int), ("payload", C.c_int),
                ("total_batches", C.c_int64), ("listen_any", C.c_int), ("scores_out", C.c_void_p),
                ("max_batches_per_conn", C.c_int64)]


class IngestStats(C.Structure):
    _fields_ = [("batches", C.c_int64), ("bytes", C.c_int64), ("seconds", C.c_double), ("connections", C.c_int)]


FR_INGEST_CONCAT, FR_INGEST_INDICES = 0, 1
_P, _I, _I64, _U32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32
# name -> (restype, argtypes); must list every symbol include/fleetrec.h declares
SIGNATURES = {
    "fr_model_builtin": (_I, [C.c_char_p, C.POINTER(ModelDesc)]),
    "fr_create": (_I, [C.POINTER(ModelDesc), _I, C.POINTER(C.c_int), C.POINTER(_P)]),
    "fr_destroy": (None, [_P]),
    "fr_last_error": (C.c_char_p, [_P]),
    "fr_set_table_rows": (_I, [_P, _I, _I64]),
    "fr_load_table": (_I, [_P, _I, _P, _I64, _I]),
    "fr_fill_table_reference": (_I, [_P, _I, _I64]),
    "fr_fill_table_hash": (_I, [_P, _I, _U32]),
    "fr_read_table": (_I, [_P, _I, _I64, _I64, _P]),
    "fr_load_mlp": (_I, [_P, _I, _P, _P]),
    "fr_set_mlp_mode": (_I, [_P, _I]),
    "fr_set_precision": (_I, [_P, _I]),
    "fr_set_option": (_I, [_P, _I, _I]),
    "fr_f16_report": (_I, [_P, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "fr_index_layout": (_I, [_P, _I, _P, _P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fr_stream_create": (_I, [_P, C.POINTER(_P)]),
    "fr_stream_destroy": (None, [_P, _P]),
    "fr_stream_cuda": (_P, [_P]),
    "fr_infer": (_I, [_P, _P, _I, _P, _P]),
    "fr_infer_many": (_I, [_P, _P, _I, _I, _P, _P]),
    "fr_gather_only": (_I, [_P, _P, _I, _P, _P]),
    "fr_mlp_only": (_I, [_P, _P, _I, _P, _P]),
    "fr_layer_only": (_I, [_P, _I, _P, _I, _P, _P]),
    "fr_sync": (_I, [_P, _P]),
    "fr_launch_count": (_I64, [_P]),
    "fr_table_bytes": (_I64, [_P]),
    "fr_graph_stats": (_I, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "fr_graph_flush": (_I, [_P, _P]),
    "fr_mark": (_I, [_P, _P, _I]),
    "fr_elapsed_ms": (_I, [_P, _P, C.POINTER(C.c_float)]),
    "fr_time_kernels": (_I, [_P, _P, _I, _I, _P, C.POINTER(C.c_float)]),
    "fr_shard_init": (_I, [_P, _I, _I, C.POINTER(C.c_int)]),
    "fr_shard_export": (_I, [_P, _P]),
    "fr_shard_import": (_I, [_P, _P]),
    "fr_shard_attach_local": (_I, [_P, C.POINTER(_P)]),
    "fr_shard_gather_push": (_I, [_P, _P, _I, _P]),
    "fr_shard_mlp": (_I, [_P, _I, _P, _P]),
    "fr_shard_infer": (_I, [_P, _P, _I, _P, _P]),
    "fr_shard_tables": (_I, [_P, _I, _P, _P]),
    "fr_shard_infer_sliced": (_I, [_P, _P, _P, _I, _P, _P]),
    "fr_shard_infer_sliced_many": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "fr_shard_read_concat": (_I, [_P, _I, _P, _P]),
    "fr_merge_index": (_I64, [_I64, _I64, _I64]),
    "fr_merge_tables": (_I, [_P, _I, _I, _I]),
    "fr_batcher_create": (_I, [_P, C.POINTER(BatcherConfig), C.POINTER(_P)]),
    "fr_batcher_submit": (_I, [_P, _P, _I, _P, C.POINTER(C.c_uint64)]),
    "fr_batcher_wait": (_I, [_P, C.c_uint64]),
    "fr_batcher_flush": (_I, [_P]),
    "fr_batcher_get_stats": (_I, [_P, C.POINTER(BatcherStats)]),
    "fr_batcher_destroy": (None, [_P]),
    "fr_ingest_start": (_I, [_P, C.POINTER(IngestConfig), C.POINTER(_P)]),
    "fr_ingest_wait": (_I, [_P, C.POINTER(IngestStats)]),
    "fr_ingest_last_scores": (_I, [_P, _I, _P, C.POINTER(C.c_int64)]),
    "fr_ingest_destroy": (None, [_P]),
    "fr_build_has_experiments": (_I, []),
}

_LIB = None


def build(force=False):
    """Compile libfleetrec.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", PKG_DIR, "libfleetrec.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}` "
                               "(there is no CPU or PyTorch fallback for the hot path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
