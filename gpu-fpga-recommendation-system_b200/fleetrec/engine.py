"""Host-side mirror of the reference's two host programs, over the C ABI.

  Engine.load_* / fill_*   <- FPGA/host/embedding_47_krnl/host.cpp:324-750 (table images into device memory)
  Engine.load_mlp          <- cuda_server.c:152-160,346-354 (weights H2D once)
  Engine.infer             <- one trip of the hot loop: lookup kernel -> wire -> cuda_server.c:406-495
  Worker                   <- one thread_consume() worker (cuda_server.c:101): own stream + buffers

numpy arrays are host buffers; objects exposing data_ptr() (torch CUDA tensors)
are passed through as device pointers.  Errors raise FleetRecError carrying
fr_last_error(); nothing falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from .catalogue import TIERS, Model


class FleetRecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fleetrec error {code}: {msg}")
        self.code = code


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(type(a))


def model_desc(model: Model, mlp_mode=_capi.FR_MLP_BIAS_RELU_SIGMOID, precision=_capi.FR_PREC_TF32,
               max_batch=16384, table_dtype=_capi.FR_TABLE_F32):
    d = _capi.ModelDesc()
    tabs = (_capi.TableDesc * model.n_tables)()
    for i, t in enumerate(model.tables):
        tabs[i] = _capi.TableDesc(TIERS[t.tier], t.tier_index, t.bank, t.round, t.rows, t.dim)
    segs = (_capi.SegmentDesc * len(model.segments))()
    for i, s in enumerate(model.segments):
        segs[i] = _capi.SegmentDesc(s.dst, s.table, s.col, s.len)
    d.name = model.name.encode()
    d.n_tables, d.tables = model.n_tables, tabs
    d.n_segments, d.segments = len(model.segments), segs
    d.concat_floats = model.concat_floats
    d.hidden = (C.c_int * 4)(*model.hidden)
    d.mlp_mode, d.precision, d.max_batch = mlp_mode, precision, max_batch
    d.table_dtype = table_dtype
    d._keep = (tabs, segs)
    return d


class Worker:
    """One in-flight batch: a CUDA stream + activation workspaces (fr_stream)."""

    def __init__(self, engine):
        self.engine = engine
        h = C.c_void_p()
        engine._chk(engine._L.fr_stream_create(engine._h, C.byref(h)))
        self._h = h

    @property
    def cuda_stream(self):
        return self.engine._L.fr_stream_cuda(self._h)

    def close(self):
        if self._h:
            self.engine._L.fr_stream_destroy(self.engine._h, self._h)
            self._h = None


class Engine:
    def __init__(self, model: Model, device=0, mlp_mode=_capi.FR_MLP_BIAS_RELU_SIGMOID,
                 precision=_capi.FR_PREC_TF32, max_batch=16384, table_dtype=_capi.FR_TABLE_F32):
        self._L = _capi.lib()
        self.model = model
        self._desc = model_desc(model, mlp_mode, precision, max_batch, table_dtype)
        h = C.c_void_p()
        dev = (C.c_int * 1)(device)
        rc = self._L.fr_create(C.byref(self._desc), 1, dev, C.byref(h))
        if rc != _capi.FR_OK:
            raise FleetRecError(rc, (self._L.fr_last_error(None) or b"").decode())
        self._h = h
        self.max_batch = max_batch

    # -- plumbing ---------------------------------------------------------
    def _chk(self, rc):
        if rc != _capi.FR_OK:
            raise FleetRecError(rc, (self._L.fr_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.fr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tables -----------------------------------------------------------
    def set_table_rows(self, t, rows):
        self._chk(self._L.fr_set_table_rows(self._h, t, rows))

    def load_table(self, t, rows_array):
        a = np.ascontiguousarray(rows_array, np.float32)
        self._chk(self._L.fr_load_table(self._h, t, a.ctypes.data, a.shape[0], a.shape[1]))

    def load_tables(self, tables):
        for t, a in enumerate(tables):
            self.load_table(t, a)

    def fill_reference(self, debug_rows=0):
        for t in range(self.model.n_tables):
            self._chk(self._L.fr_fill_table_reference(self._h, t, debug_rows))

    def fill_hash(self, seed=0x5EED):
        for t in range(self.model.n_tables):
            self._chk(self._L.fr_fill_table_hash(self._h, t, seed))

    def read_table(self, t, first_row, n_rows):
        out = np.empty((n_rows, self.model.tables[t].dim), np.float32)
        self._chk(self._L.fr_read_table(self._h, t, first_row, n_rows, out.ctypes.data))
        return out

    # -- MLP --------------------------------------------------------------
    def load_mlp(self, W, bias=None):
        for k in range(4):
            w = np.ascontiguousarray(W[k], np.float32)
            b = None if bias is None or bias[k] is None else np.ascontiguousarray(bias[k], np.float32)
            self._chk(self._L.fr_load_mlp(self._h, k, w.ctypes.data, None if b is None else b.ctypes.data))

    def set_mlp_mode(self, mode):
        self._chk(self._L.fr_set_mlp_mode(self._h, mode))

    def set_precision(self, p):
        self._chk(self._L.fr_set_precision(self._h, p))

    def set_option(self, option, value):
        """fr_set_option: FR_OPT_CUDA_GRAPHS / CHECK_INDICES / FUSE_LOOKUP / TILE_HINT / F16_OPERANDS."""
        self._chk(self._L.fr_set_option(self._h, option, int(value)))

    def index_layout(self, which=2):
        """(byte_offsets, widths, row_bytes) of one index row under FR_OPT_INDEX_FORMAT: which = 2 full rows (all tables),
        0 / 1 the owned / replicated column blocks of the sliced sharded step."""
        n, rb = C.c_int(0), C.c_int(0)
        self._chk(self._L.fr_index_layout(self._h, which, None, None, C.byref(n), C.byref(rb)))
        off = (C.c_int32 * max(n.value, 1))()
        wid = (C.c_int32 * max(n.value, 1))()
        self._chk(self._L.fr_index_layout(self._h, which, off, wid, C.byref(n), C.byref(rb)))
        return [int(off[i]) for i in range(n.value)], [int(wid[i]) for i in range(n.value)], rb.value

    def f16_report(self):
        """(active, bounds): does fr_infer compute on fp16 operands, and the range bounds that decided it
        ([max |x|, max |h1|, max |h2|, min non-zero |table element|, inexact weight-mass share])."""
        active = C.c_int(0)
        b = (C.c_float * 5)()
        self._chk(self._L.fr_f16_report(self._h, C.byref(active), b))
        return bool(active.value), [float(v) for v in b]

    # -- hot path ---------------------------------------------------------
    def infer_async(self, idx, scores, B=None, worker=None):
        B = idx.shape[0] if B is None else B
        self._chk(self._L.fr_infer(self._h, _ptr(idx), B, _ptr(scores), worker._h if worker else None))

    def infer(self, idx, worker=None):
        idx = np.ascontiguousarray(idx, np.int32)
        scores = np.empty(idx.shape[0], np.float32)
        self.infer_async(idx, scores, worker=worker)
        self.sync(worker)
        return scores

    def infer_many_async(self, idx, scores, n, B, worker=None):
        """fr_infer_many: n batches of B items, idx [n][B][T], scores [n][B]; one copy each way for host buffers."""
        self._chk(self._L.fr_infer_many(self._h, _ptr(idx), n, B, _ptr(scores), worker._h if worker else None))

    def gather_only(self, idx, worker=None):
        idx = np.ascontiguousarray(idx, np.int32)
        out = np.empty((idx.shape[0], self.model.concat_floats), np.float32)
        self._chk(self._L.fr_gather_only(self._h, idx.ctypes.data, idx.shape[0], out.ctypes.data,
                                         worker._h if worker else None))
        self.sync(worker)
        return out

    def gather_only_async(self, idx, out, B, worker=None):
        self._chk(self._L.fr_gather_only(self._h, _ptr(idx), B, _ptr(out), worker._h if worker else None))

    def mlp_only(self, x, worker=None):
        x = np.ascontiguousarray(x, np.float32)
        scores = np.empty(x.shape[0], np.float32)
        self._chk(self._L.fr_mlp_only(self._h, x.ctypes.data, x.shape[0], scores.ctypes.data,
                                      worker._h if worker else None))
        self.sync(worker)
        return scores

    def mlp_only_async(self, x, scores, B, worker=None):
        self._chk(self._L.fr_mlp_only(self._h, _ptr(x), B, _ptr(scores), worker._h if worker else None))

    def layer_only(self, k, x, n_out, worker=None):
        """One launch of the MLP chain alone (unit-test hook); n_out = out width, 1 for scores."""
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty((x.shape[0], n_out) if n_out > 1 else (x.shape[0],), np.float32)
        self._chk(self._L.fr_layer_only(self._h, k, x.ctypes.data, x.shape[0], y.ctypes.data,
                                        worker._h if worker else None))
        self.sync(worker)
        return y

    def sync(self, worker=None):
        self._chk(self._L.fr_sync(self._h, worker._h if worker else None))

    # -- introspection ----------------------------------------------------
    def launch_count(self):
        return self._L.fr_launch_count(self._h)

    def table_bytes(self):
        return self._L.fr_table_bytes(self._h)

    def graph_stats(self):
        """{replayed, captured, direct}: hot-path steps served from a cached CUDA graph / captured first / plain launches."""
        r, c, d = C.c_int64(), C.c_int64(), C.c_int64()
        self._chk(self._L.fr_graph_stats(self._h, C.byref(r), C.byref(c), C.byref(d)))
        return {"replayed": r.value, "captured": c.value, "direct": d.value}

    def graph_flush(self, worker=None):
        self._chk(self._L.fr_graph_flush(self._h, worker._h if worker else None))

    def mark(self, which, worker=None):
        self._chk(self._L.fr_mark(self._h, worker._h if worker else None, which))

    def elapsed_ms(self, worker=None):
        ms = C.c_float()
        self._chk(self._L.fr_elapsed_ms(self._h, worker._h if worker else None, C.byref(ms)))
        return ms.value

    def time_kernels(self, idx, B=None, reps=20, worker=None):
        """avg ms per launch of [gather, layer1, layer2, layer3(+4), output layer]."""
        B = idx.shape[0] if B is None else B
        ms = (C.c_float * 5)()
        self._chk(self._L.fr_time_kernels(self._h, _ptr(idx), B, reps, worker._h if worker else None, ms))
        return list(ms)

    # -- sharding ---------------------------------------------------------
    def shard_init(self, rank, world, owner):
        arr = (C.c_int * len(owner))(*owner)
        self._chk(self._L.fr_shard_init(self._h, rank, world, arr))
        self.rank, self.world = rank, world

    def shard_export(self):
        buf = C.create_string_buffer(64)
        self._chk(self._L.fr_shard_export(self._h, buf))
        return buf.raw

    def shard_import(self, handles_bytes):
        self._chk(self._L.fr_shard_import(self._h, handles_bytes))

    def shard_attach_local(self, engines):
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        self._chk(self._L.fr_shard_attach_local(self._h, arr))

    def shard_gather_push(self, idx, B_global=None, worker=None):
        B_global = idx.shape[0] if B_global is None else B_global
        self._chk(self._L.fr_shard_gather_push(self._h, _ptr(idx), B_global, worker._h if worker else None))

    def shard_mlp(self, B_global, scores, worker=None):
        self._chk(self._L.fr_shard_mlp(self._h, B_global, _ptr(scores), worker._h if worker else None))

    def shard_infer(self, idx, B_global, scores, worker=None):
        """One-call sharded step (device-side sync); scores: [B_global/world]."""
        self._chk(self._L.fr_shard_infer(self._h, _ptr(idx), B_global, _ptr(scores), worker._h if worker else None))

    def shard_tables(self, which):
        """Table ids this rank needs indices for, ascending: which=0 owned (all items), 1 replicated (its items)."""
        n = C.c_int(0)
        self._chk(self._L.fr_shard_tables(self._h, which, None, C.byref(n)))
        ids = (C.c_int32 * max(n.value, 1))()
        self._chk(self._L.fr_shard_tables(self._h, which, ids, C.byref(n)))
        return [int(ids[i]) for i in range(n.value)]

    def shard_infer_sliced(self, idx_owned, idx_repl, B_global, scores, worker=None):
        """fr_shard_infer from column-sliced blocks: idx_owned [B_global][owned], idx_repl [B_global/world][replicated]."""
        self._chk(self._L.fr_shard_infer_sliced(self._h, _ptr(idx_owned), _ptr(idx_repl), B_global, _ptr(scores),
                                                worker._h if worker else None))

    def shard_infer_sliced_many(self, idx_owned, idx_repl, n, B_global, scores, worker=None):
        """n consecutive sharded steps in one call: idx_owned [n][B_global][owned], idx_repl [n][B_global/world][repl],
        scores [n][B_global/world]."""
        self._chk(self._L.fr_shard_infer_sliced_many(self._h, _ptr(idx_owned), _ptr(idx_repl), n, B_global, _ptr(scores),
                                                     worker._h if worker else None))

    def shard_read_concat(self, B_global, worker=None):
        out = np.empty((B_global // self.world, self.model.concat_floats), np.float32)
        self._chk(self._L.fr_shard_read_concat(self._h, B_global, out.ctypes.data, worker._h if worker else None))
        self.sync(worker)
        return out

    def merge_tables(self, a, b, dst):
        self._chk(self._L.fr_merge_tables(self._h, a, b, dst))


class Batcher:
    """Request-driven batch former in front of Engine.infer (fr_batcher_*):
    replaces the reference's fixed batch hand-out (cuda_server.c:406-417).  submit() may be called
    from any number of threads (ctypes releases the GIL)."""

    def __init__(self, engine, max_batch, max_delay_us=200, n_workers=4):
        self.engine = engine
        cfg = _capi.BatcherConfig(max_batch, max_delay_us, n_workers)
        h = C.c_void_p()
        engine._chk(engine._L.fr_batcher_create(engine._h, C.byref(cfg), C.byref(h)))
        self._h = h

    def submit(self, idx, scores_out):
        """idx [n][T] int32 (copied), scores_out [n] float32 (written later); returns a ticket."""
        idx = np.ascontiguousarray(idx, np.int32)
        assert scores_out.dtype == np.float32 and scores_out.flags.c_contiguous and scores_out.shape[0] == idx.shape[0]
        t = C.c_uint64()
        self.engine._chk(self.engine._L.fr_batcher_submit(self._h, idx.ctypes.data, idx.shape[0], scores_out.ctypes.data,
                                                          C.byref(t)))
        return t.value

    def wait(self, ticket):
        self.engine._chk(self.engine._L.fr_batcher_wait(self._h, ticket))

    def flush(self):
        self.engine._chk(self.engine._L.fr_batcher_flush(self._h))

    def stats(self):
        s = _capi.BatcherStats()
        self.engine._chk(self.engine._L.fr_batcher_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def close(self):
        if self._h:
            self.engine._L.fr_batcher_destroy(self._h)
            self._h = None


class Ingest:
    """B2-compatible TCP ingest (fr_ingest_*): listens on base_port + i like the
    reference's cuda_server.c, runs the MLP (payload='concat') or lookup + MLP (payload='indices')
    on every block received."""

    def __init__(self, engine, base_port, n_conn, batch, payload="concat", total_batches=0, max_batches_per_conn=0,
                 listen_any=False):
        self.engine = engine
        self.scores = np.zeros((n_conn, max(max_batches_per_conn, 1), batch), np.float32)
        cfg = _capi.IngestConfig(base_port, n_conn, batch,
                                 _capi.FR_INGEST_CONCAT if payload == "concat" else _capi.FR_INGEST_INDICES,
                                 total_batches, int(listen_any),
                                 self.scores.ctypes.data if max_batches_per_conn > 0 else None, max_batches_per_conn)
        h = C.c_void_p()
        engine._chk(engine._L.fr_ingest_start(engine._h, C.byref(cfg), C.byref(h)))
        self._h = h

    def wait(self):
        st = _capi.IngestStats()
        self.engine._chk(self.engine._L.fr_ingest_wait(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def last_scores(self, conn):
        out = np.empty(self.scores.shape[2], np.float32)
        no = C.c_int64()
        self.engine._chk(self.engine._L.fr_ingest_last_scores(self._h, conn, out.ctypes.data, C.byref(no)))
        return out, no.value

    def close(self):
        if self._h:
            self.engine._L.fr_ingest_destroy(self._h)
            self._h = None


def pack_indices(idx, layout):
    """int32 index rows [B][n] -> rows in the layout Engine.index_layout() describes ([B][row_bytes] uint8, viewable as
    int32 [B][row_bytes / 4]): int32 columns stay int32, uint16 columns are narrowed (they belong to tables of at most
    65536 rows).  With the FR_IDX_I32 layout this is a reinterpretation of the same bytes."""
    off, wid, row_bytes = layout
    idx = np.ascontiguousarray(idx, np.int32)
    assert idx.ndim == 2 and idx.shape[1] == len(off)
    out = np.zeros((idx.shape[0], row_bytes), np.uint8)
    for c, (o, w) in enumerate(zip(off, wid)):
        col = idx[:, c]
        if w == 2:
            assert col.min(initial=0) >= 0 and col.max(initial=0) < 65536
            out[:, o:o + 2] = col.astype("<u2").view(np.uint8).reshape(-1, 2)
        else:
            out[:, o:o + 4] = col.astype("<i4").view(np.uint8).reshape(-1, 4)
    return out.view(np.int32)


def merge_index(iA, iB, rowsB):
    return _capi.lib().fr_merge_index(iA, iB, rowsB)
