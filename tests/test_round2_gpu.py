"""GPU parity tests added in round 2: full-size / 64-bit addressing, the large model end to end, the
table-sharded step on every model, CUDA-graph cache behaviour, fr_infer_many, index checking, the guarded
fp16-operand path, tile hints.  Everything goes through the C ABI (ctypes -> libfleetrec.so)."""
import os
import sys
import time

import numpy as np
import pytest

import fleetrec
from fleetrec import catalogue, shard
from oracle import oracle

sys.path.insert(0, os.path.dirname(__file__))
import f16_bound_ref  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel_err(got, exp):
    return float(np.max(np.abs(got - exp) / np.maximum(np.abs(exp), 1e-6)))


def assert_bits_equal(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _small_engine(seed=3, cap=20000, max_batch=2048, **kw):
    cat = catalogue.load("small").with_row_cap(cap)
    tables = oracle.make_tables(cat, "hash", seed=seed)
    W, b = oracle.make_weights(cat.layer_dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=max_batch, **kw)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    return cat, tables, W, b, eng


# ---------------------------------------------------------------- 64-bit addressing at full size
def test_gather_beyond_int32_offsets_medium_ddr3_full_size():
    """the largest table (medium DDR_3, 100 M rows x 32 floats = 12.8 GB) needs 64-bit row
    addressing (embedding_47_krnl.cpp:927-929 widens to long): rows past 2^26 have FLOAT offsets past 2^31 and rows
    past 2^25 BYTE offsets past 2^32.  The table is hash-filled at full size on the device; the other 97 tables are
    capped.  Rows {0, 2^25 +- 1, 2^26 + 1, 2^26 + 2^24 + 5, rows - 1} and 200 uniform rows must come back bit-exact."""
    cat = catalogue.load("medium").with_row_cap(1000)
    big = max(catalogue.load("medium").tables, key=lambda t: t.rows * t.dim)
    assert big.rows == 100_000_000 and big.dim == 32
    cat.tables[big.id].rows = big.rows
    eng = fleetrec.Engine(cat, max_batch=256)
    eng.fill_hash(seed=0x5EED)
    assert eng.table_bytes() > 12.8e9
    idx = oracle.uniform_indices(cat, 206, seed=9)
    idx[:6, big.id] = [0, (1 << 25) - 1, (1 << 25) + 1, (1 << 26) + 1, (1 << 26) + (1 << 24) + 5, big.rows - 1]
    assert idx[:, big.id].max() < big.rows
    assert int(idx[:, big.id].astype(np.int64).max()) * big.dim * 4 > 1 << 32
    assert_bits_equal(eng.gather_only(idx), oracle.gather_hashed(cat, 0x5EED, idx))
    # and the device image itself at the far end (read back through the ABI)
    assert_bits_equal(eng.read_table(big.id, big.rows - 3, 3),
                      oracle.hash_rows(0x5EED, big.id, np.arange(big.rows - 3, big.rows), big.dim))
    eng.close()


# ---------------------------------------------------------------- large model, lookup -> MLP, full size
@pytest.mark.parametrize("B", (1000, 4096))
def test_large_model_full_size_end_to_end(B):
    """The 377-table model (3968-2048-512-256-1, 63.5 GB of tables, two 12.8 GB ones) through fr_infer on ONE GPU:
    concat bit-exact against the oracle's hash fill, scores within 1e-3 of the oracle's fp32 MLP."""
    cat = catalogue.load("large")
    dims = cat.layer_dims
    eng = fleetrec.Engine(cat, max_batch=4096)
    eng.fill_hash(seed=0x5EED)
    assert eng.table_bytes() == cat.table_bytes() > 60e9
    W, b = oracle.make_weights(dims, seed=42)
    eng.load_mlp(W, b)
    idx = oracle.uniform_indices(cat, B, seed=B)
    exp_x = oracle.gather_hashed(cat, 0x5EED, idx)
    assert_bits_equal(eng.gather_only(idx), exp_x)
    exp = oracle.mlp(exp_x, dims, W, b, mode=1)
    for _ in range(2):                       # direct, then captured / replayed
        got = eng.infer(idx)
        assert rel_err(got, exp) <= TOL, rel_err(got, exp)
    eng.close()


# ---------------------------------------------------------------- table-sharded step, every model
@pytest.mark.parametrize("model,world,B", (("medium", 2, 512), ("large", 2, 512), ("large", 4, 1024), ("small", 4, 2048)))
def test_sharded_step_every_model(model, world, B):
    """fr_shard_infer and fr_shard_infer_sliced on the medium (duplicate pad), large (377 tables, H1 = 2048) and,
    with four ranks, small model: `world` engines of one process (on as many GPUs as the box has, else all on GPU
    0), concat buffers bit-exact after the exchange, scores within 1e-3, over several steps so that both exchange
    buffers and the replayed graphs are exercised."""
    import torch
    cat = catalogue.load(model).with_row_cap(3000)
    dims = cat.layer_dims
    owner = shard.plan_owners(cat, world)
    tables = oracle.make_tables(cat, "hash", seed=31)
    W, b = oracle.make_weights(dims, seed=42)
    per = B // world
    engs = []
    for r in range(world):
        e = fleetrec.Engine(cat, device=r % _n_gpus(), max_batch=B)
        e.shard_init(r, world, owner)
        for t in cat.tables:
            e.load_table(t.id, tables[t.id])
        e.load_mlp(W, b)
        engs.append(e)
    for e in engs:
        e.shard_attach_local(engs)
    full = torch.empty((B, cat.n_tables), dtype=torch.int32).pin_memory()
    blocks = []
    for r in range(world):
        o, p = shard.slice_indices(np.zeros((B, cat.n_tables), np.int32), owner, world, r)
        blocks.append((torch.from_numpy(o.copy()).pin_memory(), torch.from_numpy(p.copy()).pin_memory()))
    outs = [torch.empty(per, dtype=torch.float32).pin_memory() for _ in engs]
    for step in range(6):
        idx = oracle.zipf_indices(cat, B, seed=300 + step)
        exp_x = oracle.gather(cat, tables, idx)
        exp = oracle.mlp(exp_x, dims, W, b, mode=1)
        full.copy_(torch.from_numpy(idx))
        for r in range(world):
            o, p = shard.slice_indices(idx, owner, world, r)
            blocks[r][0].copy_(torch.from_numpy(o))
            blocks[r][1].copy_(torch.from_numpy(p))
        sliced = step % 2 == 1
        for r, e in enumerate(engs):
            if sliced:
                e.shard_infer_sliced(blocks[r][0].numpy(), blocks[r][1].numpy(), B, outs[r].numpy())
            else:
                e.shard_infer(full.numpy(), B, outs[r].numpy())
        for e in engs:
            e.sync()
        got = np.concatenate([t.numpy() for t in outs])
        assert rel_err(got, exp) <= TOL, (step, rel_err(got, exp))
        if step == 0:     # what the peers pushed, before anything else overwrites it: tf32-rounded concat vectors
            for r, e in enumerate(engs):
                x = e.shard_read_concat(B)
                rounded = ((exp_x[r * per:(r + 1) * per].view(np.uint32).astype(np.uint64) + 0x1000) & 0xFFFFE000)
                assert np.array_equal(x.view(np.uint32), rounded.astype(np.uint32)), r
    for e in engs:
        e.close()


def test_sharded_steps_grouped_equal_single_steps():
    """fr_shard_infer_sliced_many (n steps, one copy each way) gives, bit for bit, what n fr_shard_infer_sliced calls
    give, on every rank, from one pinned buffer holding [owned blocks | replicated blocks]; direct, captured, replayed,
    with an odd n so that consecutive calls start on alternating exchange buffers."""
    import torch
    world, n, B = 2, 3, 512
    per = B // world
    cat = catalogue.load("small").with_row_cap(3000)
    dims = cat.layer_dims
    owner = shard.plan_owners(cat, world, policy="contiguous")
    tables = oracle.make_tables(cat, "hash", seed=37)
    W, b = oracle.make_weights(dims, seed=42)
    engs = []
    for r in range(world):
        e = fleetrec.Engine(cat, device=r % _n_gpus(), max_batch=B)
        e.shard_init(r, world, owner)
        for t in cat.tables:
            e.load_table(t.id, tables[t.id])
        e.load_mlp(W, b)
        engs.append(e)
    for e in engs:
        e.shard_attach_local(engs)
    idx = [oracle.zipf_indices(cat, B, seed=800 + i) for i in range(n)]
    exp = np.concatenate([oracle.mlp(oracle.gather(cat, tables, ix), dims, W, b, mode=1) for ix in idx])
    single = [torch.empty(n * per, dtype=torch.float32).pin_memory() for _ in engs]
    for i in range(n):
        for r, e in enumerate(engs):
            o, p = shard.slice_indices(idx[i], owner, world, r)
            e.shard_infer_sliced(torch.from_numpy(o).pin_memory().numpy(), torch.from_numpy(p).pin_memory().numpy(), B,
                                 single[r][i * per:(i + 1) * per].numpy())
        for e in engs:
            e.sync()
    bufs, outs = [], [torch.empty(n * per, dtype=torch.float32).pin_memory() for _ in engs]
    for r in range(world):
        o = np.concatenate([shard.slice_indices(ix, owner, world, r)[0].reshape(-1) for ix in idx])
        p = np.concatenate([shard.slice_indices(ix, owner, world, r)[1].reshape(-1) for ix in idx])
        n_o = (o.size + 3) // 4 * 4
        buf = torch.zeros(n_o + p.size, dtype=torch.int32).pin_memory()
        buf[:o.size] = torch.from_numpy(o)
        buf[n_o:] = torch.from_numpy(p)
        bufs.append((buf, buf[:o.size], buf[n_o:]))
    for rep in range(4):
        for t in outs:
            t.zero_()
        for r, e in enumerate(engs):
            e.shard_infer_sliced_many(bufs[r][1].numpy(), bufs[r][2].numpy(), n, B, outs[r].numpy())
        for e in engs:
            e.sync()
        for r in range(world):
            assert_bits_equal(outs[r].numpy(), single[r].numpy())
    got = np.concatenate([np.concatenate([outs[r][i * per:(i + 1) * per].numpy() for r in range(world)]) for i in range(n)])
    assert rel_err(got, exp) <= TOL
    for e in engs:
        gs = e.graph_stats()
        assert gs["captured"] >= 1 and gs["replayed"] >= 2, gs
        e.close()


def test_sharded_step_reports_a_missing_peer():
    """A rank whose peer never issues the step gives up after ~2 s: fr_sync and later sharded calls return
    FR_ERR_STATE instead of handing out scores computed on an incomplete concat buffer."""
    import torch
    cat = catalogue.load("small").with_row_cap(2000)
    owner = shard.plan_owners(cat, 2)
    tables = oracle.make_tables(cat, "hash", seed=5)
    W, b = oracle.make_weights(cat.layer_dims, seed=42)
    engs = []
    for r in range(2):
        e = fleetrec.Engine(cat, device=r % _n_gpus(), max_batch=256)
        e.shard_init(r, 2, owner)
        for t in cat.tables:
            e.load_table(t.id, tables[t.id])
        e.load_mlp(W, b)
        engs.append(e)
    for e in engs:
        e.shard_attach_local(engs)
    idx = torch.from_numpy(oracle.zipf_indices(cat, 256, seed=1)).pin_memory()
    out = torch.empty(128, dtype=torch.float32).pin_memory()
    engs[0].shard_infer(idx.numpy(), 256, out.numpy())        # rank 1 never does
    t0 = time.perf_counter()
    with pytest.raises(fleetrec.FleetRecError) as ei:
        engs[0].sync()
    assert ei.value.code == fleetrec.FR_ERR_STATE and 1.0 < time.perf_counter() - t0 < 8.0
    with pytest.raises(fleetrec.FleetRecError):
        engs[0].shard_infer(idx.numpy(), 256, out.numpy())
    for e in engs:
        e.close()


# ---------------------------------------------------------------- CUDA-graph cache
def test_graph_cache_steady_state_lru_and_flush():
    """A steady-state loop over a fixed set of buffers only replays; more buffer combinations than the cache
    holds (64 per worker) are handled by evicting the least recently used; fr_graph_flush forgets them."""
    import torch
    cat, tables, W, b, eng = _small_engine(max_batch=256)
    w = fleetrec.Worker(eng)
    B = 256
    idx = [torch.from_numpy(oracle.zipf_indices(cat, B, seed=i)).cuda() for i in range(70)]
    sc = torch.empty(B, dtype=torch.float32, device="cuda")
    exp = [oracle.mlp(oracle.gather(cat, tables, t.cpu().numpy()), cat.layer_dims, W, b, mode=1) for t in idx[:3]]
    for i in range(3):
        eng.infer_async(idx[i], sc, B, w)
    eng.sync(w)
    s0 = eng.graph_stats()
    assert s0 == {"replayed": 0, "captured": 2, "direct": 1}, s0     # first call of (fr_infer, 256) runs un-captured
    for rep in range(5):
        for i in range(3):
            eng.infer_async(idx[i], sc, B, w)
            eng.sync(w)
            assert rel_err(sc.cpu().numpy(), exp[i]) <= TOL
    s1 = eng.graph_stats()
    assert s1["replayed"] == 14 and s1["captured"] == 3 and s1["direct"] == 1, s1   # idx[0] captured on its 2nd use
    for t in idx:                                # 70 combinations through a 64-entry cache
        eng.infer_async(t, sc, B, w)
    eng.sync(w)
    s2 = eng.graph_stats()
    assert s2["captured"] == 3 + 67 and s2["direct"] == 1, s2
    eng.infer_async(idx[69], sc, B, w)           # most recent: still cached
    eng.infer_async(idx[0], sc, B, w)            # evicted long ago: captured again
    eng.sync(w)
    s3 = eng.graph_stats()
    assert s3["replayed"] == s2["replayed"] + 1 and s3["captured"] == s2["captured"] + 1, s3
    eng.graph_flush(w)
    eng.infer_async(idx[69], sc, B, w)
    eng.sync(w)
    assert eng.graph_stats()["captured"] == s3["captured"] + 1
    assert rel_err(sc.cpu().numpy(), oracle.mlp(oracle.gather(cat, tables, idx[69].cpu().numpy()), cat.layer_dims, W, b, mode=1)) <= TOL
    w.close()
    eng.close()


@pytest.mark.parametrize("kind", ("device", "pinned"))
def test_infer_many_equals_per_batch_infer(kind):
    """fr_infer_many (n batches, one copy each way) gives, bit for bit, what n fr_infer calls give."""
    import torch
    cat, tables, W, b, eng = _small_engine(max_batch=512)      # one batch deep: the staging buffers grow on the first call
    w = fleetrec.Worker(eng)
    n, B = 4, 512
    idx = oracle.zipf_indices(cat, n * B, seed=77)
    one = np.concatenate([eng.infer(idx[i * B:(i + 1) * B], w) for i in range(n)])
    assert rel_err(one, oracle.mlp(oracle.gather(cat, tables, idx), cat.layer_dims, W, b, mode=1)) <= TOL
    ti = torch.from_numpy(idx)
    ti = ti.cuda() if kind == "device" else ti.pin_memory()
    so = torch.empty(n * B, dtype=torch.float32, device="cuda") if kind == "device" else \
        torch.empty(n * B, dtype=torch.float32).pin_memory()
    l0 = eng.launch_count()
    for _ in range(3):                            # direct, captured, replayed
        so.zero_()
        torch.cuda.synchronize()
        eng.infer_many_async(ti if kind == "device" else ti.numpy(), so if kind == "device" else so.numpy(), n, B, w)
        eng.sync(w)
        assert_bits_equal(so.cpu().numpy(), one)
    assert eng.launch_count() - l0 == 3 * n * 4   # lookup + 3 GEMM launches per batch, no extra kernels
    with pytest.raises(fleetrec.FleetRecError):
        eng.infer_many_async(ti if kind == "device" else ti.numpy(), so if kind == "device" else so.numpy(), 1, 4 * B + 1, w)   # B > max_batch
    w.close()
    eng.close()


# ---------------------------------------------------------------- packed index rows
def test_packed_index_rows_equal_int32_rows():
    """FR_IDX_PACKED (uint16 columns for tables of at most 65536 rows behind the int32 ones): the same lookups, bit for
    bit -- gather hook, fr_infer (direct, captured, replayed from pinned buffers), fr_infer_many, with the index check
    on; the layout the library reports is the documented one; 120 instead of 188 bytes per item for the small model."""
    import torch
    cat = catalogue.load("small").with_row_cap(200000)           # tables above and below 65536 rows
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=19)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=1024)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    B, n = 1000, 3
    idx = oracle.uniform_indices(cat, n * B, seed=5)
    idx[0, :] = [t.rows - 1 for t in cat.tables]
    want_x = eng.gather_only(idx[:B])
    want = np.concatenate([eng.infer(idx[i * B:(i + 1) * B]) for i in range(n)])
    off32, wid32, rb32 = eng.index_layout()
    assert rb32 == 4 * cat.n_tables and wid32 == [4] * cat.n_tables and off32 == [4 * t for t in range(cat.n_tables)]
    eng.set_option(fleetrec.FR_OPT_INDEX_FORMAT, fleetrec.FR_IDX_PACKED)
    eng.set_option(fleetrec.FR_OPT_CHECK_INDICES, 1)
    off, wid, rb = eng.index_layout()
    big = [t.id for t in cat.tables if t.rows > 65536]
    small = [t.id for t in cat.tables if t.rows <= 65536]
    assert [wid[t] for t in big] == [4] * len(big) and [wid[t] for t in small] == [2] * len(small)
    assert [off[t] for t in big] == [4 * i for i in range(len(big))]
    assert [off[t] for t in small] == [4 * len(big) + 2 * i for i in range(len(small))]
    assert rb == (4 * len(big) + 2 * len(small) + 3) // 4 * 4 < rb32
    packed = fleetrec.pack_indices(idx, (off, wid, rb))
    assert packed.shape == (n * B, rb // 4)
    w = fleetrec.Worker(eng)
    x = np.empty((B, cat.concat_floats), np.float32)
    eng._chk(eng._L.fr_gather_only(eng._h, packed[:B].ctypes.data, B, x.ctypes.data, w._h))
    eng.sync(w)
    assert_bits_equal(x, want_x)
    pin = torch.from_numpy(packed.copy()).pin_memory()
    out = torch.zeros(n * B, dtype=torch.float32).pin_memory()
    for rep in range(3):
        out.zero_()
        for i in range(n):
            eng.infer_async(pin[i * B:(i + 1) * B].numpy(), out[i * B:(i + 1) * B].numpy(), B, w)
        eng.sync(w)
        assert_bits_equal(out.numpy(), want)
    for rep in range(3):
        out.zero_()
        eng.infer_many_async(pin.numpy(), out.numpy(), n, B, w)
        eng.sync(w)
        assert_bits_equal(out.numpy(), want)
    bad = idx[:B].copy()
    bad[7, small[3]] = cat.tables[small[3]].rows          # representable in uint16, outside the table
    eng.infer_async(fleetrec.pack_indices(bad, (off, wid, rb)), np.empty(B, np.float32), B, w)
    with pytest.raises(fleetrec.FleetRecError) as ei:
        eng.sync(w)
    assert ei.value.code == fleetrec.FR_ERR_INVALID
    eng.set_option(fleetrec.FR_OPT_INDEX_FORMAT, fleetrec.FR_IDX_I32)
    assert_bits_equal(eng.infer(idx[:B]), want[:B])
    w.close()
    eng.close()


def test_packed_index_rows_through_the_sliced_sharded_step():
    """The column-sliced blocks of fr_shard_infer_sliced(_many) in FR_IDX_PACKED: same scores as with int32 blocks."""
    import torch
    world, B, n = 2, 512, 2
    per = B // world
    cat = catalogue.load("small").with_row_cap(100000)
    dims = cat.layer_dims
    owner = shard.plan_owners(cat, world, policy="contiguous")
    tables = oracle.make_tables(cat, "hash", seed=43)
    W, b = oracle.make_weights(dims, seed=42)
    engs = []
    for r in range(world):
        e = fleetrec.Engine(cat, device=r % _n_gpus(), max_batch=B)
        e.shard_init(r, world, owner)
        for t in cat.tables:
            e.load_table(t.id, tables[t.id])
        e.load_mlp(W, b)
        engs.append(e)
    for e in engs:
        e.shard_attach_local(engs)
    idx = [oracle.zipf_indices(cat, B, seed=900 + i) for i in range(n)]
    exp = np.concatenate([oracle.mlp(oracle.gather(cat, tables, ix), dims, W, b, mode=1) for ix in idx])
    res = {}
    for fmt in (fleetrec.FR_IDX_I32, fleetrec.FR_IDX_PACKED):
        bufs, outs = [], [torch.zeros(n * per, dtype=torch.float32).pin_memory() for _ in engs]
        for r, e in enumerate(engs):
            e.set_option(fleetrec.FR_OPT_INDEX_FORMAT, fmt)
            lo, lr = e.index_layout(0), e.index_layout(1)
            o = np.concatenate([fleetrec.pack_indices(shard.slice_indices(ix, owner, world, r)[0], lo).reshape(-1) for ix in idx])
            p = np.concatenate([fleetrec.pack_indices(shard.slice_indices(ix, owner, world, r)[1], lr).reshape(-1) for ix in idx])
            n_o = (o.size + 3) // 4 * 4
            buf = torch.zeros(n_o + p.size, dtype=torch.int32).pin_memory()
            buf[:o.size] = torch.from_numpy(o)
            buf[n_o:] = torch.from_numpy(p)
            bufs.append((buf, buf[:o.size], buf[n_o:]))
        for rep in range(3):
            for r, e in enumerate(engs):
                e.shard_infer_sliced_many(bufs[r][1].numpy(), bufs[r][2].numpy(), n, B, outs[r].numpy())
            for e in engs:
                e.sync()
        res[fmt] = np.concatenate([np.concatenate([outs[r][i * per:(i + 1) * per].numpy() for r in range(world)]) for i in range(n)])
    assert_bits_equal(res[fleetrec.FR_IDX_PACKED], res[fleetrec.FR_IDX_I32])
    assert rel_err(res[fleetrec.FR_IDX_PACKED], exp) <= TOL
    for e in engs:
        e.close()


# ---------------------------------------------------------------- index checking
def test_check_indices_reports_and_reads_row_zero():
    cat, tables, W, b, eng = _small_engine(max_batch=256)
    idx = oracle.zipf_indices(cat, 256, seed=8)
    good = eng.infer(idx)
    bad = idx.copy()
    bad[17, 30] = cat.tables[30].rows            # one past the end
    bad[200, 3] = -5
    eng.set_option(fleetrec.FR_OPT_CHECK_INDICES, 1)
    assert_bits_equal(eng.infer(idx), good)       # in-range batches are unaffected
    with pytest.raises(fleetrec.FleetRecError) as ei:
        eng.infer(bad)
    assert ei.value.code == fleetrec.FR_ERR_INVALID and "out-of-range" in str(ei.value)
    with pytest.raises(fleetrec.FleetRecError):
        eng.gather_only(bad)
    # the offenders read row 0: same concat as the batch with those two indices set to 0
    fixed = bad.copy()
    fixed[17, 30] = 0
    fixed[200, 3] = 0
    sc = np.empty(256, np.float32)
    eng.infer_async(bad, sc)
    with pytest.raises(fleetrec.FleetRecError):
        eng.sync()
    assert_bits_equal(sc, eng.infer(fixed))
    eng.set_option(fleetrec.FR_OPT_CHECK_INDICES, 0)
    assert_bits_equal(eng.infer(idx), good)
    eng.close()


def test_ingest_rejects_out_of_range_indices():
    """FR_INGEST_INDICES blocks come off a socket: a block with an index outside its table is refused on the host
    (never reaches the lookup kernel), the connection ends with FR_ERR_INVALID."""
    import socket
    cat, tables, W, b, eng = _small_engine(max_batch=64)
    port = 23000 + os.getpid() % 2000
    ing = fleetrec.Ingest(eng, port, 1, 64, payload="indices", max_batches_per_conn=4)
    idx = oracle.zipf_indices(cat, 64, seed=4)
    bad = idx.copy()
    bad[5, 46] = 1 << 30
    with socket.create_connection(("127.0.0.1", port)) as s:
        s.sendall(idx.tobytes())
        s.sendall(bad.tobytes())
        s.sendall(idx.tobytes())
    with pytest.raises(fleetrec.FleetRecError) as ei:
        ing.wait()
    assert ei.value.code == fleetrec.FR_ERR_INVALID and "outside its" in str(ei.value)
    exp = oracle.mlp(oracle.gather(cat, tables, idx), cat.layer_dims, W, b, mode=1)
    assert rel_err(ing.scores[0, 0], exp) <= TOL and np.all(ing.scores[0, 1] == 0)
    ing.close()
    eng.close()


# ---------------------------------------------------------------- batcher: split requests
def test_batcher_split_request_ticket_waits_for_every_part():
    """A request larger than max_batch is split over batches that run concurrently on several workers and finish
    in any order; its (single) ticket must not complete before ALL parts are scored: wait on that ticket alone and
    check the scores at once, many times."""
    cat, tables, W, b, eng = _small_engine(max_batch=512)
    bat = fleetrec.Batcher(eng, max_batch=512, max_delay_us=200, n_workers=4)
    for rep in range(20):
        n = 512 * 3 + 37 + rep                   # three full batches and a short tail that finishes first
        idx = oracle.zipf_indices(cat, n, seed=500 + rep)
        out = np.full(n, np.nan, np.float32)
        bat.wait(bat.submit(idx, out))
        assert not np.isnan(out).any(), (rep, int(np.isnan(out).sum()))
        if rep % 5 == 0:
            assert rel_err(out, oracle.mlp(oracle.gather(cat, tables, idx), cat.layer_dims, W, b, mode=1)) <= TOL
    bat.close()
    eng.close()


# ---------------------------------------------------------------- fr_mlp_only rounds like fr_infer
def test_mlp_only_rounds_operands_like_infer():
    """TF32: fr_mlp_only on the unrounded fp32 concat vectors gives the bits fr_infer gives (both round to nearest,
    ties away, before the tensor cores would truncate)."""
    cat, tables, W, b, eng = _small_engine(max_batch=1024)
    idx = oracle.zipf_indices(cat, 1000, seed=12)
    x = eng.gather_only(idx)
    assert_bits_equal(eng.mlp_only(x), eng.infer(idx))
    eng.close()


# ---------------------------------------------------------------- guarded fp16 operands
@pytest.mark.parametrize("model", ("small", "medium"))
@pytest.mark.parametrize("B", (1, 333, 2048, 4099))
def test_f16_guarded_operands_vs_oracle(model, B):
    """FR_F16_GUARDED on data the analysis can bound (hash fill: |x| < 1, N(0, 1/in) weights): the engine picks fp16
    operands, its bounds agree with the numpy restatement, scores stay within 1e-3 of the fp32 oracle, the lookup
    hook stays fp32 and bit-exact, fr_mlp_only stays TF32."""
    cat = catalogue.load(model).with_row_cap(20000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=41)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=max(B, 256))
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    idx = oracle.zipf_indices(cat, B, seed=B)
    x = oracle.gather(cat, tables, idx)
    exp = oracle.mlp(x, dims, W, b, mode=1)
    tf32 = eng.infer(idx)
    tf32_mlp = eng.mlp_only(x)
    eng.set_option(fleetrec.FR_OPT_F16_OPERANDS, fleetrec.FR_F16_GUARDED)
    active, bounds = eng.f16_report()
    assert active, bounds
    safe, rep = f16_bound_ref.f16_safe(cat, [float(np.abs(t).max()) for t in tables], W, b)
    assert safe
    assert bounds[0] == pytest.approx(rep["x"], rel=1e-6)
    assert rep["h1"] <= bounds[1] <= rep["h1"] * 1.01 and rep["h2"] <= bounds[2] <= rep["h2"] * 1.03
    assert bounds[3] == pytest.approx(min(float(np.abs(t[t != 0]).min()) for t in tables), rel=1e-6)
    assert_bits_equal(eng.gather_only(idx), x)
    for _ in range(3):
        got = eng.infer(idx)
    assert rel_err(got, exp) <= TOL, rel_err(got, exp)
    assert_bits_equal(eng.mlp_only(x), tf32_mlp)
    eng.set_option(fleetrec.FR_OPT_F16_OPERANDS, fleetrec.FR_F16_OFF)
    assert_bits_equal(eng.infer(idx), tf32)
    eng.close()


def test_f16_guard_refuses_what_it_cannot_prove():
    """Adversarial cases: each must fall back to TF32 and give the TF32 answer.
    (a) the reference's all-ones known answer (352 -> 360448 at layer 1: far beyond 65504) stays EXACT;
    (b) one 7e4 entry in one table; (c) a table of tiny (fp16-subnormal) values; (d) a weight matrix whose mass is
    mostly fp16-subnormal; (e) a NaN in a table."""
    cat = catalogue.load("small").with_row_cap(4000)
    dims = cat.layer_dims
    KAT = 47244640256.0
    # (a)
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, max_batch=64)
    eng.fill_reference()
    eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
    eng.set_option(fleetrec.FR_OPT_F16_OPERANDS, fleetrec.FR_F16_GUARDED)
    active, bounds = eng.f16_report()
    assert not active and bounds[1] == pytest.approx(352.0 * 1.001, rel=1e-5) and bounds[2] > 65504
    got = eng.infer(oracle.idx_reference(32, cat.n_tables))
    assert set(np.unique(got)) == {0.0, np.float32(KAT)}
    eng.close()
    # (b) .. (e)
    tables = oracle.make_tables(cat, "hash", seed=6)
    W, b = oracle.make_weights(dims, seed=42)
    idx = oracle.zipf_indices(cat, 300, seed=2)
    cases = {}
    t_b = [t.copy() for t in tables]
    t_b[20][123, 5] = 7e4
    cases["one 7e4 table entry"] = (t_b, W)
    t_c = [t.copy() for t in tables]
    t_c[3] *= np.float32(1e-6)
    cases["fp16-subnormal table"] = (t_c, W)
    W_d = [w.copy() for w in W]
    W_d[1] *= np.float32(1e-5)
    cases["fp16-subnormal weights"] = (tables, W_d)
    t_e = [t.copy() for t in tables]
    t_e[40][7, 1] = np.nan
    cases["NaN in a table"] = (t_e, W)
    for name, (tb, Wc) in cases.items():
        eng = fleetrec.Engine(cat, max_batch=512)
        eng.load_tables(tb)
        eng.load_mlp(Wc, b)
        tf32 = eng.infer(idx)
        eng.set_option(fleetrec.FR_OPT_F16_OPERANDS, fleetrec.FR_F16_GUARDED)
        active, bounds = eng.f16_report()
        assert not active, (name, bounds)
        assert np.array_equal(eng.infer(idx).view(np.uint32), tf32.view(np.uint32)), name
        # the decision follows the data: restoring the offending table / weights re-enables fp16
        eng.load_tables(tables)
        eng.load_mlp(W, b)
        assert eng.f16_report()[0], name
        assert rel_err(eng.infer(idx), oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)) <= TOL
        eng.close()


# ---------------------------------------------------------------- tile hints
@pytest.mark.parametrize("model", ("small", "medium"))
def test_tile_hints_change_tiles_not_results(model):
    """FR_OPT_TILE_HINT picks narrow (latency) or wide (throughput) tcgen05 tiles; both stay within tolerance at the
    batch sizes where the choice differs, and the latency tiles spread a lone batch over more SMs."""
    import ctypes as C
    from fleetrec import _capi
    cat = catalogue.load(model).with_row_cap(20000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=9)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=4096)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    raw = C.CDLL(_capi.LIB_PATH)
    raw.frdbg_layer_ctas.argtypes = [C.c_void_p, C.c_int]
    for B in (1024, 2048, 4096):
        idx = oracle.zipf_indices(cat, B, seed=B)
        exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
        ctas = {}
        for hint in (fleetrec.FR_HINT_LATENCY, fleetrec.FR_HINT_THROUGHPUT):
            eng.set_option(fleetrec.FR_OPT_TILE_HINT, hint)
            got = eng.infer(idx)
            assert rel_err(got, exp) <= TOL, (B, hint, rel_err(got, exp))
            ctas[hint] = [int(raw.frdbg_layer_ctas(eng._h, k)) for k in range(3)]
        assert ctas[fleetrec.FR_HINT_LATENCY][0] >= ctas[fleetrec.FR_HINT_THROUGHPUT][0], (B, ctas)
        assert ctas[fleetrec.FR_HINT_LATENCY][1] >= ctas[fleetrec.FR_HINT_THROUGHPUT][1], (B, ctas)
    eng.close()
