"""Multi-GPU path: host logic on CPU (world-size-2 gloo processes) and the device path (peer-store push +
device-side step flags).  The device tests run their ranks as separate engines of ONE process: on separate
GPUs when the box has them, otherwise all on GPU 0 (fr_shard_attach_local allows it: same kernels, same
flags, same exchange buffers, only the peer stores stay on-device), so a 1-GPU box covers them too."""
import os
import socket

import numpy as np
import pytest

import fleetrec
from fleetrec import catalogue, shard
from oracle import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("model", ("small", "medium", "large"))
@pytest.mark.parametrize("world", (2, 4, 8))
def test_plan_owners_covers_every_float_once(model, world):
    cat = catalogue.load(model)
    owner = shard.plan_owners(cat, world)
    assert len(owner) == cat.n_tables and all(-1 <= o < world for o in owner)
    assert all((o == -1) == (t.tier == "PLRAM") for o, t in zip(owner, cat.tables))
    pushed = [shard.owned_floats(cat, owner, r)[0] for r in range(world)]
    local = shard.owned_floats(cat, owner, 0)[1]
    assert sum(pushed) + local == cat.concat_floats
    assert max(pushed) - min(pushed) <= 64          # traffic balance: within one widest row
    assert shard.plan_owners(cat, world) == owner   # deterministic on every rank


@pytest.mark.parametrize("model", ("small", "medium", "large"))
@pytest.mark.parametrize("world", (2, 4, 8))
def test_contiguous_plan_cuts_the_wire_order_into_equal_runs(model, world):
    """policy="contiguous": every rank owns ONE run of consecutive owned tables in concat (wire) order, the runs'
    float counts differ by at most one widest row, on-chip-class tables stay replicated, every float is produced
    exactly once, and the plan is deterministic."""
    cat = catalogue.load(model)
    owner = shard.plan_owners(cat, world, policy="contiguous")
    assert owner == shard.plan_owners(cat, world, policy="contiguous")
    assert all((o == -1) == (t.tier == "PLRAM") for o, t in zip(owner, cat.tables))
    seq = []                                   # owners of the owned tables in order of first appearance on the wire
    for s in sorted(cat.segments, key=lambda s: s.dst):
        if owner[s.table] >= 0 and (not seq or seq[-1][0] != s.table) and s.table not in [t for t, _ in seq]:
            seq.append((s.table, owner[s.table]))
    ranks = [o for _, o in seq]
    assert ranks == sorted(ranks) and set(ranks) == set(range(world))      # one run per rank, in rank order, none empty
    pushed = [shard.owned_floats(cat, owner, r)[0] for r in range(world)]
    assert sum(pushed) + shard.owned_floats(cat, owner, 0)[1] == cat.concat_floats
    assert max(pushed) - min(pushed) <= 2 * max(t.dim for t in cat.tables)
    with pytest.raises(ValueError):
        shard.plan_owners(cat, world, policy="nope")


@pytest.mark.parametrize("model", ("small", "large"))
@pytest.mark.parametrize("world", (2, 8))
def test_column_sliced_index_blocks_cover_the_batch(model, world):
    """fr_shard_infer_sliced's input: rank r is given [B][owned tables] over all items and [B/world][replicated
    tables] over its own items.  Together the blocks of all ranks hold every index the un-sliced batch holds that
    any rank reads, each owned column exactly once, and far fewer indices per rank than B * T."""
    cat = catalogue.load(model)
    owner = shard.plan_owners(cat, world)
    B = 16 * world
    idx = oracle.uniform_indices(cat.with_row_cap(1000), B, seed=3)
    seen = np.zeros(cat.n_tables, np.int32)
    for r in range(world):
        owned, repl = shard.rank_tables(owner, r)
        io, ir = shard.slice_indices(idx, owner, world, r)
        assert io.shape == (B, len(owned)) and ir.shape == (B // world, len(repl))
        assert io.flags.c_contiguous and ir.flags.c_contiguous and io.dtype == np.int32
        b0, b1 = shard.item_range(B, world, r)
        assert np.array_equal(io, idx[:, owned]) and np.array_equal(ir, idx[b0:b1, repl])
        seen[owned] += 1
        assert io.size + ir.size < idx.size
        # the same blocks in the packed transport format: every column decodes to the int32 block's
        capped = cat.with_row_cap(1000)
        for blk, tabs in ((io, owned), (ir, repl)):
            off, wid, rb = shard.index_layout([capped.tables[t].rows if t % 3 else 10 ** 6 for t in tabs])
            pk = fleetrec.pack_indices(blk, (off, wid, rb)).view(np.uint8).reshape(blk.shape[0], rb)
            assert rb % 4 == 0 and rb <= 4 * len(tabs)
            for c, (o, w) in enumerate(zip(off, wid)):
                col = pk[:, o:o + w].copy().view("<u2" if w == 2 else "<i4").reshape(-1)
                assert np.array_equal(col.astype(np.int64), blk[:, c].astype(np.int64))
    assert all(seen[t] == (0 if owner[t] == -1 else 1) for t in range(cat.n_tables))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cat = catalogue.load("small").with_row_cap(500)
        owner = shard.plan_owners(cat, world)
        # every rank derives the same plan without talking
        t = torch.tensor(owner, dtype=torch.int32)
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)

        # handle exchange plumbing (64-byte blobs, one per rank, rank order preserved)
        class FakeEngine:
            def shard_export(self):
                return bytes([rank]) * 64
        blob = shard.exchange_handles(FakeEngine(), dist)
        assert len(blob) == 64 * world and all(blob[64 * r] == r for r in range(world))

        # each rank holds only its own + replicated tables and gathers them for the GLOBAL batch
        B = 64
        idx = oracle.uniform_indices(cat, B, seed=7)            # same seed => same global batch everywhere
        tables = [oracle.fill_hash(9, tb.id, tb.rows, tb.dim) if owner[tb.id] in (-1, rank)
                  else np.zeros((tb.rows, tb.dim), np.float32) for tb in cat.tables]
        mine = oracle.gather(cat, tables, idx)
        # emulate the push all-to-all with gloo: every rank publishes its full-width gather and
        # each destination keeps, per piece, the rows the owning rank produced for its items
        per = B // world
        b0, b1 = shard.item_range(B, world, rank)
        allg = [torch.zeros(B, cat.concat_floats) for _ in range(world)]
        dist.all_gather(allg, torch.from_numpy(mine))
        recv = [g[b0:b1] for g in allg]
        out = np.zeros((per, cat.concat_floats), np.float32)
        for s in cat.segments:
            o = owner[s.table]
            src = rank if o == -1 else o                          # replicated pieces come from myself
            out[:, s.dst:s.dst + s.len] = recv[src].numpy()[:, s.dst:s.dst + s.len]
        full_tables = [oracle.fill_hash(9, tb.id, tb.rows, tb.dim) for tb in cat.tables]
        exp = oracle.gather(cat, full_tables, idx)[b0:b1]
        assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_simulated_exchange_hits_every_float_once():
    cat = catalogue.load("medium").with_row_cap(300)
    world = 4
    owner = shard.plan_owners(cat, world)
    full_tables = oracle.make_tables(cat, "hash", seed=2)
    idx = oracle.uniform_indices(cat, 32, seed=1)

    def per_rank(rank, idx):
        tabs = [t if owner[i] in (-1, rank) else np.zeros_like(t) for i, t in enumerate(full_tables)]
        return oracle.gather(cat, tabs, idx)
    out, hits = shard.simulate_sharded_gather(cat, owner, world, per_rank, idx)
    exp = oracle.gather(cat, full_tables, idx)
    for r in range(world):
        assert np.all(hits[r] == 1)            # incl. the duplicate pad piece: written once
        b0, b1 = shard.item_range(32, world, r)
        assert np.array_equal(out[r].view(np.uint32), exp[b0:b1].view(np.uint32))


# --------------------------------------------------------------------------- device path
def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", (1, 2))
def test_sharded_device_path_single_process(world):
    """Two engines of one process (fr_shard_attach_local): peer-store push, then the one-call step with the
    device-side flags, against the oracle."""
    import torch
    cat = catalogue.load("small").with_row_cap(5000)
    dims = cat.layer_dims
    owner = shard.plan_owners(cat, world)
    tables = oracle.make_tables(cat, "hash", seed=21)
    W, b = oracle.make_weights(dims, seed=42)
    B = 256
    engs = []
    for r in range(world):
        e = fleetrec.Engine(cat, device=r % _n_gpus(), max_batch=B)
        e.shard_init(r, world, owner)
        for t in cat.tables:
            e.load_table(t.id, tables[t.id])         # no-op for tables this rank does not own
        e.load_mlp(W, b)
        engs.append(e)
    assert sum(e.table_bytes() for e in engs) == cat.table_bytes() + (world - 1) * sum(
        t.rows * t.dim * 4 for t in cat.tables if owner[t.id] == -1)
    for e in engs:
        e.shard_attach_local(engs)
    idx = oracle.zipf_indices(cat, B, seed=5)
    exp_x = oracle.gather(cat, tables, idx)
    exp_s = oracle.mlp(exp_x, dims, W, b, mode=1)
    per = B // world
    # two-phase form: host barrier between push and MLP; FP32 so the exchanged bytes are unrounded
    for e in engs:
        e.set_precision(fleetrec.FR_PREC_FP32)
        e.shard_gather_push(idx)
    for e in engs:
        e.sync()
    for r, e in enumerate(engs):
        got = e.shard_read_concat(B)
        assert np.array_equal(got.view(np.uint32), exp_x[r * per:(r + 1) * per].view(np.uint32))
    # one-call form, several steps with fresh indices (exercises both buffer parities)
    for e in engs:
        e.set_precision(fleetrec.FR_PREC_TF32)
    for step in range(5):
        idx = oracle.zipf_indices(cat, B, seed=50 + step)
        exp_s = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
        # pinned buffers: a pageable D2H would block the host inside rank 0's call while its
        # device-side wait needs rank 1's kernels, which this single thread has not enqueued yet
        idx_p = torch.from_numpy(idx).pin_memory()
        outs = [torch.empty(per, dtype=torch.float32).pin_memory() for _ in engs]
        for e, o in zip(engs, outs):
            e.shard_infer(idx_p.numpy(), B, o.numpy())
        for e in engs:
            e.sync()
        got = np.concatenate([o.numpy() for o in outs])
        err = float(np.max(np.abs(got - exp_s) / np.maximum(np.abs(exp_s), 1e-6)))
        assert err <= 1e-3, (step, err)
    if world > 1:
        with pytest.raises(fleetrec.FleetRecError):      # un-sharded entry points refuse a sharded engine
            engs[0].gather_only(idx)
    for e in engs:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", (1, 2))
def test_sharded_multi_worker_graph_replay(world):
    """Several sharded steps in flight: every worker stream owns an exchange slot (own concat
    buffers, flags and device-side step counter), and a step seen before on the same buffers is
    replayed as a CUDA graph (one per buffer parity).  New index CONTENTS in the same pinned
    buffers must give new, correct scores on every rank, for both workers, over many steps."""
    import torch
    cat = catalogue.load("small").with_row_cap(5000)
    dims = cat.layer_dims
    owner = shard.plan_owners(cat, world)
    tables = oracle.make_tables(cat, "hash", seed=23)
    W, b = oracle.make_weights(dims, seed=42)
    B, per, n_workers = 512, 512 // world, 2
    engs, workers = [], []
    for r in range(world):
        e = fleetrec.Engine(cat, device=r % _n_gpus(), max_batch=B)
        e.shard_init(r, world, owner)
        for t in cat.tables:
            e.load_table(t.id, tables[t.id])
        e.load_mlp(W, b)
        engs.append(e)
        workers.append([fleetrec.Worker(e) for _ in range(n_workers)])
    for e in engs:
        e.shard_attach_local(engs)
    idx_p = [torch.empty((B, cat.n_tables), dtype=torch.int32).pin_memory() for _ in range(n_workers)]
    outs = [[torch.empty(per, dtype=torch.float32).pin_memory() for _ in range(n_workers)] for _ in engs]
    l0 = [e.launch_count() for e in engs]
    for step in range(7):
        exp = []
        for w in range(n_workers):
            idx = oracle.zipf_indices(cat, B, seed=900 + 10 * step + w)
            idx_p[w].copy_(torch.from_numpy(idx))
            exp.append(oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1))
        for w in range(n_workers):              # both workers' steps are in flight on every rank
            for r, e in enumerate(engs):
                e.shard_infer(idx_p[w].numpy(), B, outs[r][w].numpy(), workers[r][w])
        for r, e in enumerate(engs):
            for w in range(n_workers):
                e.sync(workers[r][w])
        for w in range(n_workers):
            got = np.concatenate([outs[r][w].numpy() for r in range(world)])
            err = float(np.max(np.abs(got - exp[w]) / np.maximum(np.abs(exp[w]), 1e-6)))
            assert err <= 1e-3, (step, w, err)
    # same kernels per step whether launched directly or replayed from the graph
    per_step = (engs[0].launch_count() - l0[0]) / (7 * n_workers)
    # push (+ replicated lookup) + flag kernel (publish, wait) + 3 GEMM launches
    assert per_step == (6 if world > 1 else 5), per_step
    for e in engs:   # every step after a worker's first replayed a graph (both buffer parities captured together)
        gs = e.graph_stats()
        assert gs["direct"] <= 1 and gs["captured"] <= n_workers and gs["replayed"] >= 7 * n_workers - n_workers - 1, gs
    for r, e in enumerate(engs):
        for w in workers[r]:
            w.close()
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", (1, 2))
def test_sharded_step_from_column_sliced_indices_matches_full_rows(world):
    """fr_shard_infer_sliced (each rank uploads only [B][owned tables] + [B/world][replicated tables]) gives the
    bits fr_shard_infer gives from the full [B][T] rows: same lookups, same exchange, same MLP; direct and
    graph-replayed, from pinned host blocks whose CONTENTS change between steps."""
    import torch
    cat = catalogue.load("small").with_row_cap(5000)
    dims = cat.layer_dims
    owner = shard.plan_owners(cat, world)
    tables = oracle.make_tables(cat, "hash", seed=29)
    W, b = oracle.make_weights(dims, seed=42)
    B, per = 1024, 1024 // world
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
    for r, e in enumerate(engs):
        assert (e.shard_tables(0), e.shard_tables(1)) == shard.rank_tables(owner, r)
    full = torch.empty((B, cat.n_tables), dtype=torch.int32).pin_memory()
    blocks = []
    for r in range(world):
        o, p = shard.slice_indices(np.zeros((B, cat.n_tables), np.int32), owner, world, r)
        blocks.append((torch.from_numpy(o.copy()).pin_memory(), torch.from_numpy(p.copy()).pin_memory()))
    out_full = [torch.empty(per, dtype=torch.float32).pin_memory() for _ in engs]
    out_sl = [torch.empty(per, dtype=torch.float32).pin_memory() for _ in engs]
    for step in range(5):
        idx = oracle.zipf_indices(cat, B, seed=700 + step)
        full.copy_(torch.from_numpy(idx))
        for r in range(world):
            o, p = shard.slice_indices(idx, owner, world, r)
            blocks[r][0].copy_(torch.from_numpy(o))
            blocks[r][1].copy_(torch.from_numpy(p))
        for r, e in enumerate(engs):
            e.shard_infer(full.numpy(), B, out_full[r].numpy())
        for e in engs:
            e.sync()
        for r, e in enumerate(engs):
            e.shard_infer_sliced(blocks[r][0].numpy(), blocks[r][1].numpy(), B, out_sl[r].numpy())
        for e in engs:
            e.sync()
        exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
        got = np.concatenate([t.numpy() for t in out_sl])
        assert np.array_equal(got.view(np.uint32), np.concatenate([t.numpy() for t in out_full]).view(np.uint32)), step
        assert float(np.max(np.abs(got - exp) / np.maximum(np.abs(exp), 1e-6))) <= 1e-3, step
    for e in engs:
        e.close()
