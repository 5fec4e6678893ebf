"""Table-sharded multi-GPU host logic: one process per GPU.

The reference splits its large model by table over FPGA0 / FPGA1 / CPU0, each
sending its slice of every item's concat vector to the GPU server over TCP
(GPU/final_network_cublasLt_3_nodes_no_FIFO_scatter/constant.h:25-27,
cuda_server.c:513-587).  Here every rank owns a table subset, gathers it for the
GLOBAL batch and stores the pieces straight into the concat buffer of the rank
that owns each item (NVLink peer stores from the lookup kernel); the MLP is
batch-parallel.  torch.distributed is only the plumbing that carries the 64-byte
CUDA-IPC handles and the end-of-run barrier.
"""
import numpy as np


def plan_owners(model, world, replicate_tiers=("PLRAM",), replicate_max_bytes=16 << 20, replicate_below_bytes=0,
                policy="balanced"):
    """owner[t] for every table: -1 = replicated on every rank, else the owning rank.

    On-chip-class tables (the reference's PLRAM tier: <= 10 000 rows, L2-resident
    here) are replicated, which removes their floats from the exchange.
    `replicate_below_bytes` > 0 additionally replicates ANY table smaller than that, whatever its
    tier: a table that fits L2 many times over costs next to nothing to replicate, and every
    replicated table leaves the exchange.  The rest are owned by one rank each:

    policy "balanced"    greedily by descending bytes-per-item traffic then bytes, to the currently lightest
                         rank (ties: lowest rank): traffic within one widest row, capacity spread.
    policy "contiguous"  in concat (wire) order, cut into `world` runs of about equal floats: every rank pushes ONE
                         contiguous run of every item's vector (long NVLink stores, every lane of the push kernel
                         busy) -- traffic stays balanced, resident bytes are whatever the runs hold.

    Deterministic, so every rank computes the same plan without communicating."""
    owner = [None] * model.n_tables
    if world == 1:
        return [0] * model.n_tables

    def replicated(t):
        return (t.tier in replicate_tiers and t.rows * t.dim * 4 <= replicate_max_bytes) or \
            t.rows * t.dim * 4 < replicate_below_bytes
    if policy == "contiguous":
        order = []                                      # owned tables by first appearance in the concat vector
        for s in sorted(model.segments, key=lambda s: s.dst):
            t = model.tables[s.table]
            if replicated(t):
                owner[t.id] = -1
            elif t.id not in order:
                order.append(t.id)
        total = sum(model.tables[t].dim for t in order)
        acc, r = 0, 0
        for t in order:
            # move on to the next rank once this one holds its share (never leave a later rank empty-handed)
            if r < world - 1 and acc >= (r + 1) * total / world:
                r += 1
            owner[t] = r
            acc += model.tables[t].dim
        for t in model.tables:                          # tables no segment reads (none in the catalogues)
            if owner[t.id] is None:
                owner[t.id] = -1 if replicated(t) else 0
        return owner
    if policy != "balanced":
        raise ValueError(policy)
    load = [0.0] * world          # gathered bytes per item (traffic balance)
    size = [0] * world            # resident bytes (capacity balance)
    order = sorted(model.tables, key=lambda t: (-t.dim, -t.rows * t.dim, t.id))
    for t in order:
        if replicated(t):
            owner[t.id] = -1
            continue
        r = min(range(world), key=lambda k: (load[k], size[k], k))
        owner[t.id] = r
        load[r] += t.dim * 4
        size[r] += t.rows * t.dim * 4
    return owner


def owned_floats(model, owner, rank):
    """floats of one item's concat vector this rank produces (pushed + replicated-local)."""
    pushed = sum(s.len for s in model.segments if owner[s.table] == rank)
    local = sum(s.len for s in model.segments if owner[s.table] == -1)
    return pushed, local


def item_range(B_global, world, rank):
    per = B_global // world
    return rank * per, (rank + 1) * per


def rank_tables(owner, rank):
    """(owned, replicated) table ids of `rank`, ascending -- the column order of its sliced index blocks
    (what fr_shard_tables reports)."""
    return ([t for t, o in enumerate(owner) if o == rank], [t for t, o in enumerate(owner) if o == -1])


def index_layout(rows, packed=True):
    """(byte_offsets, widths, row_bytes) of one index row over tables of `rows` rows each, in the engine's transport
    formats -- what fr_index_layout reports for the same column list (include/fleetrec.h, FR_OPT_INDEX_FORMAT).
    packed: int32 columns of the tables above 65536 rows first, in list order, then uint16 columns, row padded to 4."""
    rows = list(rows)
    if not packed:
        return [4 * i for i in range(len(rows))], [4] * len(rows), 4 * len(rows)
    off, wid, pos = [0] * len(rows), [4 if r > 65536 else 2 for r in rows], 0
    for w in (4, 2):
        for i, wi in enumerate(wid):
            if wi == w:
                off[i], pos = pos, pos + w
    return off, wid, (pos + 3) // 4 * 4


def slice_indices(idx, owner, world, rank):
    """Column-slice a global index batch [B][T] for one rank, as the reference's index source does per FPGA
    (each device is sent only its own tables' indices): returns (idx_owned [B][n_owned] over ALL items,
    idx_repl [B/world][n_repl] over the rank's own items), both C-contiguous int32."""
    owned, repl = rank_tables(owner, rank)
    b0, b1 = item_range(idx.shape[0], world, rank)
    return (np.ascontiguousarray(idx[:, owned], dtype=np.int32),
            np.ascontiguousarray(idx[b0:b1, repl], dtype=np.int32))


def exchange_handles(engine, dist, device=None):
    """All-gather every rank's 64-byte CUDA-IPC handle; returns world*64 bytes."""
    import torch
    mine = torch.frombuffer(bytearray(engine.shard_export()), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    world = dist.get_world_size()
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def simulate_sharded_gather(model, owner, world, per_rank_gather, idx):
    """Host model of the exchange: rank r contributes the pieces of the tables it owns for
    ALL items, every rank contributes replicated pieces for its own items.  Returns the
    list of per-rank concat buffers [B/world][D].  `per_rank_gather(rank, idx)` must
    return the full-width [B][D] gather as rank `rank` would compute it (zeros where it
    has no table).  Used by the CPU tests to check the plan covers every float once."""
    B = idx.shape[0]
    per = B // world
    out = [np.zeros((per, model.concat_floats), np.float32) for _ in range(world)]
    hits = [np.zeros((per, model.concat_floats), np.int32) for _ in range(world)]
    for r in range(world):
        full = per_rank_gather(r, idx)
        for s in model.segments:
            o = owner[s.table]
            if o == r:
                for dst in range(world):
                    b0, b1 = item_range(B, world, dst)
                    out[dst][:, s.dst:s.dst + s.len] = full[b0:b1, s.dst:s.dst + s.len]
                    hits[dst][:, s.dst:s.dst + s.len] += 1
            elif o == -1:
                b0, b1 = item_range(B, world, r)
                out[r][:, s.dst:s.dst + s.len] = full[b0:b1, s.dst:s.dst + s.len]
                hits[r][:, s.dst:s.dst + s.len] += 1
    return out, hits
