#!/usr/bin/env python3
"""bench.py -- inferences/s of the FleetRec hot path (lookup+concat -> MLP) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on host cores

Workload: batches of 2048 items of the small (47-table, 352-1024-512-256-1) model -- BASELINE.json
configs[1] -- through gather+MLP, dealt to `--streams` worker streams (the reference's THREAD_NUM
workers, cuda_server.c:554-556) so that several batches are in flight.  A STEP is a fixed quantum of
that stream: `--rounds` batches on every worker (default 64 x 12 = 768 batches = 1.57 M items), so a
step is long against the pipeline's ramp and drain and ms_per_step does not depend on --steps; the
timed region is bracketed by barrier + synchronize and timed with CUDA events that fork from / join
into the worker streams.

  value  indices already resident in HBM, scores left in HBM (fr_infer per batch)
  e2e    the public call on pinned HOST index and score buffers (fr_infer_many: `--group` batches per
         call, one H2D copy of their indices and one D2H copy of their scores, both inside the
         timed region)
  roofline      the step's slowest kernel, timed alone with CUDA events on its stream
  cpu_baseline  the oracle's C port of the same workload on this box's host cores (+ BASELINE.md
                section 3's config-1 protocol: batch 256, 1 core and all cores, both MLP modes)
  latency       p50 / p99 of one batch of 2048 alone (small and medium model), BASELINE.json's 2nd metric
  stress        configs[4] per-GPU slice: lookup+concat GB/s against the HBM roofline, the 3rd metric
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))

def trace(msg):
    """Progress marks on stderr (BENCH_TRACE=1): tells which phase a hung run was in."""
    if os.environ.get("BENCH_TRACE"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


METRIC = "inferences/sec (gather+MLP)"
UNIT = "inferences/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], bf16=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained"),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel_substr):
    """DRAM bytes per launch (read + write) of the newest committed `ncu --set full` capture of a kernel whose
    name contains `kernel_substr` (profiles/*_step.csv / *_gather_stress.csv, written by tools/ncu_summary.py);
    None when no capture is committed.  The capture is of the same command at the same batch size."""
    import csv
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step.csv")) +
                       glob.glob(os.path.join(ROOT, "profiles", "r*_gather_stress.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 2:
            continue
        hdr = rows[0]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        cols = []
        for i, h in enumerate(hdr):
            if h.startswith("dram__bytes_read.sum") or h.startswith("dram__bytes_write.sum"):
                cols.append((i, scale.get(h[h.index("[") + 1:-1], 1.0)))
        norm = lambda n: n.replace("(int)", "").replace("(bool)", "").replace(" ", "")   # noqa: E731
        vals = [sum(float(r[i]) * k for i, k in cols) for r in rows[1:] if norm(kernel_substr) in norm(r[0])]
        if vals and len(cols) == 2:
            best = dict(bytes_per_launch=sum(vals) / len(vals), source=os.path.relpath(path, ROOT), launches=len(vals))
    return best


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm (host cores)
def host_threads():
    """Cores this process may use.  Passed to the port explicitly: torchrun exports OMP_NUM_THREADS=1, which would
    silently turn the N > 1 reference arm into a one-core run."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def quantum(args):
    """batches per step"""
    return args.rounds * args.streams


def cpu_config1(cat, tables, dims, W, b, threads_all):
    """BASELINE.md section 3 / SURVEY.md 8(d) config 1: small model, batch 256, LINEAR and BIAS_RELU_SIGMOID, on 1
    core and on all cores, 20 warm-up + 200 timed iterations, per-batch latency percentiles."""
    from oracle import oracle
    out = []
    batches = [oracle.zipf_indices(cat, 256, seed=4000 + i) for i in range(8)]
    alg = 256 * cat.gather_bytes_per_item(materialised=True)
    for th in (1, threads_all):
        for mode, name in ((0, "linear"), (1, "bias_relu_sigmoid")):
            lat = []
            for i in range(220):
                t0 = time.perf_counter()
                x = oracle.gather(cat, tables, batches[i % 8], threads=th)
                t1 = time.perf_counter()
                oracle.mlp(x, dims, W, b if mode else None, mode=mode, threads=th)
                t2 = time.perf_counter()
                if i >= 20:
                    lat.append((t2 - t0, t1 - t0))
            tot = np.array([a for a, _ in lat])
            g = np.array([c for _, c in lat])
            out.append(dict(cores=th, mlp_mode=name, batch=256, iterations=len(lat), inferences_per_s=256 * len(lat) / float(tot.sum()),
                            batch_ms_p50=float(np.percentile(tot, 50) * 1e3), batch_ms_p99=float(np.percentile(tot, 99) * 1e3),
                            gather_gbs=alg / float(np.median(g)) / 1e9))
    return out


def cpu_port_run(args, steps, warmup, budget_s=None, total_budget_s=None, config1=False):
    """The reference's algorithm (oracle C port, OpenMP) on the host: lookup+concat in wire order, then the
    4-layer MLP, batch by batch.  A step is the same quantum as the CUDA arm's (`--rounds` x `--streams` batches
    of args.batch items) or -- when `total_budget_s` would be exceeded -- a bounded sample of it (its first
    `n_b` batches)."""
    from fleetrec import catalogue
    from oracle import oracle
    cat = catalogue.load(args.model)
    dims = cat.layer_dims
    cores = host_threads()
    tables = oracle.make_tables(cat, "hash", seed=0x5EED)
    W, b = oracle.make_weights(dims, seed=42)
    batches = [oracle.zipf_indices(cat, args.batch, seed=1234 + i) for i in range(4)]

    def one(i, th=cores):
        oracle.mlp(oracle.gather(cat, tables, batches[i % 4], threads=th), dims, W, b, mode=1, threads=th)
    for i in range(2):
        t0 = time.perf_counter()
        one(i)
        t_batch = time.perf_counter() - t0
    n_b = quantum(args)
    if total_budget_s and (steps + warmup) * n_b * t_batch > total_budget_s:
        n_b = max(1, int(total_budget_s / ((steps + warmup) * t_batch)))
    for i in range(warmup * n_b if total_budget_s else 0):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        for j in range(n_b):
            one(i * n_b + j)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    res = dict(value=done * n_b * args.batch / dt, unit=UNIT, cores=cores, kind="port",
               sample=f"{done} steps of {n_b} batches of {args.batch} items (a full step is {quantum(args)} batches), "
                      f"{args.model} model full-size tables ({cat.table_bytes() / 1e9:.2f} GB, hash fill), Zipf(1.05) "
                      f"indices, oracle/fr_oracle.c gather + fp32 MLP, OpenMP num_threads({cores}) set explicitly")
    if config1:
        res["config1"] = cpu_config1(cat, tables, dims, W, b, cores)
    return res, dt / max(done, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, per_step = cpu_port_run(args, args.steps, args.warmup, total_budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference algorithm on host cores: the FPGA lookup cannot run here and the reference has no CPU "
                    "MLP, so this is the oracle's C port (cpu_baseline.kind = port)"}
    print(json.dumps(line))


def workload_config(args, n):
    return {"workload": f"FleetRec {args.model} model (BASELINE.json configs[1]): all tables HBM-resident, "
                        f"gather+MLP at batch {args.batch}, Zipf(1.05) indices",
            "model_tables": args.model, "batch": args.batch, "global_batch": args.batch * n,
            "streams": args.streams, "step": f"{args.rounds} batches on each of {args.streams} worker streams = "
                                             f"{quantum(args)} batches = {quantum(args) * args.batch * n} items",
            "batches_per_step": quantum(args), "e2e_batches_per_copy": args.group,
            "mlp_mode": "bias_relu_sigmoid", "precision": args.precision,
            "index_rows": "int32 per table (the reference's index stream)" if args.index_format == "i32" else
                          "packed transport format: uint16 for tables of at most 65536 rows, int32 otherwise (FR_IDX_PACKED)",
            "sharding": "single GPU" if n == 1 else (
                "tables sharded across ranks (on-chip-class tables replicated), every rank fed the index columns of its "
                "own tables, pieces pushed over NVLink by the lookup kernel, batch-parallel MLP" if args.shard == "tables" else "replicated tables, independent batches"),
            "l2": "tables (1.4 GB) exceed L2; a pool of distinct index batches is rotated so no step repeats "
                  "the previous step's inputs"}


# --------------------------------------------------------------------------- CUDA arm
def run_ours(args):
    import torch

    import fleetrec
    from fleetrec import catalogue
    from oracle import oracle  # index/weight generators + the cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cat = catalogue.load(args.model)
    dims = cat.layer_dims
    B, T, G = args.batch, cat.n_tables, args.group
    prec = fleetrec.FR_PREC_TF32 if args.precision == "tf32" else fleetrec.FR_PREC_FP32
    # N > 1, --shard tables: every rank owns a table subset, looks it up for the GLOBAL batch and
    # pushes the pieces over NVLink into the concat buffer of the rank that owns the item; the MLP is
    # batch-parallel.  Per-GPU work is fixed (weak scaling): global batch = N x 2048.
    sharded = world > 1 and args.shard == "tables"
    Bg = B * world if sharded else B
    extras = world == 1 and not args.quick          # the per-kernel / large-batch / latency / stress / fp16 legs
    mb = max(Bg, args.gather_batch if extras else 0)
    eng = fleetrec.Engine(cat, device=local, precision=prec, max_batch=(mb + world - 1) // world * world)
    eng.set_option(fleetrec.FR_OPT_TILE_HINT, fleetrec.FR_HINT_THROUGHPUT)   # `--streams` batches in flight
    if sharded:
        from fleetrec import shard
        owner = shard.plan_owners(cat, world, replicate_below_bytes=args.replicate_mb << 20, policy=args.plan)
        eng.shard_init(rank, world, owner)   # every worker stream gets its own exchange slot
    eng.fill_hash(seed=0x5EED)
    W, b = oracle.make_weights(dims, seed=42)
    eng.load_mlp(W, b)
    if sharded:
        eng.shard_import(shard.exchange_handles(eng, dist, device="cuda"))
        dist.barrier()
    # index rows travel in the engine's transport format: int32 columns (the reference's index stream), or with
    # --index-format packed uint16 columns for the tables of at most 65536 rows (packed on the host, outside the timed
    # regions, as the index source would produce them)
    eng.set_option(fleetrec.FR_OPT_INDEX_FORMAT, fleetrec.FR_IDX_PACKED if args.index_format == "packed" else fleetrec.FR_IDX_I32)
    lay_full = eng.index_layout(2)
    lay_own, lay_rep = (eng.index_layout(0), eng.index_layout(1)) if sharded else (None, None)

    def pack(a, lay=None):
        return fleetrec.pack_indices(a, lay or lay_full)
    workers = [fleetrec.Worker(eng) for _ in range(args.streams)]
    wstreams = [torch.cuda.ExternalStream(w.cuda_stream) for w in workers]
    main = torch.cuda.current_stream()

    pool = 32
    # sharded: every rank sees the same global batch (same seed); replicated: its own batches
    idx_np = [oracle.zipf_indices(cat, Bg, seed=1234 + (0 if sharded else 1000 * rank) + i) for i in range(pool)]
    idx_host = [torch.from_numpy(pack(a)).pin_memory() for a in idx_np]
    idx_dev = [t.cuda(non_blocking=True) for t in idx_host]
    # e2e, single GPU / replicated: fr_infer_many takes `G` batches per call from ONE pinned buffer [G][B][T]
    grp_host = grp_sc = None
    if not sharded:
        n_grp = pool // G
        grp_host = [torch.from_numpy(pack(np.concatenate(idx_np[g * G:(g + 1) * G]))).pin_memory() for g in range(n_grp)]
        grp_sc = [torch.empty(G * B, dtype=torch.float32).pin_memory() for _ in range(args.streams)]
    # sharded: every rank is fed the column slices it needs (fr_shard_infer_sliced) -- the indices of the tables it
    # owns for ALL items and of the replicated tables for its own items -- sliced on the host outside the timed
    # region, as the reference's index source feeds every FPGA only its own tables' indices
    sl_host = sl_dev = None
    if sharded:
        def sliced(a):   # this rank's two column blocks of a global batch, in the transport format
            o, r = shard.slice_indices(a, owner, world, rank)
            return pack(o, lay_own), pack(r, lay_rep)

        def packed(a):   # one pinned buffer, the replicated block right behind the owned one: ONE copy per step
            o, r = sliced(a)
            n_o = (o.size + 3) // 4 * 4
            buf = torch.empty(n_o + r.size, dtype=torch.int32).pin_memory()
            buf[:o.size] = torch.from_numpy(o.reshape(-1))
            buf[n_o:] = torch.from_numpy(r.reshape(-1))
            return buf, (buf[:o.size], buf[n_o:])
        sl_bufs = [packed(a) for a in idx_np]            # keep the buffers alive
        sl_host = [views for _, views in sl_bufs]
        sl_dev = [tuple(a.cuda(non_blocking=True) for a in pair) for pair in sl_host]

        def packed_group(arrs):   # G consecutive steps for fr_shard_infer_sliced_many: [owned blocks of all G | replicated blocks]
            o = np.concatenate([sliced(a)[0].reshape(-1) for a in arrs])
            r = np.concatenate([sliced(a)[1].reshape(-1) for a in arrs])
            n_o = (o.size + 3) // 4 * 4
            buf = torch.zeros(n_o + r.size, dtype=torch.int32).pin_memory()
            buf[:o.size] = torch.from_numpy(o)
            buf[n_o:] = torch.from_numpy(r)
            return buf, (buf[:o.size], buf[n_o:])
        slg_bufs = [packed_group(idx_np[g * G:(g + 1) * G]) for g in range(pool // G)]
        slg_host = [views for _, views in slg_bufs]
        slg_sc = [torch.empty(G * B, dtype=torch.float32).pin_memory() for _ in range(args.streams)]
    sc_dev = [torch.empty(B, dtype=torch.float32, device="cuda") for _ in range(args.streams)]
    sc_host = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(args.streams)]
    torch.cuda.synchronize()

    trace("setup done, parity gate")
    # correctness gate before timing anything: one batch against the oracle (hash-filled tables)
    i0 = idx_np[0]
    lo, hi = (rank * B, (rank + 1) * B) if sharded else (0, B)
    exp_x = oracle.gather_hashed(cat, 0x5EED, i0[lo:hi])
    exp_s = oracle.mlp(exp_x, dims, W, b, mode=1)
    tol = 1e-3 if args.precision == "tf32" else 2e-5

    def gate(what):
        err = float(np.max(np.abs(sc_host[0].numpy() - exp_s) / np.maximum(np.abs(exp_s), 1e-6)))
        assert err <= tol, f"score parity gate failed ({what}): {err}"
        return err
    if sharded:
        eng.shard_infer_sliced(sl_host[0][0].numpy(), sl_host[0][1].numpy(), Bg, sc_host[0].numpy(), workers[0])
        eng.sync(workers[0])
        dist.barrier()
    else:
        got = eng.gather_only(idx_host[0].numpy())
        assert np.array_equal(got.view(np.uint32), exp_x.view(np.uint32)), "concat not bit-exact"
        eng.infer_async(idx_host[0].numpy(), sc_host[0].numpy(), B, workers[0])
        eng.sync(workers[0])
    gate_err = gate("tf32" if args.precision == "tf32" else "fp32")
    if not sharded:   # the grouped call computes what the per-batch call computes (group 0 starts with batch 0)
        eng.infer_many_async(grp_host[0].numpy(), grp_sc[0].numpy(), G, B, workers[0])
        eng.sync(workers[0])
        assert np.array_equal(grp_sc[0][:B].numpy().view(np.uint32), sc_host[0].numpy().view(np.uint32)), "fr_infer_many"

    host_us = {}   # host time to enqueue one batch (if it approaches the device time per batch the host is the limit)
    graph_misses = {}
    S, R = args.streams, args.rounds

    def timed(batch_fn, per_call, steps, warmup, name):
        """batch_fn(j): enqueue call j of the endless stream (worker j % S); per_call = batches per call.
        One step = R * S batches."""
        calls = R * S // per_call
        trace(f"timed({name}, {steps}, {warmup}) graphs")
        # one-time setup, like loading weights: every (index buffer, score buffer, worker) combination the stream
        # uses is shown to the engine once, so its CUDA graphs are instantiated before the W warm-up steps begin
        # (the first call of a batch size on an engine runs un-captured: + S)
        n_pool = pool // per_call
        for j in range(n_pool * S // np.gcd(n_pool, S) + S):
            batch_fn(j)
        barrier()
        trace("warmup")
        for i in range(warmup):
            for j in range(calls):
                batch_fn(i * calls + j)
        barrier()
        trace("timed region")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0, g0 = eng.launch_count(), eng.graph_stats()
        e0.record(main)
        for s_ in wstreams:
            s_.wait_event(e0)
        t_host = time.perf_counter()
        for j in range(steps * calls):
            batch_fn(j)
        host_us[name] = (time.perf_counter() - t_host) / (steps * calls * per_call) * 1e6
        for s_ in wstreams:
            ev = torch.cuda.Event()
            ev.record(s_)
            main.wait_event(ev)
        e1.record(main)
        trace("enqueued, waiting")
        barrier()
        trace("done")
        g1 = eng.graph_stats()
        graph_misses[name] = {k: g1[k] - g0[k] for k in ("captured", "direct")}
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, eng.launch_count() - l0

    # raw addresses once, outside the loops: the timed regions should measure the library, not numpy/ctypes
    # attribute lookups (the buffers stay referenced by the lists above)
    p_idx_dev = [t.data_ptr() for t in idx_dev]
    p_sc_dev, p_sc_host = [t.data_ptr() for t in sc_dev], [t.data_ptr() for t in sc_host]
    if sharded:
        p_sl_dev = [(a.data_ptr(), r.data_ptr()) for a, r in sl_dev]
        p_slg_host = [(a.data_ptr(), r.data_ptr() if r.numel() else None) for a, r in slg_host]
        p_slg_sc = [t.data_ptr() for t in slg_sc]
    else:
        p_grp_host, p_grp_sc = [t.data_ptr() for t in grp_host], [t.data_ptr() for t in grp_sc]

    def batch_dev(j):
        w = j % S
        if sharded:
            eng.shard_infer_sliced(p_sl_dev[j % pool][0], p_sl_dev[j % pool][1], Bg, p_sc_dev[w], workers[w])
        else:
            eng.infer_async(p_idx_dev[j % pool], p_sc_dev[w], B, workers[w])

    def batch_e2e(j):
        w = j % S
        if sharded:
            g = j % len(p_slg_host)
            eng.shard_infer_sliced_many(p_slg_host[g][0], p_slg_host[g][1], G, Bg, p_slg_sc[w], workers[w])
        else:
            eng.infer_many_async(p_grp_host[j % len(p_grp_host)], p_grp_sc[w], G, B, workers[w])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(batch_dev, 1, args.steps, args.warmup, "value")
    ms_e2e, _ = timed(batch_e2e, G, args.steps, args.warmup, "e2e")
    clocks = sampler.stop() if rank == 0 else None

    items_per_step = world * R * S * B
    value = args.steps * items_per_step / (ms_dev * 1e-3)
    e2e = args.steps * items_per_step / (ms_e2e * 1e-3)
    ms_batch = ms_dev / (args.steps * R * S)         # device time per batch of B items per GPU

    if rank != 0:
        eng.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # bytes uploaded per step, all ranks together, counted from the tensors copied
    if sharded:   # (every rank's blocks, from the plan: no collective here -- the other ranks have already left)
        def row_bytes(tabs):
            return shard.index_layout([cat.tables[t].rows for t in tabs], packed=args.index_format == "packed")[2]
        per_rank = [Bg * row_bytes(shard.rank_tables(owner, r)[0]) + B * row_bytes(shard.rank_tables(owner, r)[1]) for r in range(world)]
        per_rank[rank] = Bg * lay_own[2] + B * lay_rep[2]     # this rank's: what fr_index_layout reported
    else:
        per_rank = [B * lay_full[2]] * world
    h2d_bytes, h2d_rank_max = sum(per_rank) * R * S, max(per_rank) * R * S
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (fp32 storage, fp32 accumulate)" if args.precision == "tf32" else "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * 4 * world * R * S,
                    "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_rank_max": h2d_rank_max,
                    "call": f"fr_shard_infer_sliced_many, {G} steps per call" if sharded else f"fr_infer_many, {G} batches per call"},
            "us_per_batch": ms_batch * 1e3, "gpu_launches": int(launches), "host_enqueue_us_per_batch": host_us,
            "graph_misses_in_timed_region": graph_misses, "parity_gate_max_rel_err": gate_err, "clocks": clocks}

    pk = peaks()
    tensor_peak = pk["bf16"] / 2 if args.precision == "tf32" else 2 * 148 * 128 * 1.965e-3   # TF/s
    flops = [0, 2.0 * B * dims[0] * dims[1], 2.0 * B * dims[1] * dims[2],
             2.0 * B * (dims[2] * dims[3] + (dims[3] if args.precision == "tf32" else 0)), 2.0 * B * dims[3]]
    step_flops = sum(flops[1:4])
    step_tf = step_flops / (ms_batch * 1e-3) / 1e12
    whole = dict(bound="tensor", achieved=step_tf, peak=tensor_peak, unit="TFLOP/s", frac=step_tf / tensor_peak,
                 flops_per_batch=step_flops,
                 note="all MLP FLOPs of one batch / device time per batch, %d worker streams in flight (per GPU)" % S)
    if not extras:
        line["roofline"] = dict(whole, kernel="whole step (per-kernel legs run at N = 1 only)", traffic=None,
                                peak_source=pk["src"] + "; tf32 tensor peak taken as half the measured dense bf16 rate")
        eng.close()
        print(json.dumps(line))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # Everything below explains the two numbers above; a failure in one of these legs must not lose them.
    try:
        # ---- per-kernel times (each kernel alone, CUDA events on its own stream) and rooflines
        trace("per-kernel times")
        kms = eng.time_kernels(idx_dev[1], B, reps=args.kernel_reps, worker=workers[0])
        names = ["gather_concat", "mlp_layer1", "mlp_layer2", "mlp_layer3+out", "mlp_out"]
        # SURVEY 8(d)'s per-item figure with the index bytes this run actually reads (packed rows are shorter than T*4)
        item_bytes = cat.gather_bytes_per_item(materialised=True) - 4 * T + lay_full[2]
        gather_bytes = B * item_bytes
        # SMs every MLP launch occupies (one CTA per SM): a 2048-item batch is 8..32 tiles, so a kernel timed ALONE runs
        # on 8..32 of the 148 SMs -- `frac` (the contract's definition, against the whole device) is small by
        # construction; `frac_of_occupied_sms` relates it to the tensor peak of the SMs it actually held, and
        # roofline.whole_step to what the device sustains with `--streams` batches in flight.
        import ctypes as C
        from fleetrec import _capi
        raw = C.CDLL(_capi.LIB_PATH)
        raw.frdbg_layer_ctas.argtypes = [C.c_void_p, C.c_int]
        ctas = [0] + [int(raw.frdbg_layer_ctas(eng._h, k)) for k in range(3)] + [0]
        kernels = []
        for ki, (n, ms, fl) in enumerate(zip(names, kms, flops)):
            if ms <= 0:
                continue
            if n == "gather_concat":
                a = gather_bytes / (ms * 1e-3) / 1e9
                kernels.append(dict(name=n, ms=ms, bound="hbm", achieved=a, peak=pk["hbm"], unit="GB/s", frac=a / pk["hbm"]))
            else:
                a = fl / (ms * 1e-3) / 1e12
                k = dict(name=n, ms=ms, bound="tensor", achieved=a, peak=tensor_peak, unit="TFLOP/s", frac=a / tensor_peak)
                if ctas[ki] > 0:
                    k.update(sms_occupied=ctas[ki], frac_of_occupied_sms=a / (tensor_peak * ctas[ki] / 148.0))
                kernels.append(k)
        dom = max(kernels, key=lambda k: k["ms"])
        # the committed ncu --set full capture of the same command: which kernel instance is the dominant one
        ncu_name = {"gather_concat": "gather_concat_kernel", "mlp_layer1": "tc_pair_kernel<256,3,0,4,2",
                    "mlp_layer2": "tc_pair_kernel<512,4,0,4,1", "mlp_layer3+out": "tc_pair_kernel<256,3,1,4,2"}.get(dom["name"])
        tr = ncu_traffic(ncu_name) if (ncu_name and args.model == "small" and B == 2048) else None
        roofline = dict(bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"], unit=dom["unit"], frac=dom["frac"],
                        traffic=tr["bytes_per_launch"] if tr else None, traffic_source=tr["source"] if tr else None,
                        kernel=dom["name"], ms_per_launch=dom["ms"],
                        peak_source=pk["src"] + ("; tf32 tensor peak taken as half the measured dense bf16 rate"
                                                 if dom["bound"] == "tensor" and args.precision == "tf32" else ""),
                        share_of_step=dom["ms"] / sum(k["ms"] for k in kernels),
                        sms_occupied=dom.get("sms_occupied"), frac_of_occupied_sms=dom.get("frac_of_occupied_sms"))
        # every MLP kernel launched on ALL worker streams at once: what that kernel sustains at the occupancy it has
        # inside the timed region (alone it holds `sms_occupied` SMs)
        layer_of = {"mlp_layer1": 0, "mlp_layer2": 1, "mlp_layer3+out": 2}
        raw.frdbg_enqueue_layer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]

        def burst(kd, reps):
            for _ in range(reps):
                for w in workers:
                    rc = raw.frdbg_enqueue_layer(eng._h, kd, B, w._h)
                    assert rc == 0, eng._L.fr_last_error(eng._h)
        reps_c = 20
        for k in kernels:
            if k["name"] not in layer_of:
                continue
            kd = layer_of[k["name"]]
            burst(kd, 3)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(5e6))   # ~2.5 ms gate: the launches below queue up behind it, so the host's launch
            c0.record(main)               # rate (~4 us per un-graphed launch) is not what is measured
            for st_ in wstreams:
                st_.wait_event(c0)
            burst(kd, reps_c)
            for st_ in wstreams:
                ev = torch.cuda.Event()
                ev.record(st_)
                main.wait_event(ev)
            c1.record(main)
            torch.cuda.synchronize()
            ms_eff = c0.elapsed_time(c1) / (reps_c * len(workers))
            a_c = flops[1 + kd] / (ms_eff * 1e-3) / 1e12
            k["at_step_occupancy"] = dict(
                achieved=a_c, peak=tensor_peak, unit="TFLOP/s", frac=a_c / tensor_peak, ms_per_launch_effective=ms_eff,
                note="the same kernel on all %d worker streams at once, %d launches: FLOPs of all launches / elapsed" %
                     (len(workers), reps_c * len(workers)))
        if "at_step_occupancy" in dom:
            roofline["at_step_occupancy"] = dom["at_step_occupancy"]
        # the step as a whole: its kernels overlap across the worker streams, so the dominant kernel timed
        # alone (above) understates what the device sustains -- all MLP FLOPs of a batch over the time per batch
        roofline["whole_step"] = whole

        # ---- the same kernels at a large batch (north star: tensor-pipe utilisation at batch >= 4096)
        large = None
        if args.gather_batch >= 4096:
            LB = args.gather_batch
            lidx = torch.from_numpy(pack(oracle.zipf_indices(cat, LB, seed=99))).cuda()
            lms = eng.time_kernels(lidx, LB, reps=max(args.kernel_reps // 2, 2), worker=workers[0])
            lfl = [f * LB / B for f in flops]
            ltot = sum(lfl[1:4])
            lctas = [0] + [int(raw.frdbg_layer_ctas(eng._h, k)) for k in range(3)] + [0]
            large = dict(batch=LB, kernels=[dict(name=n, ms=ms, achieved=fl / (ms * 1e-3) / 1e12, unit="TFLOP/s",
                                                 frac=fl / (ms * 1e-3) / 1e12 / tensor_peak, sms_occupied=c_)
                                            for n, ms, fl, c_ in zip(names, lms, lfl, lctas) if ms > 0 and fl > 0],
                         mlp_ms=sum(lms[1:]), mlp_tflops=ltot / (sum(lms[1:]) * 1e-3) / 1e12,
                         mlp_frac=ltot / (sum(lms[1:]) * 1e-3) / 1e12 / tensor_peak, peak=tensor_peak)

        # ---- stand-alone gather at a large batch, uniform indices (the lookup of THIS model against the HBM roofline)
        GB = args.gather_batch
        gidx = torch.from_numpy(pack(oracle.uniform_indices(cat, GB, seed=4321))).cuda()
        gout = torch.empty(GB, cat.concat_floats, dtype=torch.float32, device="cuda")
        for _ in range(3):
            eng.gather_only_async(gidx, gout, GB, workers[0])
        eng.sync(workers[0])
        eng.mark(0, workers[0])
        for _ in range(20):
            eng.gather_only_async(gidx, gout, GB, workers[0])
        eng.mark(1, workers[0])
        gms = eng.elapsed_ms(workers[0]) / 20
        g_alg = GB * item_bytes
        gather = dict(batch=GB, indices="uniform", ms=gms, achieved=g_alg / (gms * 1e-3) / 1e9, peak=pk["hbm"],
                      unit="GB/s", frac=g_alg / (gms * 1e-3) / 1e9 / pk["hbm"],
                      bytes_per_item=item_bytes)
        del gidx, gout

        # ---- the same timed stream on fp16 operands, where the engine's range analysis allows them
        f16 = None
        if args.precision == "tf32":
            trace("fp16-guarded leg")
            eng.set_option(fleetrec.FR_OPT_F16_OPERANDS, fleetrec.FR_F16_GUARDED)
            active, bounds = eng.f16_report()
            f16 = dict(active=active, bounds=dict(zip(("x", "h1", "h2", "min_nonzero_table", "inexact_weight_share"), bounds)))
            if active:
                eng.infer_async(idx_host[0].numpy(), sc_host[0].numpy(), B, workers[0])
                eng.sync(workers[0])
                f16["parity_gate_max_rel_err"] = gate("fp16 operands")
                ms16, _ = timed(batch_dev, 1, args.steps, args.warmup, "value_f16")
                ms16e, _ = timed(batch_e2e, G, args.steps, args.warmup, "e2e_f16")
                f16.update(value=args.steps * items_per_step / (ms16 * 1e-3), e2e=args.steps * items_per_step / (ms16e * 1e-3),
                           unit=UNIT, us_per_batch=ms16 / (args.steps * R * S) * 1e3,
                           dtype="f16 operands and activations (same 11-bit significand as tf32), f32 accumulate; chosen by "
                                 "the engine's range analysis (FR_F16_GUARDED), tf32 otherwise",
                           whole_step_tflops=step_flops / (ms16 / (args.steps * R * S) * 1e-3) / 1e12)
            eng.set_option(fleetrec.FR_OPT_F16_OPERANDS, fleetrec.FR_F16_OFF)

        for w in workers:
            w.close()
        eng.close()
        del idx_dev, sc_dev
        torch.cuda.empty_cache()

        line.update(roofline=roofline, kernels=kernels, gather_standalone=gather, mlp_large_batch=large, f16_guarded=f16)
        if not args.no_latency:
            trace("latency leg")
            line["latency"] = latency_leg(args, ("small", "medium"), (B,), args.latency_launches)
        if not args.no_stress:
            trace("stress leg")
            line["stress"] = stress_leg(args, local, 0, 1, None)
        trace("cpu baseline")
        line["cpu_baseline"] = cpu_port_run(args, 10 ** 9, 1, budget_s=args.cpu_seconds, config1=True)[0] if args.cpu_seconds > 0 else None
    except Exception as ex:   # noqa: BLE001 -- report and keep the headline
        import traceback
        traceback.print_exc(file=sys.stderr)
        line["legs_error"] = f"{type(ex).__name__}: {ex}"
        line.setdefault("roofline", dict(whole, kernel="whole step (per-kernel legs failed)", traffic=None,
                                         peak_source=pk["src"] + "; tf32 tensor peak taken as half the measured dense bf16 rate"))
        line.setdefault("cpu_baseline", None)
    print(json.dumps(line))


# --------------------------------------------------------------------------- latency of one batch alone (2nd metric)
def latency_leg(args, models, batches, n):
    """p50 / p99 of ONE batch in flight (BASELINE.json: 'p99 batch latency'): per model and batch size, device time
    of one fr_infer (CUDA events around each graph launch on the worker's stream, `n` launches) and host wall
    time of fr_infer -> fr_sync on pinned buffers.  The engine has one worker: latency-oriented tiles."""
    import torch

    import fleetrec
    from fleetrec import catalogue
    from oracle import oracle
    out = []
    for model in models:
        cat = catalogue.load(model)
        dims = cat.layer_dims
        prec = fleetrec.FR_PREC_TF32 if args.precision == "tf32" else fleetrec.FR_PREC_FP32
        eng = fleetrec.Engine(cat, device=torch.cuda.current_device(), precision=prec, max_batch=max(batches))
        eng.fill_hash(seed=0x5EED)
        W, b = oracle.make_weights(dims, seed=42)
        eng.load_mlp(W, b)
        w = fleetrec.Worker(eng)
        ws = torch.cuda.ExternalStream(w.cuda_stream)
        for B in batches:
            idx_d = [torch.from_numpy(oracle.zipf_indices(cat, B, seed=77 + i)).cuda() for i in range(4)]
            idx_h = [t.cpu().pin_memory() for t in idx_d]
            sc_d = torch.empty(B, dtype=torch.float32, device="cuda")
            sc_h = torch.empty(B, dtype=torch.float32).pin_memory()
            for i in range(12):                       # graphs instantiated, caches warm
                eng.infer_async(idx_d[i % 4], sc_d, B, w)
                eng.infer_async(idx_h[i % 4].numpy(), sc_h.numpy(), B, w)
            eng.sync(w)
            if B <= 4096:                             # parity gate inside the leg
                exp = oracle.mlp(oracle.gather_hashed(cat, 0x5EED, idx_h[3].numpy()), dims, W, b, mode=1)
                err = float(np.max(np.abs(sc_h.numpy() - exp) / np.maximum(np.abs(exp), 1e-6)))
                assert err <= (1e-3 if args.precision == "tf32" else 2e-5), err
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for i, (e0, e1) in enumerate(evs):
                e0.record(ws)
                eng.infer_async(idx_d[i % 4], sc_d, B, w)
                e1.record(ws)
                if i % 64 == 63:
                    eng.sync(w)                       # isolated launches: latency, not pipelined throughput
            eng.sync(w)
            dev = np.array([e0.elapsed_time(e1) for e0, e1 in evs]) * 1e3
            wall = []
            for i in range(min(n, 300)):
                t0 = time.perf_counter()
                eng.infer_async(idx_h[i % 4].numpy(), sc_h.numpy(), B, w)
                eng.sync(w)
                wall.append((time.perf_counter() - t0) * 1e6)
            wall = np.array(wall)
            out.append(dict(model=model, batch=B, dev_p50_us=float(np.percentile(dev, 50)), dev_p99_us=float(np.percentile(dev, 99)),
                            host_p50_us=float(np.percentile(wall, 50)), host_p99_us=float(np.percentile(wall, 99)),
                            inferences_per_s=B / (float(np.percentile(dev, 50)) * 1e-6), launches=n))
        w.close()
        eng.close()
        torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- config 5: gather HBM-roofline stress
def stress_leg(args, local, rank, world, dist):
    """BASELINE.json configs[4], per-GPU slice: T tables x R rows x dim 64 (256-byte rows),
    lookup+concat only, uniform indices (tables >> L2: algorithmic bytes ~ DRAM bytes, the honest
    HBM-roofline test) and Zipf(1.05) indices (hot heads become L2-resident).  With N ranks every rank
    holds its own T x R tables (the 1000 tables split 125 per GPU; weak scaling, no exchange: the lookup of
    disjoint table sets is embarrassingly parallel).  Returns the result dict on rank 0 (None elsewhere)."""
    import torch

    import fleetrec
    from fleetrec import catalogue
    from oracle import oracle

    T, dim, B = args.stress_tables, 64, args.gather_batch
    rows, eng = args.stress_rows, None
    tdt = {"f32": fleetrec.FR_TABLE_F32, "f16": fleetrec.FR_TABLE_F16, "bf16": fleetrec.FR_TABLE_BF16}[args.table_dtype]
    esize = 4 if tdt == fleetrec.FR_TABLE_F32 else 2
    while eng is None:
        cat = catalogue.synthetic(T, rows, dim)
        eng = fleetrec.Engine(cat, device=local, max_batch=B, table_dtype=tdt)
        try:
            eng.fill_hash(seed=0x5EED)
        except fleetrec.FleetRecError as ex:        # HBM too small for this many rows: halve and say so
            if ex.code != fleetrec.FR_ERR_OOM or rows <= 65536:
                raise
            eng.close()
            eng, rows = None, rows // 2
    w = fleetrec.Worker(eng)
    out = torch.empty(B, cat.concat_floats, dtype=torch.float32, device="cuda")
    pk = peaks()
    # algorithmic bytes per item: rows at their storage width + int32 indices + the fp32 concat written
    per_item = T * dim * esize + T * 4 + cat.concat_floats * 4
    alg = B * per_item
    steps = args.stress_steps
    res = {}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for kind in ("uniform", "zipf"):
        gen = oracle.uniform_indices if kind == "uniform" else oracle.zipf_indices
        pool = [torch.from_numpy(gen(cat, B, seed=4321 + 17 * i + 1000 * rank)).cuda() for i in range(4)]
        # parity gate on a sample: bit-exact against the oracle's hash fill
        eng.gather_only_async(pool[0], out, B, w)
        eng.sync(w)
        exp = oracle.quantize_dequantize(oracle.gather_hashed(cat, 0x5EED, pool[0][:128].cpu().numpy()), tdt)
        assert np.array_equal(out[:128].cpu().numpy().view(np.uint32), exp.view(np.uint32)), "concat not bit-exact"
        for i in range(5):
            eng.gather_only_async(pool[i % 4], out, B, w)
        eng.sync(w)
        if dist is not None:
            dist.barrier()
        eng.mark(0, w)
        for i in range(steps):
            eng.gather_only_async(pool[i % 4], out, B, w)
        eng.mark(1, w)
        ms = eng.elapsed_ms(w)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        per = ms / steps
        res[kind] = dict(ms_per_launch=per, items_per_s=world * B / (per * 1e-3), launches=steps,
                         achieved=alg / (per * 1e-3) / 1e9, frac=alg / (per * 1e-3) / 1e9 / pk["hbm"])
    clocks = sampler.stop() if rank == 0 else None
    w.close()
    eng.close()
    del out
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    u = res["uniform"]
    return {"workload": f"BASELINE.json configs[4] per-GPU slice: {T} tables x {rows} rows x dim {dim} "
                        f"({T * rows * dim * esize / 1e9:.1f} GB per GPU, {args.table_dtype} rows), lookup+concat only, "
                        f"batch {B}, {world} GPU(s) each holding its own {T} tables",
            "rows_requested": args.stress_rows, "rows_used": rows,
            "note": "the literal config (1000 x 10M x 64 fp32 = 2.56 TB) exceeds 8 x 180 GB; rows are scaled so one GPU's "
                    "slice fits HBM, tables stay >> L2 (126 MB)",
            "l2": "uniform indices over tables far larger than L2; a pool of 4 index batches rotates",
            "gather_hbm_gbs_per_gpu": u["achieved"], "gather_hbm_gbs_all_gpus": world * u["achieved"],
            "roofline": {"bound": "hbm", "achieved": u["achieved"], "peak": pk["hbm"], "unit": "GB/s", "frac": u["frac"],
                         "traffic_ncu": ncu_traffic("gather_concat_kernel<0"),   # same launch shape, 1 M-row tables
                         "kernel": "gather_concat", "ms_per_launch": u["ms_per_launch"],
                         "peak_source": pk["src"], "algorithmic_bytes_per_item": per_item},
            "uniform": res["uniform"], "zipf": res["zipf"], "clocks": clocks, "n_gpus": world}


def run_stress(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    st = stress_leg(args, local, rank, world, dist)
    if rank == 0:
        u = st["uniform"]
        esize = 4 if args.table_dtype == "f32" else 2
        line = {"metric": "gather HBM GB/s (lookup+concat, stress tables)", "value": st["gather_hbm_gbs_all_gpus"], "unit": "GB/s",
                "n_gpus": world, "steps": u["launches"], "warmup": 5, "ms_per_step": u["ms_per_launch"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 rows (byte copy)" if esize == 4 else f"{args.table_dtype} rows widened to f32 (exact)",
                "data": "synthetic", "config": {k: st[k] for k in ("workload", "rows_requested", "rows_used", "note", "l2")},
                "gpu_launches": int(2 * (u["launches"] + 6)), "roofline": dict(st["roofline"], traffic=None),
                "uniform": st["uniform"], "zipf": st["zipf"], "clocks": st["clocks"]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- config 3: batch sweep, latency vs throughput
def run_sweep(args):
    """BASELINE.json configs[2]: medium model (98 tables, 15.1 GB), B in 1..16384, one batch in flight."""
    import torch
    torch.cuda.set_device(0)
    sampler = ClockSampler(0)
    sampler.start()
    rows = latency_leg(args, (args.model,), [1 << k for k in range(0, 15)], args.sweep_launches)
    clocks = sampler.stop()
    best = max(rows, key=lambda r: r["inferences_per_s"])
    print(json.dumps({"metric": "p99 batch latency / throughput sweep", "unit": "us", "n_gpus": 1,
                      "config": {"workload": f"BASELINE.json configs[2]: FleetRec {args.model} model, batch sweep 1-16384, "
                                             "one batch in flight (latency-oriented tiles), Zipf(1.05) indices, "
                                             "device-resident indices for dev_*; pinned host buffers + sync for host_*",
                                 "precision": args.precision},
                      "value": best["dev_p99_us"], "best_batch": best["batch"], "data": "synthetic", "sweep": rows,
                      "clocks": clocks}))


def arm_hard_limit():
    """A run that has not finished after BENCH_HARD_LIMIT_S (default 900 s; the whole default run takes
    under a minute) is hung: say so and leave, instead of holding the GPU until somebody else's limit."""
    limit = float(os.environ.get("BENCH_HARD_LIMIT_S", "900"))

    def fire():
        print(f"bench.py: no result after {limit:.0f} s -- device or collective hung; aborting", file=sys.stderr, flush=True)
        os._exit(3)
    t = threading.Timer(limit, fire)
    t.daemon = True
    t.start()


def main():
    arm_hard_limit()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="small")
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--streams", type=int, default=0,
                    help="worker streams = batches in flight; 0 = 12 on one GPU (more only adds L2 pressure: 8 -> 184, 10-12 -> 209, 16 -> 204 M/s),\n"
                         "16 for a table-sharded multi-GPU run (more steps in flight hide the exchange: 12 -> 1314, 16 -> 1453 M/s at 8 GPUs)")
    ap.add_argument("--rounds", type=int, default=64, help="a step = this many batches on every worker stream")
    ap.add_argument("--group", type=int, default=4, help="e2e: batches per fr_infer_many call (one copy each way per call)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--index-format", default="packed", choices=["packed", "i32"],
                    help="transport format of the index rows (FR_OPT_INDEX_FORMAT): packed = uint16 columns for small tables")
    ap.add_argument("--tiles", default="", help="FR_TC_TILES override: N1,N2,N3,ctas")
    ap.add_argument("--shard", default="tables", choices=["replicated", "tables"],
                    help="N > 1: shard tables across ranks with the NVLink push exchange (north star), or replicate")
    ap.add_argument("--plan", default="contiguous", choices=["balanced", "contiguous"],
                    help="table sharding: owner plan (fleetrec.shard.plan_owners policy)")
    ap.add_argument("--replicate-mb", type=int, default=0,
                    help="table sharding: also replicate any table smaller than this many MiB (0: only the on-chip class)")
    ap.add_argument("--gather-batch", type=int, default=16384)
    ap.add_argument("--kernel-reps", type=int, default=50)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--quick", action="store_true", help="N = 1: only value / e2e (no per-kernel, latency, stress, fp16, cpu legs)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-stress", action="store_true")
    ap.add_argument("--latency-launches", type=int, default=1000)
    ap.add_argument("--workload", default="model", choices=["model", "stress", "sweep"],
                    help="model: gather+MLP of --model (the headline); stress: configs[4] gather HBM roofline; "
                         "sweep: configs[2] latency/throughput batch sweep")
    ap.add_argument("--stress-tables", type=int, default=125)
    ap.add_argument("--stress-rows", type=int, default=4000000)
    ap.add_argument("--stress-steps", type=int, default=50)
    ap.add_argument("--table-dtype", default="f32", choices=["f32", "f16", "bf16"],
                    help="stress workload: storage type of the tables (SURVEY.md 8(f)4)")
    ap.add_argument("--sweep-launches", type=int, default=1000)
    args = ap.parse_args()
    if args.streams <= 0:
        args.streams = 12 if int(os.environ.get("WORLD_SIZE", "1")) == 1 or args.shard != "tables" else 16
    if args.tiles:
        os.environ["FR_TC_TILES"] = args.tiles
    if args.rounds % args.group:
        ap.error("--rounds must be a multiple of --group")
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "stress":
        run_stress(args)
    elif args.workload == "sweep":
        if args.model == "small":
            args.model = "medium"
        run_sweep(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
