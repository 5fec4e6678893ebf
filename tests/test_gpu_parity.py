"""Parity of the CUDA path (through the C ABI) against the oracle -- needs a B200.

Bit-exact for the lookup/concat (memcmp), tolerance for the MLP:
  max_i |s_i - s^_i| / max(|s^_i|, 1e-6) <= 1e-3   (BASELINE.json north_star)
against the oracle's fp32 MLP, in both arithmetic modes.
"""
import hashlib
import os

import numpy as np
import pytest

import fleetrec
from fleetrec import catalogue
from oracle import oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MODELS3 = ("small", "medium", "large_half")
TOL = 1e-3
# FR_TC_TILES (test hook): tcgen05 tile width of layers 1..3 and CTAs per tile (2 = CTA pair: the production kernel;
# 1 = cta_group::1 tiles of the older kernel, experiments build only)
TILES = ("128,128,256,2", "256,256,256,2", "512,512,256,2", "256,512,256,2", "512,128,256,2")
TILES_1CTA = ("128,128,256,1", "256,256,256,1")


def _has_experiments():
    from fleetrec import _capi
    return bool(_capi.lib().fr_build_has_experiments())


# The measured-slower kernel variants exist only in libfleetrec_exp.so (`make exp`).  Their tests
# are deselected when the release library is loaded (tests/conftest.py) and run by test_experimental_build_variants,
# which re-runs pytest on them in a child process with FLEETREC_LIB pointing at the experiments build.
experimental = pytest.mark.experimental


def test_experimental_build_variants():
    import subprocess
    import sys
    from fleetrec import _capi
    if _has_experiments():
        pytest.skip("already running on the experiments build")
    exp = os.path.join(_capi.PKG_DIR, "libfleetrec_exp.so")
    assert os.path.exists(exp), "build it with `make -C gpu-fpga-recommendation-system_b200 exp` (__graft_entry__.build does)"
    env = dict(os.environ, FLEETREC_LIB=exp)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu and experimental",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-1500:] + r.stderr[-500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], tail


def rel_err(got, exp):
    return float(np.max(np.abs(got - exp) / np.maximum(np.abs(exp), 1e-6)))


def assert_bits_equal(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---------------------------------------------------------------- lookup + concat
@pytest.mark.parametrize("model", MODELS3)
def test_gather_matches_reference_golden(model):
    """CUDA gather vs the wire bytes the reference's own kernel emitted (tests/golden)."""
    g = np.load(os.path.join(GOLD, f"ref_{model}.npz"))
    cat = catalogue.load(model).with_row_cap(int(g["row_cap"]))
    eng = fleetrec.Engine(cat, max_batch=128)
    for t in cat.tables:
        tab = oracle.fill_reference(t.rows, t.dim) if t.tier == "PLRAM" else \
            oracle.fill_hash(int(g["seed"]), t.id, t.rows, t.dim)
        eng.load_table(t.id, tab)
    out = eng.gather_only(oracle.idx_reference(96, cat.n_tables))
    assert_bits_equal(out[:32], g["wire_first32"])
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["wire96_sha256"])
    eng.close()


@pytest.mark.parametrize("model", ("small", "medium", "large_half", "large"))
@pytest.mark.parametrize("B", (1, 37, 1000))
def test_gather_bit_exact_vs_oracle(model, B):
    cat = catalogue.load(model).with_row_cap(3000)
    eng = fleetrec.Engine(cat, max_batch=1024)
    tables = oracle.make_tables(cat, "hash", seed=11)
    eng.load_tables(tables)
    idx = oracle.uniform_indices(cat, B, seed=B)
    assert_bits_equal(eng.gather_only(idx), oracle.gather(cat, tables, idx))
    eng.close()


def test_gather_edge_cases_and_device_fills():
    cat = catalogue.load("small").with_row_cap(257)          # odd row count
    eng = fleetrec.Engine(cat, max_batch=64)
    eng.fill_hash(seed=5)
    for t in (0, 29, 46):
        assert_bits_equal(eng.read_table(t, 0, cat.tables[t].rows),
                          oracle.fill_hash(5, t, cat.tables[t].rows, cat.tables[t].dim))
    assert eng.gather_only(np.zeros((0, 47), np.int32)).shape == (0, 352)          # empty batch
    idx = np.stack([np.array([t.rows - 1 for t in cat.tables], np.int32), np.zeros(47, np.int32)])
    assert_bits_equal(eng.gather_only(idx), oracle.gather_hashed(cat, 5, idx))       # first / last rows
    eng.close()
    # reference fill + reference index list: the lookup KAT (all-ones / all-zeros items)
    for model in MODELS3:
        cat = catalogue.load(model).with_row_cap(201)
        eng = fleetrec.Engine(cat, max_batch=64)
        eng.fill_reference()
        out = eng.gather_only(oracle.idx_reference(32, cat.n_tables))
        for j, r in enumerate(catalogue.IDX_RANDOM):
            assert np.all(out[j] == (1.0 if r % 2 == 0 else 0.0)), (model, j)
        assert_bits_equal(eng.read_table(1, 0, 201), oracle.fill_reference(201, cat.tables[1].dim))
        eng.close()
    cat = catalogue.load("small").with_row_cap(1000)
    eng = fleetrec.Engine(cat, max_batch=64)
    eng.fill_reference(debug_rows=200)                        # host.cpp:75 `#define DEBUG`
    assert_bits_equal(eng.read_table(0, 0, 1000), oracle.fill_reference(1000, cat.tables[0].dim, 200))
    eng.close()


def test_gather_full_size_small_model():
    """BASELINE configs[1] tables at their real sizes (1.4 GB, 10 M-row DDR1): the
    expected bytes come from the position-encoding hash, no host tables needed."""
    cat = catalogue.load("small")
    eng = fleetrec.Engine(cat, max_batch=4096)
    eng.fill_hash(seed=0x5EED)
    assert eng.table_bytes() == cat.table_bytes()
    for idx in (oracle.zipf_indices(cat, 2048), oracle.uniform_indices(cat, 4096),
                np.array([[t.rows - 1 for t in cat.tables]], np.int32)):
        assert_bits_equal(eng.gather_only(idx), oracle.gather_hashed(cat, 0x5EED, idx))
    eng.close()


def test_errors_are_reported_not_fatal():
    cat = catalogue.load("small").with_row_cap(100)
    eng = fleetrec.Engine(cat, max_batch=32)
    with pytest.raises(fleetrec.FleetRecError) as ei:
        eng.gather_only(np.zeros((4, 47), np.int32))          # tables not loaded
    assert ei.value.code == 4 and "not loaded" in str(ei.value)
    eng.fill_hash()
    with pytest.raises(fleetrec.FleetRecError):
        eng.gather_only(np.zeros((33, 47), np.int32))         # B > max_batch
    with pytest.raises(fleetrec.FleetRecError) as ei:
        eng.infer(np.zeros((4, 47), np.int32))                # MLP not loaded
    assert "layer" in str(ei.value)
    with pytest.raises(fleetrec.FleetRecError):
        eng.load_table(0, np.zeros((100, 12), np.float32))    # wrong dim
    eng.close()


# ---------------------------------------------------------------- MLP
KAT = {"small": 47244640256.0, "medium": 118111600640.0, "large": 1065151889408.0}


@pytest.mark.parametrize("prec,tiles", ((fleetrec.FR_PREC_FP32, ""), (fleetrec.FR_PREC_TF32, "auto"), (fleetrec.FR_PREC_TF32, TILES[0]),
                                        (fleetrec.FR_PREC_TF32, TILES[1]), (fleetrec.FR_PREC_TF32, TILES[2]),
                                        (fleetrec.FR_PREC_TF32, TILES[3]), (fleetrec.FR_PREC_TF32, TILES[4])))
@pytest.mark.parametrize("model", ("small", "medium", "large"))
def test_mlp_all_ones_known_answer(model, prec, tiles, monkeypatch):
    """README.md:7-11: all-ones input and weights -> IN*H1*H2*H3, exact in fp32 and tf32."""
    if tiles == "auto":
        monkeypatch.delenv("FR_TC_TILES", raising=False)
    else:
        monkeypatch.setenv("FR_TC_TILES", tiles or TILES[1])
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, precision=prec, max_batch=256)
    eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
    out = eng.mlp_only(np.ones((130, dims[0]), np.float32))
    assert np.all(out == np.float32(KAT[model])), out[:4]
    eng.close()


@pytest.mark.parametrize("prec,tol", ((fleetrec.FR_PREC_FP32, 2e-5), (fleetrec.FR_PREC_TF32, TOL)))
@pytest.mark.parametrize("model", ("small", "medium", "large"))
@pytest.mark.parametrize("B", (1, 100, 2048, 4099))
def test_mlp_bias_relu_sigmoid_vs_oracle(model, prec, tol, B):
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    W, b = oracle.make_weights(dims, seed=42)
    x = np.random.default_rng(B).uniform(-1, 1, (B, dims[0])).astype(np.float32)
    eng = fleetrec.Engine(cat, precision=prec, max_batch=4224)
    eng.load_mlp(W, b)
    got = eng.mlp_only(x)
    exp = oracle.mlp(x, dims, W, b, mode=1)
    assert rel_err(got, exp) <= tol, rel_err(got, exp)
    eng.close()


@pytest.mark.parametrize("prec,tol", ((fleetrec.FR_PREC_FP32, 2e-5), (fleetrec.FR_PREC_TF32, TOL)))
def test_mlp_linear_mode_vs_oracle(prec, tol):
    """Reference-exact mode (no bias / activation).  Positive weights and inputs keep
    every output far from zero, so the element-wise relative error is well defined;
    random-sign weights are checked against the output scale instead."""
    cat = catalogue.load("small").with_row_cap(64)
    dims = cat.layer_dims
    rng = np.random.default_rng(3)
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, precision=prec, max_batch=1024)
    Wp = [(rng.uniform(0, 2, (dims[k], dims[k + 1])) / dims[k]).astype(np.float32) for k in range(4)]
    x = rng.uniform(0, 1, (777, dims[0])).astype(np.float32)
    eng.load_mlp(Wp)
    exp = oracle.mlp(x, dims, Wp, None, mode=0)
    assert rel_err(eng.mlp_only(x), exp) <= tol
    W, _ = oracle.make_weights(dims, seed=9)
    x = rng.uniform(-1, 1, (777, dims[0])).astype(np.float32)
    eng.load_mlp(W)
    exp = oracle.mlp(x, dims, W, None, mode=0, acc64=True)
    assert np.max(np.abs(eng.mlp_only(x) - exp)) / np.std(exp) <= 5 * tol
    eng.close()


# ---------------------------------------------------------------- end to end
@pytest.mark.parametrize("prec,tol", ((fleetrec.FR_PREC_FP32, 2e-5), (fleetrec.FR_PREC_TF32, TOL)))
@pytest.mark.parametrize("model", ("small", "medium"))
def test_infer_end_to_end(model, prec, tol):
    cat = catalogue.load(model).with_row_cap(20000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=77)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, precision=prec, max_batch=2048)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    workers = [fleetrec.Worker(eng) for _ in range(2)]
    for B, w in ((2048, None), (256, workers[0]), (333, workers[1])):
        idx = oracle.zipf_indices(cat, B, seed=B)
        exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
        got = eng.infer(idx, worker=w)
        assert rel_err(got, exp) <= tol, (B, rel_err(got, exp))
    assert eng.launch_count() > 0
    for w in workers:
        w.close()
    eng.close()


def test_reference_chain_kat_through_infer():
    """reference fill + index list + all-ones LINEAR MLP -> KAT value or 0."""
    cat = catalogue.load("small").with_row_cap(200)
    dims = cat.layer_dims
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, max_batch=64)
    eng.fill_reference()
    eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
    out = eng.infer(oracle.idx_reference(32, 47))
    for j, r in enumerate(catalogue.IDX_RANDOM):
        assert out[j] == (np.float32(KAT["small"]) if r % 2 == 0 else 0.0)
    eng.close()


def test_cartesian_merge_on_device():
    """gather(M, remap(iA,iB)) == gather(A,iA) || gather(B,iB)  (SURVEY 8c-b)."""
    T = catalogue.Table
    tabs = [T(0, "HBM", 0, 0, 0, 37, 4, 1, 0), T(1, "HBM", 1, 1, 0, 11, 8, 2, 0), T(2, "HBM", 2, 2, 0, 37 * 11, 12, 3, 0),
            T(3, "HBM", 3, 3, 0, 8, 8, 2, 0)]
    S = catalogue.Segment
    cat = catalogue.Model("merge", tabs, [S(0, 0, 0, 4), S(4, 1, 0, 8), S(12, 2, 0, 12), S(24, 3, 0, 8)], 32, 32,
                          [128, 128, 256, 1])
    eng = fleetrec.Engine(cat, max_batch=512)
    for t in (0, 1, 3):
        eng._chk(eng._L.fr_fill_table_hash(eng._h, t, 9))
    eng.merge_tables(0, 1, 2)
    assert_bits_equal(eng.read_table(2, 0, 37 * 11),
                      oracle.merge_tables(oracle.fill_hash(9, 0, 37, 4), oracle.fill_hash(9, 1, 11, 8)))
    ia, ib = np.meshgrid(np.arange(37), np.arange(11), indexing="ij")
    idx = np.zeros((37 * 11, 4), np.int32)
    idx[:, 0], idx[:, 1] = ia.ravel(), ib.ravel()
    idx[:, 2] = [fleetrec.merge_index(a, b, 11) for a, b in zip(ia.ravel(), ib.ravel())]
    out = eng.gather_only(idx)
    assert_bits_equal(out[:, 0:12], out[:, 12:24])
    eng.close()


def tf32_rna(x):
    b = x.view(np.uint32).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


@experimental
@pytest.mark.parametrize("tiles", TILES_1CTA)
@pytest.mark.parametrize("k", (0, 1, 2))
def test_tf32_single_layer_one_cta_tiles(k, tiles, monkeypatch):
    """The older kernel's cta_group::1 tiles (experiments build)."""
    test_tf32_single_layer_vs_numpy(k, 300, tiles, "medium", monkeypatch)


@pytest.mark.parametrize("tiles", TILES)
@pytest.mark.parametrize("k", (0, 1, 2))
@pytest.mark.parametrize("B", (128, 300, 1100))
@pytest.mark.parametrize("model", ("medium", "small"))
def test_tf32_single_layer_vs_numpy(k, B, tiles, model, monkeypatch):
    """Each tcgen05 GEMM configuration alone (128-, 256-, 512-wide pair tiles, layer 3 with the folded output layer):
    inputs pre-rounded to TF32, so the only difference to float64 is summation order.  medium: K = 880 = 27.5 K slices
    (one slice per TMA box, tail zero-filled inside the slice); small: K = 352 = 11 slices (two slices per box, the
    12th zero-filled and skipped); hidden layers: 32 and 16 slices; B = 1100: several M tiles and an M tail."""
    monkeypatch.setenv("FR_TC_TILES", tiles)                   # tile width per layer, CTAs per tile
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    W, b = oracle.make_weights(dims, seed=5)
    eng = fleetrec.Engine(cat, max_batch=2048)
    eng.load_mlp(W, b)
    x = tf32_rna(np.random.default_rng(k).uniform(-1, 1, (B, dims[k])).astype(np.float32))
    h = np.maximum(x.astype(np.float64) @ tf32_rna(W[k]).astype(np.float64) + b[k], 0)
    if k < 2:
        got = eng.layer_only(k, x, dims[k + 1])
        assert got.shape == h.shape
        assert np.max(np.abs(got - h)) <= 2e-3 * max(1.0, np.abs(h).max())
        assert np.array_equal(got.view(np.uint32) & 0x1FFF, np.zeros_like(got, np.uint32))   # stored tf32-rounded
    else:
        s = 1 / (1 + np.exp(-(h @ W[3].astype(np.float64)[:, 0] + b[3][0])))
        got = eng.layer_only(k, x, 1)
        assert rel_err(got, s.astype(np.float32)) <= 2e-4
    eng.close()


def test_infer_graph_replay_matches_direct():
    """fr_infer replays a CUDA graph from the 2nd call on for the same (idx, scores) buffers:
    new index CONTENTS in the same device / pinned buffers must give new, correct scores."""
    import torch
    cat = catalogue.load("small").with_row_cap(20000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=3)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=512)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    w = fleetrec.Worker(eng)
    B = 512
    for kind in ("device", "pinned"):
        idx_buf = torch.empty((B, 47), dtype=torch.int32, device="cuda") if kind == "device" else \
            torch.empty((B, 47), dtype=torch.int32).pin_memory()
        sc_buf = torch.empty(B, dtype=torch.float32, device="cuda") if kind == "device" else \
            torch.empty(B, dtype=torch.float32).pin_memory()
        l0 = eng.launch_count()
        for it in range(4):
            idx = oracle.zipf_indices(cat, B, seed=100 + it)
            idx_buf.copy_(torch.from_numpy(idx))
            torch.cuda.synchronize()
            eng.infer_async(idx_buf if kind == "device" else idx_buf.numpy(),
                            sc_buf if kind == "device" else sc_buf.numpy(), B, w)
            eng.sync(w)
            got = sc_buf.cpu().numpy() if kind == "device" else sc_buf.numpy().copy()
            exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
            assert rel_err(got, exp) <= TOL, (kind, it)
        # lookup + 3 GEMM launches per batch, replayed or not (experiments build with FR_ZEROCOPY: + a staging kernel)
        per = 4 + (1 if kind == "pinned" and _has_experiments() and os.environ.get("FR_ZEROCOPY", "0") != "0" else 0)
        assert eng.launch_count() - l0 == 4 * per
    w.close()
    eng.close()


@experimental
@pytest.mark.parametrize("model,B,clusters", (("small", 513, 0), ("small", 2048, 0), ("small", 1300, 1),
                                              ("medium", 4099, 2), ("large", 2304, 3), ("small", 40000, 0)))
def test_tf32_cp_async_a_operand_bit_identical_to_tma(model, B, clusters, monkeypatch):
    """Throughput-sized batches feed the A operand through the LSU: four loader warps copy every 128 x 32
    activation slice with 16-byte cp.async into the 128B-swizzled slot TMA would have filled (zero fill past
    the batch and past K: medium has K = 880 = 27.5 slices), post one arrival per warp on the pair leader's
    full barrier after fence.proxy.async, and TMA carries the weights only.  Same MMAs, same order: the scores
    must be identical bit for bit to FR_TC_ALSU=0 (everything through TMA); capped grids wrap the ring."""
    if clusters:
        monkeypatch.setenv("FR_TC_MAX_CLUSTERS", str(clusters))
    else:
        monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    W, b = oracle.make_weights(dims, seed=33)
    x = np.random.default_rng(B).uniform(-1, 1, (B, dims[0])).astype(np.float32)
    got = {}
    for v in ("1", "0"):
        monkeypatch.setenv("FR_TC_ALSU", v)
        eng = fleetrec.Engine(cat, max_batch=B)
        eng.load_mlp(W, b)
        for _ in range(3):
            got[v] = eng.mlp_only(x)
        eng.close()
    monkeypatch.delenv("FR_TC_ALSU", raising=False)
    monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    # (the variant runs the older kernel, the default the production one: same operands, same K order per MMA, another
    # order of the final 256-term dot product of the folded output layer)
    assert rel_err(got["1"], got["0"]) <= 2e-6
    assert rel_err(got["1"], oracle.mlp(x, dims, W, b, mode=1)) <= TOL


@experimental
@pytest.mark.parametrize("model,B,clusters", (("small", 513, 0), ("small", 2048, 0), ("small", 1300, 1),
                                              ("medium", 4099, 2), ("large", 2304, 3), ("small", 40000, 0)))
def test_tf32_multicast_clusters_bit_identical_to_pair_clusters(model, B, clusters, monkeypatch):
    """FR_TC_MCAST=1: throughput-sized batches run in 4-CTA clusters: two MMA pairs on adjacent 256-row tiles share every
    weight slice by TMA multicast (each CTA loads half of its share and multicasts it to the CTA of its
    parity in the other pair; a smem slot is refilled only after BOTH pairs released it).  Same MMAs in the
    same K order as the 2-CTA clusters, so the scores must be identical bit for bit; odd numbers of 256-row
    tiles leave the second pair of the last cluster with no rows (TMA zero fill / clipped stores); capped
    grids make the barrier parities wrap."""
    if clusters:
        monkeypatch.setenv("FR_TC_MAX_CLUSTERS", str(clusters))
    else:
        monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    W, b = oracle.make_weights(dims, seed=31)
    x = np.random.default_rng(B).uniform(-1, 1, (B, dims[0])).astype(np.float32)
    got = {}
    for mc in ("1", "0"):
        monkeypatch.setenv("FR_TC_MCAST", mc)
        eng = fleetrec.Engine(cat, max_batch=B)
        eng.load_mlp(W, b)
        for _ in range(3):
            got[mc] = eng.mlp_only(x)
        eng.close()
    monkeypatch.delenv("FR_TC_MCAST", raising=False)
    monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    # (the variant runs the older kernel, the default the production one: same operands, same K order per MMA, another
    # order of the final 256-term dot product of the folded output layer)
    assert rel_err(got["1"], got["0"]) <= 2e-6
    assert rel_err(got["1"], oracle.mlp(x, dims, W, b, mode=1)) <= TOL


@experimental
@pytest.mark.parametrize("B", (1, 333, 2048))
def test_pinned_buffers_without_copy_engine_match_memcpy_path(B, monkeypatch):
    """Page-locked caller buffers: indices are fetched over PCIe by a staging kernel and the scores are written
    to the host buffer by the last MLP kernel itself (no memcpy nodes) with FR_ZEROCOPY=1; 0 = cudaMemcpyAsync.
    Both must give the same bits, direct and graph-replayed, with new index contents in the same buffer
    (B * 47 * 4 bytes is a multiple of 16 only for some B: the others take the memcpy path for the indices)."""
    import torch
    cat = catalogue.load("small").with_row_cap(5000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=5)
    W, b = oracle.make_weights(dims, seed=42)
    got = {}
    for zc in ("1", "0", "55"):     # all over PCIe by the SMs | all by the copy engine | split
        monkeypatch.setenv("FR_ZEROCOPY", zc)
        eng = fleetrec.Engine(cat, max_batch=B)
        eng.load_tables(tables)
        eng.load_mlp(W, b)
        w = fleetrec.Worker(eng)
        idx_buf = torch.empty((B, 47), dtype=torch.int32).pin_memory()
        sc_buf = torch.zeros(B, dtype=torch.float32).pin_memory()
        outs = []
        for it in range(4):     # calls 2.. replay the captured graph
            idx = oracle.zipf_indices(cat, B, seed=40 + it)
            idx_buf.copy_(torch.from_numpy(idx))
            eng.infer_async(idx_buf.numpy(), sc_buf.numpy(), B, w)
            eng.sync(w)
            outs.append(sc_buf.numpy().copy())
            if zc == "1":
                exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
                assert rel_err(outs[-1], exp) <= TOL, it
        got[zc] = outs
        w.close()
        eng.close()
    monkeypatch.delenv("FR_ZEROCOPY", raising=False)
    for a, c, d in zip(got["1"], got["0"], got["55"]):
        assert_bits_equal(a, c)
        assert_bits_equal(d, c)


@pytest.mark.parametrize("tiles", TILES)
@pytest.mark.parametrize("clusters,pdl", ((1, "1"), (3, "0"), (0, "1")))
def test_tf32_persistent_tile_loop(tiles, clusters, pdl, monkeypatch):
    """The tcgen05 kernels are persistent: a cluster walks tiles c, c+G, ...  with two TMEM
    accumulator stages and two TMA-store buffers per epilogue warp.  Capping the grid
    (FR_TC_MAX_CLUSTERS) makes every cluster run many tiles, so ring positions, accumulator
    stage parities and store-buffer reuse all wrap several times; clusters=0 is the production
    grid at a batch with > 148 tiles.  Checked per layer against float64 on TF32-rounded
    inputs and through the whole chain against the oracle, with and without PDL."""
    monkeypatch.setenv("FR_TC_TILES", tiles)
    monkeypatch.setenv("FR_PDL", pdl)
    if clusters:
        monkeypatch.setenv("FR_TC_MAX_CLUSTERS", str(clusters))
    else:
        monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    B = 1300 if clusters else 20000
    cat = catalogue.load("small").with_row_cap(64)
    dims = cat.layer_dims
    W, b = oracle.make_weights(dims, seed=11)
    eng = fleetrec.Engine(cat, max_batch=B)
    eng.load_mlp(W, b)
    rng = np.random.default_rng(clusters)
    for k in (0, 1):
        x = tf32_rna(rng.uniform(-1, 1, (B, dims[k])).astype(np.float32))
        h = np.maximum(x.astype(np.float64) @ tf32_rna(W[k]).astype(np.float64) + b[k], 0)
        got = eng.layer_only(k, x, dims[k + 1])
        assert np.max(np.abs(got - h)) <= 2e-3 * max(1.0, np.abs(h).max()), k
    x = rng.uniform(-1, 1, (B, dims[0])).astype(np.float32)
    for _ in range(2):   # twice: the second chain runs back to back with the first on the same stream
        got = eng.mlp_only(x)
    assert rel_err(got, oracle.mlp(x, dims, W, b, mode=1)) <= TOL
    eng.close()
    monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)


@experimental
@pytest.mark.parametrize("model,B,clusters", (("small", 513, 0), ("small", 2048, 0), ("small", 1300, 1),
                                              ("medium", 3000, 2), ("large", 2304, 3), ("small", 40000, 0)))
def test_tf32_chain_kernel_bit_identical_to_per_layer_kernels(model, B, clusters, monkeypatch):
    """FR_CHAIN=1: batches above 512 run the whole MLP as ONE persistent launch (tc_mlp_chain_kernel): a CTA pair
    walks its 256 items through every layer, H1 / H2 stored with TMA and re-loaded by the same CTA behind
    `ready` barriers.  The K order of every accumulation is that of the per-layer kernels, so the scores
    must be IDENTICAL bit for bit (a stale H1 / H2 row, i.e. a broken store -> load hand-over, could not
    be); capped grids make one pair run many item tiles (barrier parities wrap), B = 40000 is the
    production grid with more item tiles than CTA pairs.  Also against the oracle and the README KAT."""
    if clusters:
        monkeypatch.setenv("FR_TC_MAX_CLUSTERS", str(clusters))
    else:
        monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    W, b = oracle.make_weights(dims, seed=21)
    x = np.random.default_rng(B).uniform(-1, 1, (B, dims[0])).astype(np.float32)
    got = {}
    for chain in ("1", "0"):
        monkeypatch.setenv("FR_CHAIN", chain)
        eng = fleetrec.Engine(cat, max_batch=B)
        eng.load_mlp(W, b)
        l0 = eng.launch_count()
        for _ in range(3):   # back to back on one stream: the next launch overwrites H1 / H2 of the previous one
            got[chain] = eng.mlp_only(x)
        assert eng.launch_count() - l0 == 3 * (1 + (1 if chain == "1" else 3))   # operand rounding + chain | 3 layers
        if chain == "1":     # reference KAT through the chain (LINEAR mode, all-ones): exact
            eng.close()
            eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, max_batch=B)
            eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
            out = eng.mlp_only(np.ones((min(B, 1000), dims[0]), np.float32))
            assert np.all(out == np.float32(KAT[model])), out[:4]
        eng.close()
    monkeypatch.delenv("FR_CHAIN", raising=False)
    monkeypatch.delenv("FR_TC_MAX_CLUSTERS", raising=False)
    # (the variant runs the older kernel, the default the production one: same operands, same K order per MMA, another
    # order of the final 256-term dot product of the folded output layer)
    assert rel_err(got["1"], got["0"]) <= 2e-6
    assert rel_err(got["1"], oracle.mlp(x, dims, W, b, mode=1)) <= TOL


def test_merge_planner_engine_matches_unmerged_engine():
    """plan merges under a byte budget, build the merged tables on the device
    (fr_merge_tables), drive the merged engine with remapped indices: concat vectors bit-identical
    to the un-merged engine's and to the oracle's, scores within tolerance, fewer index columns read."""
    from fleetrec import merge
    cat = catalogue.load("small").with_row_cap(2000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=31)
    W, b = oracle.make_weights(dims, seed=42)
    plan = merge.plan_merges(cat, 256 << 20)
    assert plan.lookups_saved >= 4
    mm = merge.apply_merges(cat, plan.pairs)
    eng = merge.build(mm, tables, max_batch=512)
    eng.load_mlp(W, b)
    idx = oracle.zipf_indices(cat, 500, seed=3)
    idx[0, :] = 0
    idx[1, :] = [t.rows - 1 for t in cat.tables]
    exp = oracle.gather(cat, tables, idx)
    got = eng.gather_only(mm.remap(idx))
    assert_bits_equal(got, exp)
    assert rel_err(eng.infer(mm.remap(idx)), oracle.mlp(exp, dims, W, b, mode=1)) <= TOL
    eng.close()


# ---------------------------------------------------------------- request-driven batching front-end (8f-2)
def test_batcher_many_producers_match_oracle():
    """fr_batcher: requests of random sizes from several threads are packed into batches (full or
    closed by the deadline), scored on worker streams, and every request gets exactly its own
    scores back -- against the oracle on the union of all requests."""
    import threading
    cat = catalogue.load("small").with_row_cap(20000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=13)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=512)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    bat = fleetrec.Batcher(eng, max_batch=512, max_delay_us=500, n_workers=3)
    n_threads, per_thread = 4, 40
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 300, (n_threads, per_thread))
    sizes[0, 0] = 1500                                         # larger than max_batch: split over batches
    reqs = [[oracle.zipf_indices(cat, int(sizes[t, i]), seed=1000 * t + i) for i in range(per_thread)]
            for t in range(n_threads)]
    outs = [[np.full(int(sizes[t, i]), np.nan, np.float32) for i in range(per_thread)] for t in range(n_threads)]
    errors = []

    def producer(t):
        try:
            tickets = [bat.submit(reqs[t][i], outs[t][i]) for i in range(per_thread)]
            for tk in tickets:
                bat.wait(tk)
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)
    threads = [threading.Thread(target=producer, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not errors, errors
    allidx = np.concatenate([r for rs in reqs for r in rs])
    exp = oracle.mlp(oracle.gather(cat, tables, allidx), dims, W, b, mode=1)
    got = np.concatenate([o for os_ in outs for o in os_])
    assert not np.isnan(got).any()
    assert rel_err(got, exp) <= TOL
    st = bat.stats()
    assert st["items"] == allidx.shape[0] and st["requests"] >= n_threads * per_thread
    assert st["batches"] >= allidx.shape[0] // 512 and st["latency_p99_us"] >= st["latency_p50_us"] > 0
    bat.close()
    eng.close()


def test_batcher_deadline_flush_and_errors():
    import time
    cat = catalogue.load("small").with_row_cap(5000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=14)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=4096)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    with pytest.raises(fleetrec.FleetRecError):
        fleetrec.Batcher(eng, max_batch=8192)                  # above the engine's max_batch
    bat = fleetrec.Batcher(eng, max_batch=4096, max_delay_us=3000, n_workers=2)
    idx = oracle.zipf_indices(cat, 7, seed=1)
    out = np.zeros(7, np.float32)
    t0 = time.perf_counter()
    bat.wait(bat.submit(idx, out))                             # 7 of 4096 items: only the deadline closes it
    waited = time.perf_counter() - t0
    assert 0.002 <= waited < 2.0, waited
    exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
    assert rel_err(out, exp) <= TOL
    assert bat.stats()["closed_by_deadline"] == 1
    bat.close()
    bat = fleetrec.Batcher(eng, max_batch=4096, max_delay_us=10_000_000, n_workers=1)
    out2 = np.zeros(7, np.float32)
    tk = bat.submit(idx, out2)
    bat.flush()                                                # explicit flush beats a 10 s deadline
    t0 = time.perf_counter()
    bat.wait(tk)
    assert time.perf_counter() - t0 < 2.0
    assert np.array_equal(out2, out)
    with pytest.raises(fleetrec.FleetRecError):
        bat.wait(10 ** 9)                                      # unknown ticket
    out3 = np.zeros(5, np.float32)
    bat.submit(oracle.zipf_indices(cat, 5, seed=2), out3)
    bat.close()                                                # destroy drains what was submitted
    assert np.all(out3 > 0)
    eng.close()


# ---------------------------------------------------------------- lookup fused into layer 1
@pytest.mark.parametrize("model,B", (("small", 1), ("small", 300), ("small", 2048), ("small", 20000),
                                     ("medium", 777), ("large", 515)))
def test_fused_lookup_layer1_matches_unfused_and_oracle(model, B, monkeypatch):
    """With FR_OPT_FUSE_LOOKUP fr_infer's TF32 chain gathers straight into the first GEMM's A tile (no concat
    in global memory).  Same indices through (a) the fused chain, (b) the default (lookup kernel, concat
    materialised, TMA-fed layer 1; the default) and (c) the oracle: (a) vs (c) within the MLP tolerance, (a) vs
    (b) equal to the last bit or two (same TF32-rounded operands, same K order, fp32 accumulate).
    Covers K tails (medium: 27.5 slices), M tails, > 1 tile per cluster (B = 20000) and N = 2048."""
    cat = catalogue.load(model).with_row_cap(3000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=17)
    W, b = oracle.make_weights(dims, seed=42)
    idx = oracle.zipf_indices(cat, B, seed=B)
    idx[0, :] = [t.rows - 1 for t in cat.tables]
    got = {}
    for fuse in ("1", "0"):
        eng = fleetrec.Engine(cat, max_batch=B)
        eng.set_option(fleetrec.FR_OPT_FUSE_LOOKUP, int(fuse))
        eng.load_tables(tables)
        eng.load_mlp(W, b)
        l0 = eng.launch_count()
        got[fuse] = eng.infer(idx)
        # lookup fused into layer 1 + layers 2, 3; else lookup + (one chain launch above 512 items | 3 layers)
        chain = os.environ.get("FR_CHAIN", "0") == "1" and B > 512
        assert eng.launch_count() - l0 == (3 if fuse == "1" else (2 if chain else 4))
        eng.close()
    exp = oracle.mlp(oracle.gather(cat, tables, idx), dims, W, b, mode=1)
    assert rel_err(got["1"], exp) <= TOL, rel_err(got["1"], exp)
    assert np.max(np.abs(got["1"] - got["0"])) <= 2e-6, np.max(np.abs(got["1"] - got["0"]))


def test_fused_chain_reference_kat(monkeypatch):
    """The reference's own known answer through the fused chain: T1 fill + I1 indices + all-ones
    weights, LINEAR mode -> IN*H1*H2*H3 for even rows, 0 for odd (README.md:7-11)."""
    cat = catalogue.load("small").with_row_cap(200)
    dims = cat.layer_dims
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, max_batch=64)
    eng.set_option(fleetrec.FR_OPT_FUSE_LOOKUP, 1)
    eng.fill_reference()
    eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
    idx = oracle.idx_reference(64, cat.n_tables)
    out = eng.infer(idx)
    exp = np.where(idx[:, 0] % 2 == 0, np.float32(KAT["small"]), np.float32(0))
    assert np.array_equal(out, exp)
    eng.close()


# ---------------------------------------------------------------- reduced-precision table storage (8f-4)
@pytest.mark.parametrize("dt", (fleetrec.FR_TABLE_F16, fleetrec.FR_TABLE_BF16, fleetrec.FR_TABLE_FP8))
def test_reduced_precision_tables_bit_exact_after_stated_dequant(dt):
    """Tables stored as f16 / bf16 / fp8 (E4M3, saturating) in HBM: rows uploaded as fp32 are converted on the device
    (RNE), device fills produce the same values, the lookup widens exactly -- concat == float32(round(rows)) bit for
    bit, scores within the MLP tolerance of the oracle run on the dequantised concat.  Table bytes halve / quarter.
    Edge values (ties, saturation, underflow, signed zero) sit in the first rows of one table."""
    cat = catalogue.load("medium").with_row_cap(3000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=0x5EED)
    edge = np.array([0.0, -0.0, 1.0625, 1.1875, 17.0, 19.0, 448.0, 449.0, 480.0, 1e9, -1e9, 2.0 ** -9, 2.0 ** -10, 1.0001 * 2.0 ** -10,
                     0.00097, 65504.0, 65520.0, 3.0e38, 6.0e-8, 2.0 ** -24, 2.0 ** -25, 1.0009765625, 1.00048828125, -7.0e4], np.float32)
    edge = np.resize(edge, tables[5][:8].size).reshape(tables[5][:8].shape)
    if dt == fleetrec.FR_TABLE_F16:
        edge = np.clip(edge, -65504.0, 65504.0)     # (fp16 storage does not saturate: keep the edge rows finite)
    plain5 = tables[5].copy()
    tables[5][:8] = edge
    W, b = oracle.make_weights(dims, seed=42)
    idx = oracle.zipf_indices(cat, 700, seed=2)
    idx[0, :] = [t.rows - 1 for t in cat.tables]
    exp = oracle.quantize_dequantize(oracle.gather(cat, tables, idx), dt)
    eng = fleetrec.Engine(cat, max_batch=1024, table_dtype=dt)
    eng.load_tables(tables)                                     # fp32 in, converted on the device
    assert eng.table_bytes() == cat.table_bytes() // (4 if dt == fleetrec.FR_TABLE_FP8 else 2)
    assert_bits_equal(eng.read_table(5, 0, 8), oracle.quantize_dequantize(tables[5][:8], dt))
    assert_bits_equal(eng.gather_only(idx), exp)
    assert_bits_equal(eng.read_table(5, 10, 20), oracle.quantize_dequantize(tables[5][10:30], dt))
    eng.load_mlp(W, b)
    idx2 = idx.copy()
    idx2[:, 5] = np.maximum(idx2[:, 5], 8)                      # (the MLP check stays off the edge rows: 1e9, 3e38 ...)
    exp2 = oracle.quantize_dequantize(oracle.gather(cat, tables, idx2), dt)
    assert rel_err(eng.infer(idx2), oracle.mlp(exp2, dims, W, b, mode=1)) <= TOL
    eng.close()
    tables[5] = plain5
    exp = oracle.quantize_dequantize(oracle.gather(cat, tables, idx), dt)
    eng = fleetrec.Engine(cat, max_batch=1024, table_dtype=dt)
    eng.fill_hash(seed=0x5EED)                                  # device-side fill: same values
    assert_bits_equal(eng.gather_only(idx), exp)
    eng.close()
    eng = fleetrec.Engine(cat, max_batch=64, table_dtype=dt)
    eng.fill_reference()                                        # 1.0 / 0.0 are exact in both types
    i1 = oracle.idx_reference(32, cat.n_tables)
    assert_bits_equal(eng.gather_only(i1), oracle.gather(cat, oracle.make_tables(cat, "reference"), i1))
    with pytest.raises(fleetrec.FleetRecError):
        eng.merge_tables(0, 1, 2)                               # merged images are built in fp32 only
    eng.close()


# ---------------------------------------------------------------- B2-compatible TCP ingest (8f-3)
def _send_blocks(port, blocks):
    """What multiple_connections_network_client_sender.c:55-100 does: connect, then send() raw little-endian bytes block
    after block, no header (tests/wire.py; tests/test_wire_format.py checks it against the reference's own sender)."""
    import wire
    wire.send_blocks(port, blocks)


def _free_base_port(n):
    import wire
    base = wire.free_base_port(n)
    if base is None:
        pytest.skip("no free port range")
    return base


@pytest.mark.parametrize("payload", ("concat", "indices"))
def test_ingest_accepts_the_reference_wire_format(payload):
    """Senders speaking the reference's B2 format (raw fp32 [item][INPUT_SIZE] on PORT+i, or raw int32
    index rows) against fr_ingest on loopback: every block of every connection is scored, in order,
    against the oracle; batch numbers come off one shared counter."""
    import threading
    cat = catalogue.load("small").with_row_cap(20000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=19)
    W, b = oracle.make_weights(dims, seed=42)
    eng = fleetrec.Engine(cat, max_batch=128)
    eng.load_tables(tables)
    eng.load_mlp(W, b)
    n_conn, per_conn, B = 3, 5, 128
    base = _free_base_port(n_conn)
    ing = fleetrec.Ingest(eng, base, n_conn, B, payload=payload, max_batches_per_conn=per_conn)
    idx = [[oracle.zipf_indices(cat, B, seed=100 * c + k) for k in range(per_conn)] for c in range(n_conn)]
    x = [[oracle.gather(cat, tables, idx[c][k]) for k in range(per_conn)] for c in range(n_conn)]
    blocks = x if payload == "concat" else idx
    th = [threading.Thread(target=_send_blocks, args=(base + c, blocks[c])) for c in range(n_conn)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=60)
    st = ing.wait()
    assert st["batches"] == n_conn * per_conn and st["connections"] == n_conn
    assert st["bytes"] == n_conn * per_conn * blocks[0][0].nbytes
    for c in range(n_conn):
        for k in range(per_conn):
            exp = oracle.mlp(x[c][k], dims, W, b, mode=1)
            assert rel_err(ing.scores[c, k], exp) <= TOL, (c, k)
        last, no = ing.last_scores(c)
        # a connection draws its batch number BEFORE it reads the block (cuda_server.c:408-417), so a connection
        # waiting for its sender's EOF holds a number no block will use: real blocks can be numbered up to
        # n_conn * per_conn + (n_conn - 1)
        assert np.array_equal(last, ing.scores[c, per_conn - 1]) and 0 <= no < n_conn * (per_conn + 1)
    ing.close()
    eng.close()


def test_ingest_total_batches_and_truncated_stream():
    import threading
    cat = catalogue.load("small").with_row_cap(64)
    dims = cat.layer_dims
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, max_batch=32)
    eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
    base = _free_base_port(2)
    # TOTAL_BATCH_NUM = 3 over two connections: the shared counter stops the servers (cuda_server.c:408-412)
    ing = fleetrec.Ingest(eng, base, 2, 32, total_batches=3, max_batches_per_conn=4)
    ones = np.ones((32, dims[0]), np.float32)                 # the reference's own traffic: all-ones vectors
    th = [threading.Thread(target=_send_blocks, args=(base + c, [ones, ones])) for c in range(2)]
    for t in th:
        t.start()
    st = ing.wait()
    for t in th:
        t.join(timeout=30)
    assert st["batches"] == 3
    assert np.all(ing.scores[0, 0] == np.float32(KAT["small"]))                       # README.md:7-11 known answer
    ing.close()
    # a sender that dies inside a block is an error, not a hang and not a silently short batch
    base = _free_base_port(1)
    ing = fleetrec.Ingest(eng, base, 1, 32, max_batches_per_conn=2)
    t = threading.Thread(target=_send_blocks, args=(base, [ones, ones[:5]]))
    t.start()
    t.join(timeout=30)
    with pytest.raises(fleetrec.FleetRecError):
        ing.wait()
    assert np.all(ing.scores[0, 0] == np.float32(KAT["small"]))
    ing.close()
    eng.close()
