"""The oracle against the reference's golden vectors and known answers (CPU).

Golden vectors (tests/golden/ref_*.npz) were produced by executing the
reference's own HLS lookup kernels (tests/golden/make_golden.py); the MLP
known answers are the reference README's (GPU/final_network_cublasLt_1_node_
no_FIFO_scatter/README.md:7-11).
"""
import hashlib
import os

import numpy as np
import pytest

from fleetrec import catalogue
from oracle import oracle, ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MODELS3 = ("small", "medium", "large_half")


def golden_tables(cat, seed):
    return [oracle.fill_reference(t.rows, t.dim) if t.tier == "PLRAM"
            else oracle.fill_hash(seed, t.id, t.rows, t.dim) for t in cat.tables]


@pytest.mark.parametrize("model", MODELS3)
def test_gather_matches_reference_wire_golden(model):
    g = np.load(os.path.join(GOLD, f"ref_{model}.npz"))
    cat = catalogue.load(model).with_row_cap(int(g["row_cap"]))
    tables = golden_tables(cat, int(g["seed"]))
    out = oracle.gather(cat, tables, oracle.idx_reference(96, cat.n_tables))
    assert np.array_equal(out[:32].view(np.uint32), g["wire_first32"].view(np.uint32))
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["wire96_sha256"])


@pytest.mark.parametrize("model", MODELS3)
def test_segments_match_reference_stream_order(model):
    """Catalogue segments vs the reference's gather_embeddings() fed tagged words:
    float f of stream s of item i carries i*65536 + s*256 + f."""
    g = np.load(os.path.join(GOLD, f"ref_{model}.npz"))
    cat = catalogue.load(model)
    n_pl = max(t.bank for t in cat.tables if t.tier == "PLRAM") + 1
    where = {}
    for tier, base, nb in (("HBM", 0, 28), ("DDR", 28, 2), ("PLRAM", 30, n_pl)):
        for b in range(nb):
            off = 0
            for t in [t for t in cat.tables if t.tier == tier and t.bank == b]:
                where[t.id] = (base + b, off)
                off += t.dim
    for item, key in ((0, "tagged_item0"), (31, "tagged_item31")):
        exp = np.zeros(cat.concat_floats, np.float32)
        for s in cat.segments:
            st, off = where[s.table]
            exp[s.dst:s.dst + s.len] = item * 65536 + st * 256 + off + s.col + np.arange(s.len)
        assert np.array_equal(exp, g[key])


@pytest.mark.parametrize("model", MODELS3)
@pytest.mark.skipif(not all(ref.available(m) for m in MODELS3), reason="oracle/_ref not built")
def test_live_reference_kernel_matches_oracle(model):
    """Run the compiled reference kernel now (oracle/_ref) with a different seed."""
    cat = catalogue.load(model).with_row_cap(128)
    tables = golden_tables(cat, 0xBEEF)
    out, _ = ref.run_top(model, cat, tables, batch_num=2)
    exp = oracle.gather(cat, tables, oracle.idx_reference(64, cat.n_tables))
    assert out.shape == exp.shape
    assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))


def test_reference_fill_and_lookup_kat():
    """SURVEY 8(c) lookup KAT: with the reference fill and index list, item j's whole
    concat vector is 1.0 when idx_random[j] is even, else 0.0 (incl. the medium pad)."""
    for model in MODELS3:
        cat = catalogue.load(model).with_row_cap(200)
        tables = oracle.make_tables(cat, "reference")
        out = oracle.gather(cat, tables, oracle.idx_reference(32, cat.n_tables))
        for j, r in enumerate(catalogue.IDX_RANDOM):
            assert np.all(out[j] == (1.0 if r % 2 == 0 else 0.0)), (model, j)
    ones = [j for j, r in enumerate(catalogue.IDX_RANDOM) if r % 2 == 0]
    assert ones == [2, 3, 7, 8, 9, 11, 15, 16, 17, 20, 21, 22, 23, 25, 27, 28, 29, 30, 31]


def test_fill_reference_debug_truncation():
    t = oracle.fill_reference(1000, 8, debug_rows=200)      # host.cpp:75 `#define DEBUG`
    assert np.all(t[0:200:2] == 1.0) and np.all(t[1:200:2] == 0.0) and np.all(t[200:] == 0.0)
    t = oracle.fill_reference(7, 4)                          # odd row count: last row untouched
    assert t[:, 0].tolist() == [1, 0, 1, 0, 1, 0, 0]


def test_hash_fill_is_finite_normal_and_distinct():
    t = oracle.fill_hash(7, 3, 4096, 16)
    bits = t.view(np.uint32)
    expo = (bits >> 23) & 0xFF
    assert expo.min() >= 118 and expo.max() <= 126
    assert np.isfinite(t).all() and np.abs(t).max() < 1.0 and np.abs(t).min() >= 2.0 ** -9
    assert len(np.unique(bits)) > 0.999 * bits.size
    assert oracle.hash_bits(7, 3, 5, 2) == int(bits[5, 2])


@pytest.mark.parametrize("inp,h1,expect", [(352, 1024, 47244640256.0), (512, 1024, 68719476736.0),
                                           (880, 1024, 118111600640.0), (1024, 1024, 137438953472.0),
                                           (3968, 2048, 1065151889408.0)])
def test_mlp_all_ones_known_answer(inp, h1, expect):
    """README.md:9,11: all-ones input and weights -> IN*H1*H2*H3 (exact in fp32)."""
    dims = [inp, h1, 512, 256, 1]
    W = [np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)]
    x = np.ones((5, inp), np.float32)
    out = oracle.mlp(x, dims, W, None, mode=0)
    assert np.all(out == np.float32(expect))
    assert np.all(oracle.mlp(x, dims, W, None, mode=0, acc64=True) == np.float32(expect))


def test_mlp_matches_numpy_float64():
    dims = [352, 1024, 512, 256, 1]
    W, b = oracle.make_weights(dims, seed=1)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (64, 352)).astype(np.float32)
    h = x.astype(np.float64)
    for k in range(4):
        h = h @ W[k].astype(np.float64) + b[k]
        if k < 3:
            h = np.maximum(h, 0)
    exp = 1 / (1 + np.exp(-h[:, 0]))
    got = oracle.mlp(x, dims, W, b, mode=1)
    assert np.max(np.abs(got - exp) / exp) < 1e-5
    h = x.astype(np.float64)
    for k in range(4):
        h = h @ W[k].astype(np.float64)
    got = oracle.mlp(x, dims, W, None, mode=0, acc64=True)
    assert np.allclose(got, h[:, 0], rtol=1e-5, atol=1e-6)


def test_gather_edge_cases():
    cat = catalogue.load("small").with_row_cap(64)
    tables = oracle.make_tables(cat, "hash")
    assert oracle.gather(cat, tables, np.zeros((0, 47), np.int32)).shape == (0, 352)   # empty batch
    idx = np.stack([np.array([t.rows - 1 for t in cat.tables], np.int32),                # last rows
                    np.zeros(47, np.int32)])                                              # first rows
    out = oracle.gather(cat, tables, idx)
    for s in cat.segments:
        assert np.array_equal(out[0, s.dst:s.dst + s.len], tables[s.table][-1, s.col:s.col + s.len])
        assert np.array_equal(out[1, s.dst:s.dst + s.len], tables[s.table][0, s.col:s.col + s.len])


def test_large_model_is_cpu_block_then_two_halves():
    big, half = catalogue.load("large"), catalogue.load("large_half")
    assert big.n_tables == 377 and big.concat_floats == 3968 and big.hidden[0] == 2048
    assert big.segments[0].len == 64 and big.tables[0].tier == "CPU"
    for s_h, s_a, s_b in zip(half.segments, big.segments[1:189], big.segments[189:]):
        assert (s_a.dst, s_a.table, s_a.len) == (64 + s_h.dst, 1 + s_h.table, s_h.len)
        assert (s_b.dst, s_b.table, s_b.len) == (64 + 1952 + s_h.dst, 189 + s_h.table, s_h.len)


def test_cartesian_merge_self_consistency():
    """gather(M, remap(iA,iB)) == gather(A,iA) || gather(B,iB)."""
    A, B = oracle.fill_hash(1, 0, 37, 4), oracle.fill_hash(1, 1, 11, 8)
    M = oracle.merge_tables(A, B)
    assert M.shape == (37 * 11, 12)
    for ia, ib in ((0, 0), (36, 10), (0, 10), (36, 0), (17, 5)):
        r = oracle.merge_index(ia, ib, 11)
        assert np.array_equal(M[r], np.concatenate([A[ia], B[ib]]))
    assert oracle.merge_index(99_999_999, 9_999_999, 10_000_000) == 99_999_999 * 10_000_000 + 9_999_999


def test_catalogue_bytes_and_flops():
    m = {n: catalogue.load(n) for n in catalogue.MODEL_NAMES}
    assert [m[n].gather_bytes_per_item() for n in ("small", "medium", "large")] == [1596, 3896, 17380]
    assert [m[n].mlp_flops_per_item() for n in ("small", "medium", "large")] == [2032128, 3113472, 18612736]
    assert abs(m["small"].table_bytes() / 1e9 - 1.415) < 1e-3


# ---------------------------------------------------------------- Cartesian-merge planner
def test_merge_planner_preserves_the_concat_vector():
    """plan -> merged catalogue -> remapped indices: gathering from the merged tables gives the
    ORIGINAL concat vector bit for bit (incl. corner rows), with fewer lookups per item, inside the
    byte budget and the int32 index range."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "gpu-fpga-recommendation-system_b200"))
    from fleetrec import catalogue, merge
    cat = catalogue.load("small").with_row_cap(300)
    budget = 40 << 20
    plan = merge.plan_merges(cat, budget)
    assert plan.pairs and plan.extra_bytes <= budget
    flat = [t for p in plan.pairs for t in p]
    assert len(flat) == len(set(flat))                                   # every table merged at most once
    mm = merge.apply_merges(cat, plan.pairs)
    assert merge.lookups_per_item(mm) == cat.n_tables - plan.lookups_saved
    assert mm.model.concat_floats == cat.concat_floats
    tables = oracle.make_tables(cat, "hash", seed=4)
    new_tables = [None] * mm.model.n_tables
    for new, old in enumerate(mm.kept):
        new_tables[new] = tables[old]
    for k, (a, b) in enumerate(plan.pairs):
        new_tables[mm.merged_ids[k]] = oracle.merge_tables(tables[a], tables[b])
        new_tables[mm.source_ids[k][0]], new_tables[mm.source_ids[k][1]] = tables[a], tables[b]
        assert new_tables[mm.merged_ids[k]].shape == (mm.model.tables[mm.merged_ids[k]].rows,
                                                      mm.model.tables[mm.merged_ids[k]].dim)
    idx = oracle.uniform_indices(cat, 64, seed=8)
    idx[0, :] = 0                                                        # corner rows: first / last of every table
    idx[1, :] = [t.rows - 1 for t in cat.tables]
    idx[2, ::2] = 0
    got = oracle.gather(mm.model, new_tables, mm.remap(idx))
    exp = oracle.gather(cat, tables, idx)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_merge_planner_respects_budget_and_int32():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "gpu-fpga-recommendation-system_b200"))
    from fleetrec import catalogue, merge
    cat = catalogue.load("medium")
    assert merge.plan_merges(cat, 0).pairs == []
    small = merge.plan_merges(cat, 1 << 20)
    big = merge.plan_merges(cat, 64 << 30)
    assert small.extra_bytes <= 1 << 20 and len(small.pairs) <= len(big.pairs)
    for a, b in big.pairs:
        assert cat.tables[a].rows * cat.tables[b].rows <= merge.INT32_MAX
    # an over-large product is refused outright (host.cpp:379-382 only warns)
    with pytest.raises(OverflowError):
        merge.apply_merges(cat, [(max(cat.tables, key=lambda t: t.rows).id,
                                  sorted(cat.tables, key=lambda t: t.rows)[-2].id)])
    mm = merge.apply_merges(cat, big.pairs[:1])
    bad = np.zeros((1, cat.n_tables), np.int64)
    a, b = big.pairs[0]
    bad[0, a], bad[0, b] = 2 ** 31, 5                                    # index outside the table: remap must not wrap
    with pytest.raises(OverflowError):
        mm.remap(bad)


def test_stated_dequant_is_rne_and_idempotent():
    """SURVEY.md 8(f)4 parity contract: f16 / bf16 table storage returns float32(round_to_nearest_even(x))."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1, 1, 4096).astype(np.float32), oracle.fill_hash(3, 7, 64, 16).ravel(),
                        np.float32([0.0, 1.0, -1.0, 0.5, 2.0 ** -9, 1 + 2.0 ** -8, 1 + 2.0 ** -11, 1 + 3 * 2.0 ** -11])])
    for dt, mant in ((1, 10), (2, 7)):
        q = oracle.quantize_dequantize(x, dt)
        assert np.array_equal(oracle.quantize_dequantize(q, dt), q)                       # idempotent
        assert np.all(np.abs(q - x) <= np.abs(x) * 2.0 ** -(mant + 1) + 1e-30)          # half an ulp
        assert np.array_equal(q.view(np.uint32) & ((1 << (23 - mant)) - 1), np.zeros(q.shape, np.uint32))
    assert oracle.quantize_dequantize(np.float32([1 + 2.0 ** -8]), 2)[0] == np.float32(1.0)            # tie -> even
    assert oracle.quantize_dequantize(np.float32([1 + 3 * 2.0 ** -8]), 2)[0] == np.float32(1 + 2.0 ** -6)
    assert np.array_equal(oracle.quantize_dequantize(x, 0), x)


def test_fp8_e4m3_dequant_restatement():
    """oracle.quantize_dequantize(x, 3): what a table stored as FP8 E4M3 returns -- round to nearest, ties to the even
    code, saturating at +-448, exact on representable values, idempotent, monotone, sign-symmetric."""
    q = lambda a: oracle.quantize_dequantize(np.asarray(a, np.float32), 3)   # noqa: E731
    codes = np.arange(127)
    grid = np.where(codes >> 3 == 0, (codes & 7) * 2.0 ** -9, (1 + (codes & 7) / 8.0) * 2.0 ** ((codes >> 3) - 7.0)).astype(np.float32)
    assert grid[-1] == 448.0 and grid[1] == 2.0 ** -9 and len(np.unique(grid)) == 127
    assert np.array_equal(q(grid), grid) and np.array_equal(q(-grid), -grid)              # exact on the grid
    x = np.random.default_rng(0).uniform(-500, 500, 20000).astype(np.float32)
    y = q(x)
    assert np.array_equal(q(y), y)                                                        # idempotent
    assert np.all(np.isin(np.abs(y), grid)) and np.all(np.abs(y) <= 448.0)
    xs = np.sort(x)
    assert np.all(np.diff(q(xs)) >= 0)                                                    # monotone
    assert np.all(np.abs(y - x)[np.abs(x) <= 448] <= np.abs(x[np.abs(x) <= 448]) * 2.0 ** -4 + 2.0 ** -10)   # half an ulp of 3 mantissa bits
    # ties go to the even code; saturation; underflow; signed zero; NaN
    assert list(q([1.0625, 1.1875, 17.0, 19.0, 449.0, 1e9, -1e9, 2.0 ** -10, 1.0001 * 2.0 ** -10])) == \
        [1.0, 1.25, 16.0, 20.0, 448.0, 448.0, -448.0, 0.0, 2.0 ** -9]
    assert np.signbit(q([-0.0]))[0] and np.isnan(q([np.nan]))[0]
