"""The numpy restatement of the fp16 range analysis (tests/f16_bound_ref.py) behaves as specified -- no GPU.
The library's own analysis (csrc/fr_precision.cu) is checked against it in tests/test_round2_gpu.py."""
import numpy as np

import f16_bound_ref as precision
from fleetrec import catalogue
from oracle import oracle


def test_bound_dominates_every_stored_activation():
    """The propagated bound is a true upper bound of what the kernels would store (x, h1, h2), on random data."""
    cat = catalogue.load("small").with_row_cap(2000)
    dims = cat.layer_dims
    tables = oracle.make_tables(cat, "hash", seed=3)
    W, b = oracle.make_weights(dims, seed=42)
    tmax = [float(np.max(np.abs(t))) for t in tables]
    safe, rep = precision.f16_safe(cat, tmax, W, b)
    x = oracle.gather(cat, tables, oracle.zipf_indices(cat, 512, seed=1))
    h1 = np.maximum(x @ W[0] + b[0], 0)
    h2 = np.maximum(h1 @ W[1] + b[1], 0)
    assert np.abs(x).max() <= rep["x"] and h1.max() <= rep["h1"] and h2.max() <= rep["h2"]
    assert safe and rep["h2"] < precision.F16_MAX / 2          # hash-filled tables (|x| <= 1), N(0, 1/in) weights


def test_reference_known_answer_is_not_fp16_safe():
    """README.md:7-11's all-ones KAT passes 352 * 1024 at layer 2: the analysis must refuse fp16 operands."""
    cat = catalogue.load("small")
    dims = cat.layer_dims
    W = [np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)]
    safe, rep = precision.f16_safe(cat, [1.0] * cat.n_tables, W, None)
    assert not safe
    assert rep["h1"] == 352.0 and rep["h2"] == 352.0 * 1024.0


def test_medium_duplicate_pad_and_large_tables_are_covered():
    """Every concat position gets its table's bound (including the medium model's duplicated 4 floats); a single
    huge table value makes the whole engine unsafe."""
    cat = catalogue.load("medium")
    tmax = [1.0] * cat.n_tables
    ub = precision.concat_bounds(cat, tmax)
    assert ub.shape == (cat.concat_floats,) and np.all(ub == 1.0)
    tmax[5] = 1e6
    W, b = oracle.make_weights(cat.layer_dims, seed=1)
    safe, rep = precision.f16_safe(cat, tmax, W, b)
    assert not safe and rep["x"] == 1e6
