"""numpy restatement of the range analysis behind FR_OPT_F16_OPERANDS = FR_F16_GUARDED (csrc/fr_precision.cu).
TEST HELPER: the library decides on the device; the tests cross-check its bounds against this.

The tcgen05 kernels can run on fp16 operands and activations (csrc/fr_mlp_tc.cu, ELT = 2): the 11-bit
significand the TF32 path keeps, in half the bytes -- and with fp16's range.  An operand above 65 504 is
infinite there, so the path may only be chosen when nothing the kernels store can get that large.  What
they store: the concat vector x (table values), the weights, and the activations h1, h2 (h3 never leaves
TMEM, the output layer runs in fp32).  The bound is propagated unit by unit:

    |x_i|   <= max |value| of the table position i comes from
    |h_o|   <= sum_i |W[i][o]| * ub_in[i] + |b[o]|         (ReLU only lowers it; LINEAR mode: the same bound on |h|)

It is a worst-case bound (every input at its largest magnitude with the sign that hurts), so "safe" is a
guarantee and "unsafe" only means "not provable" -- the reference's all-ones known answer is genuinely
unsafe (352 * 1024 at layer 2).
"""
import numpy as np

F16_MAX = 65504.0


def concat_bounds(model, table_max_abs):
    """Upper bound of |x| per concat position: the largest magnitude of the table it is copied from
    (`table_max_abs[t]`; the medium model's duplicate pad copies a table position like any other)."""
    ub = np.zeros(model.concat_floats, np.float64)
    for s in model.segments:
        ub[s.dst:s.dst + s.len] = float(table_max_abs[s.table])
    return ub


def layer_bounds(ub_in, W, b=None):
    """Per-unit bound of |W^T x + b| given per-input bounds; W is [in][out] as the reference keeps it."""
    out = np.abs(np.asarray(W, np.float64)).T @ np.asarray(ub_in, np.float64)
    if b is not None:
        out = out + np.abs(np.asarray(b, np.float64))
    return out


def f16_safe(model, table_max_abs, W, b=None, margin=2.0):
    """(safe, report): may the engine compute on fp16 operands with these tables and weights?

    safe   -- every stored operand (x, W1..W3, h1, h2) is provably below F16_MAX / margin
    report -- dict of the largest bound per stored tensor, for the log
    """
    rep = {}
    ub = concat_bounds(model, table_max_abs)
    rep["x"] = float(ub.max())
    for k in range(3):
        rep[f"W{k + 1}"] = float(np.max(np.abs(W[k])))
    for k in range(2):   # h1, h2 are stored; h3 stays in TMEM and the output layer is fp32
        ub = layer_bounds(ub, W[k], None if b is None else b[k])
        rep[f"h{k + 1}"] = float(ub.max())
    safe = all(np.isfinite(v) and v <= F16_MAX / margin for v in rep.values())
    return safe, rep
