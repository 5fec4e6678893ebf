"""Where do the pipelines of tc_linear_kernel spend their time?  (experiments build, run on the GPU box)

  FLEETREC_LIB=gpu-fpga-recommendation-system_b200/libfleetrec_exp.so FR_TC_PROF=1 python tools/tc_prof.py small 2048 16384

Per layer and batch size: one launch alone; the first CTA pair's cycle counters (clock64): the TMA producer's total
and the part spent waiting for a free ring slot, the MMA issuer's total and the parts spent waiting for operands
(full barrier) and for a free accumulator stage, one epilogue thread's total and the parts spent waiting for an
accumulator and for a staging buffer (+ the named barrier)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
import fleetrec  # noqa: E402
from fleetrec import _capi, catalogue  # noqa: E402
from oracle import oracle  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "small"
batches = [int(a) for a in sys.argv[2:]] or [2048, 16384]
cat = catalogue.load(model).with_row_cap(200000)
dims = cat.layer_dims
W, b = oracle.make_weights(dims)
eng = fleetrec.Engine(cat, max_batch=max(batches))
eng.set_option(fleetrec.FR_OPT_TILE_HINT, fleetrec.FR_HINT_THROUGHPUT)
eng.fill_hash()
eng.load_mlp(W, b)
raw = C.CDLL(_capi.LIB_PATH)
raw.frdbg_enqueue_layer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
raw.frdbg_tc_prof.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
raw.frdbg_layer_ctas.argtypes = [C.c_void_p, C.c_int]
MHZ = 1965.0
print(f"# {model}, tiles {os.environ.get('FR_TC_TILES', 'auto')}; times in us at {MHZ:.0f} MHz; CTA 0 = pair leader, CTA 1 = its peer")
for B in batches:
    idx = oracle.zipf_indices(cat, B)
    eng.infer(idx)
    ms = eng.time_kernels(idx, B, reps=20)
    for k in range(3):
        for _ in range(4):
            assert raw.frdbg_enqueue_layer(eng._h, k, B, None) == 0
        eng.sync()
        buf = (C.c_longlong * 128)()
        assert raw.frdbg_tc_prof(eng._h, buf, 128) == 128
        ctas = raw.frdbg_layer_ctas(eng._h, k)
        print(f"B={B} layer {k + 1} (K={dims[k]}, N={dims[k + 1]}): {ms[1 + k] * 1e3:.1f} us alone, {ctas} CTAs")
        for c in range(2):
            v = [buf[c * 16 + i] / MHZ for i in range(12)]
            print(f"   CTA {c}: producer {v[0]:6.1f} (waiting for a slot {v[1]:6.1f}, {buf[c * 16 + 2]} slices) | "
                  f"MMA {v[4]:6.1f} (operands {v[5]:6.1f}, accumulator {v[6]:5.1f}, {buf[c * 16 + 7]} slices) | "
                  f"epilogue {v[8]:6.1f} (accumulator {v[9]:6.1f}, staging {v[10]:5.1f}, {buf[c * 16 + 11]} tiles)")
        n = buf[127]
        tl = sorted(tuple(buf[112 + 3 * j + i] for i in range(3)) for j in range(4))
        print("   launch timeline of CTA 0 (globaltimer, us): " + "; ".join(
            f"prologue {(b_ - a) / 1e3:.2f} body {(c_ - b_) / 1e3:.2f}" + (f" gap-to-next-entry {(tl[j + 1][0] - c_) / 1e3:.2f}" if j < 3 else "")
            for j, (a, b_, c_) in enumerate(tl)) + f"   ({n} launches)")
eng.close()
