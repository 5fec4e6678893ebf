import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "experimental: exercises a kernel variant that exists only in libfleetrec_exp.so "
                                       "(FLEETREC_LIB=...; deselected otherwise)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _lib_has_experiments():
    try:
        from fleetrec import _capi
        return bool(_capi.lib().fr_build_has_experiments())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if not _lib_has_experiments():     # tests of the experiments build: not part of this library's suite
        drop = [it for it in items if "experimental" in it.keywords]
        if drop:
            config.hook.pytest_deselected(items=drop)
            items[:] = [it for it in items if "experimental" not in it.keywords]
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
