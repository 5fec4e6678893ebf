"""fleetrec -- Python host side of the B200-native FleetRec inference hot path.

The product is ../libfleetrec.so (hand-written sm_100a kernels behind the C ABI
in /include/fleetrec.h); this package only binds it (ctypes) and carries the
model catalogues.  Nothing here computes on the CPU.
"""
from . import catalogue  # noqa: F401
from ._capi import (FR_ERR_CUDA, FR_ERR_INVALID, FR_ERR_OOM, FR_ERR_STATE, FR_ERR_UNSUPPORTED, FR_MLP_BIAS_RELU_SIGMOID,  # noqa: F401
                    FR_MLP_LINEAR, FR_OK, FR_PREC_FP32, FR_PREC_TF32, FR_TABLE_BF16, FR_TABLE_F16, FR_TABLE_F32)
from .engine import Batcher, Engine, FleetRecError, Ingest, Worker, merge_index  # noqa: F401
