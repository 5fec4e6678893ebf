"""fleetrec -- Python host side of the B200-native FleetRec inference hot path.

The product is ../libfleetrec.so (hand-written sm_100a kernels behind the C ABI
in /include/fleetrec.h); this package only binds it (ctypes) and carries the
model catalogues.  Nothing here computes on the CPU.
"""
from . import catalogue  # noqa: F401
from ._capi import (FR_ERR_CUDA, FR_ERR_INVALID, FR_ERR_OOM, FR_ERR_STATE, FR_ERR_UNSUPPORTED, FR_F16_GUARDED, FR_F16_OFF,  # noqa: F401
                    FR_HINT_AUTO, FR_HINT_LATENCY, FR_HINT_THROUGHPUT, FR_IDX_I32, FR_IDX_PACKED, FR_OPT_INDEX_FORMAT, FR_MLP_BIAS_RELU_SIGMOID, FR_MLP_LINEAR, FR_OK,
                    FR_OPT_CHECK_INDICES, FR_OPT_CUDA_GRAPHS, FR_OPT_F16_OPERANDS, FR_OPT_FUSE_LOOKUP, FR_OPT_TILE_HINT,
                    FR_PREC_FP32, FR_PREC_TF32, FR_TABLE_BF16, FR_TABLE_F16, FR_TABLE_F32, FR_TABLE_FP8)
from .engine import Batcher, Engine, FleetRecError, Ingest, Worker, merge_index, pack_indices  # noqa: F401
