"""vectorindex_b200 -- B200-native (sm_100a) implementation of gifton/VectorIndex's batched search hot
path behind the reference's own operator API.  Everything computes in ``libvindex_b200.so`` (hand-written
CUDA, C ABI in include/); this package is the host-side mirror of the reference interface.  There is no
CPU fallback: importing the kernels without the built library, or calling them without a GPU, fails."""
from ._lib import (INDEX_FLAT, INDEX_IVF_FLAT, INDEX_IVF_PQ, LIB_PATH, METRIC_IP, METRIC_L2, ORDER_MAX, ORDER_MIN,
                   VectorIndexError)

__all__ = ["kernels", "index", "datagen", "VectorIndexError", "METRIC_L2", "METRIC_IP", "ORDER_MIN", "ORDER_MAX",
           "INDEX_FLAT", "INDEX_IVF_FLAT", "INDEX_IVF_PQ", "LIB_PATH"]
