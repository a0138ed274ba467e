"""rcppsparse_b200 — B200-native backend for the column-sweep hot path of zdebruine/RcppSparse.

Only what the path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI of libsparse_b200),
``matrix`` (host-side mirror of ``RcppSparse::Matrix`` and ``columnSums``), ``shard`` (column
sharding across ranks), ``synth`` (deterministic dgCMatrix generators for tests and benchmarks).
"""
from ._lib import SparseB200Error
from .matrix import DeviceMatrix, Matrix, ShardedHostMatrix, columnSums

__all__ = ["Matrix", "DeviceMatrix", "ShardedHostMatrix", "columnSums", "SparseB200Error"]
