"""B200-native drop-in for pyliger's integrative-NMF hot path.

Public names mirror pyliger (/root/reference/src/pyliger/__init__.py:4-5):
``optimize_ALS``, ``online_iNMF`` and the operator ``nnlsm_blockpivot``.
"""
from .containers import AnnDataLike, Liger, liger_from_matrices
from .factorization import nnlsm_blockpivot, online_iNMF, optimize_ALS

__all__ = ["optimize_ALS", "online_iNMF", "nnlsm_blockpivot", "Liger", "AnnDataLike",
           "liger_from_matrices"]
__version__ = "0.1.0"
