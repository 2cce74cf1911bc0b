"""B200-native drop-in for pyliger's integrative-NMF hot path.

Public names mirror pyliger (/root/reference/src/pyliger/__init__.py:4-5):
``optimize_ALS``, ``online_iNMF``, ``iNMF_HALS`` and the operator ``nnlsm_blockpivot``; plus the two O(nnz) stages that
produce their input, ``normalize`` and ``scale_not_center`` (preprocessing/_normalization.py, _scale.py), and the stage that consumes
H, ``quantile_norm`` (tools/_quantile_norm.py).
"""
from .containers import AnnDataLike, Liger, liger_from_matrices, liger_from_raw
from .factorization import iNMF_HALS, nnlsm_blockpivot, online_iNMF, optimize_ALS
from .preprocessing import normalize, scale_not_center
from .tools import quantile_norm

__all__ = ["optimize_ALS", "online_iNMF", "iNMF_HALS", "nnlsm_blockpivot", "normalize", "scale_not_center", "quantile_norm", "Liger",
           "AnnDataLike", "liger_from_matrices", "liger_from_raw"]
__version__ = "0.1.0"
