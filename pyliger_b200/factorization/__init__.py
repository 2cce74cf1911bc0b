from ._iNMF_ANLS import optimize_ALS
from ._iNMF_HALS import iNMF_HALS
from ._online_iNMF import online_iNMF
from ._utilities import nnlsm_blockpivot

__all__ = ["optimize_ALS", "online_iNMF", "iNMF_HALS", "nnlsm_blockpivot"]
