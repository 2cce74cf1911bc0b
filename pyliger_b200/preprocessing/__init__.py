from ._normalization import normalize
from ._scale import scale_not_center

__all__ = ["normalize", "scale_not_center"]
