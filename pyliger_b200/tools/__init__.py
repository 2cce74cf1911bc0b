from ._quantile_norm import quantile_norm

__all__ = ["quantile_norm"]
