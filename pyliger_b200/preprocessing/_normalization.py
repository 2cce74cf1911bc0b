"""normalize on B200: same call and same slots as pyliger's
(/root/reference/src/pyliger/preprocessing/_normalization.py:15-56, in-memory branch :91-99).

Writes ``layers["norm_data"]`` (scipy CSR float64) and ``var["norm_sum"]``, ``var["norm_sum_sq"]``,
``var["norm_mean"]`` of every dataset.  The row normalisation and the per-gene sums run in one CUDA pass
(``inmf_normalize_rows_*``); the device copy of norm_data is kept for scale_not_center (devcache).
"""
import numpy as np
import torch

from .. import _lib, devcache
from ..engine import _SUFFIX, _stream, require_cuda, resolve_dtype
from ._device import to_device_csr, to_host_csr


def normalize(liger_object, remove_missing=True, chunk_size=1000, *, dtype="float64", device=None):
    device = require_cuda(device)
    dt = resolve_dtype(dtype)
    sfx = _SUFFIX[dt]
    with torch.cuda.device(device):
        for idx, adata in enumerate(liger_object.adata_list):
            if adata.isbacked:  # on-disk mode: chunk by chunk against the store (_normalization.py:44-46, 61-88)
                from ._backed import normalize_online
                norm_sum, norm_sum_sq = normalize_online(adata, chunk_size, dtype, device)
                adata.var["norm_sum"] = norm_sum
                adata.var["norm_sum_sq"] = norm_sum_sq
                adata.var["norm_mean"] = norm_sum / adata.shape[0]
                continue
            n, g = adata.shape
            rowptr, colidx, vals, _ = to_device_csr(adata.X, dt, device)
            norm_vals = torch.empty_like(vals)
            sums = torch.zeros(2, g, dtype=torch.float64, device=device)
            n_missing = torch.zeros(1, dtype=torch.int64, device=device)
            _lib.call("inmf_normalize_rows", sfx, rowptr, colidx, vals, n, norm_vals, sums[0], sums[1], n_missing,
                      _stream())
            missing = int(n_missing.item())
            if remove_missing and missing > 0:
                # The reference drops those cells from a LOCAL copy (_normalization.py:93-94) and then fails to
                # store the shorter norm_data into the original AnnData (:51).  Same outcome, said clearly.
                print("Removing {} cells not expressing any genes in {}.".format(missing, adata.uns["sample_name"]))
                raise ValueError(
                    "normalize: %d cells of %s express no genes; remove them before normalize (the reference "
                    "fails here with an AnnData layer-shape error)" % (missing, adata.uns["sample_name"]))
            norm_data = to_host_csr(rowptr, colidx, norm_vals, (n, g))
            devcache.register(norm_data, device, rowptr, colidx, norm_vals)
            sums_h = sums.cpu().numpy()
            target = liger_object.adata_list[idx]
            target.layers["norm_data"] = norm_data
            target.var["norm_sum"] = sums_h[0]
            target.var["norm_sum_sq"] = sums_h[1]
            target.var["norm_mean"] = sums_h[0] / n
    return None
