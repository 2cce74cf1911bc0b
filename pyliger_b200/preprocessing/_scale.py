"""scale_not_center on B200: same call and same slots as pyliger's
(/root/reference/src/pyliger/preprocessing/_scale.py:12-57, in-memory branch :92-102).

Every dataset is cut down to the variable genes (``adata.raw`` keeps the full one) and gets
``layers["scale_data"]`` = norm_data[:, var genes] * 1/sqrt(norm_sum_sq / (n - 1)) as scipy CSR float64.
Gene selection + scaling is one compaction on the device (``inmf_select_count`` / ``inmf_select_scale_*``);
the device copy of scale_data is kept for optimize_ALS / online_iNMF (devcache), so the factorisation does not
upload it again.
"""
import numpy as np
import torch

from .. import _lib, devcache
from ..engine import _SUFFIX, _stream, require_cuda, resolve_dtype
from ._device import to_device_csr, to_host_csr


def scale_not_center(liger_object, remove_missing=True, chunk_size=1000, *, dtype="float64", device=None):
    device = require_cuda(device)
    dt = resolve_dtype(dtype)
    sfx = _SUFFIX[dt]
    with torch.cuda.device(device):
        for idx, adata in enumerate(liger_object.adata_list):
            var_gene_idx = np.asarray(adata.var.index.isin(liger_object.var_genes))
            if adata.isbacked:  # on-disk mode (_scale.py:49-51, 62-89)
                from ._backed import scale_online
                liger_object.adata_list[idx] = scale_online(adata, var_gene_idx, chunk_size, dtype, device)
                continue
            n, g = adata.shape
            norm = adata.layers["norm_data"]
            rowptr, colidx, vals, _ = to_device_csr(norm, dt, device)

            # slice adata after keeping a raw version (_scale.py:94-96)
            adata.raw = adata
            sub = adata[:, var_gene_idx].copy()
            g_new = int(var_gene_idx.sum())
            with np.errstate(divide="ignore"):
                scaler = 1 / np.sqrt(sub.var["norm_sum_sq"].to_numpy() / (n - 1))  # :100
            colmap_h = np.full(g, -1, dtype=np.int32)
            colmap_h[var_gene_idx] = np.arange(g_new, dtype=np.int32)
            colmap = torch.from_numpy(colmap_h).to(device)
            scaler_d = torch.from_numpy(np.ascontiguousarray(scaler, dtype=np.float64)).to(device)

            counts = torch.zeros(n, dtype=torch.int64, device=device)
            _lib.call("inmf_select_count", "", rowptr, colidx, n, colmap, counts, _stream())
            new_rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
            torch.cumsum(counts, 0, out=new_rowptr[1:])
            nnz_new = int(new_rowptr[-1].item())
            new_colidx = torch.empty(nnz_new, dtype=torch.int32, device=device)
            new_vals = torch.empty(nnz_new, dtype=dt, device=device)
            _lib.call("inmf_select_scale", sfx, rowptr, colidx, vals, n, colmap, scaler_d, new_rowptr, new_colidx,
                      new_vals, _stream())
            scale_data = to_host_csr(new_rowptr, new_colidx, new_vals, (n, g_new))
            devcache.register(scale_data, device, new_rowptr, new_colidx, new_vals)
            sub.layers["scale_data"] = scale_data
            liger_object.adata_list[idx] = sub
    return None
