"""Shared plumbing of the preprocessing stages: host CSR <-> device arrays."""
import numpy as np
import scipy.sparse as sps
import torch

from .. import devcache


def to_device_csr(m, dt, device):
    """scipy matrix -> (rowptr int64, colidx int32, vals dt) on the device; reuses a registered device twin.
    Returns the arrays and the bytes copied host->device."""
    hit = devcache.lookup(m, device)
    if hit is not None:
        rowptr, colidx, vals = hit
        return rowptr, colidx, vals.to(dt), 0
    if not sps.isspmatrix_csr(m):
        m = sps.csr_matrix(m)
    if not m.has_sorted_indices:
        m = m.sorted_indices()
    ti = torch.from_numpy(np.ascontiguousarray(m.indptr))
    tc = torch.from_numpy(np.ascontiguousarray(m.indices))
    tv = torch.from_numpy(np.ascontiguousarray(m.data))
    nbytes = sum(t.numel() * t.element_size() for t in (ti, tc, tv))
    rowptr = ti.to(device).to(torch.int64)
    colidx = tc.to(device).to(torch.int32)
    vals = tv.to(device).to(dt)  # cast on the device (counts may be ints or float32)
    return rowptr, colidx, vals, nbytes


def to_host_csr(rowptr, colidx, vals, shape):
    """Device arrays -> scipy CSR float64 (what pyliger stores in the layers)."""
    data = vals.to(torch.float64).cpu().numpy()
    indices = colidx.cpu().numpy()
    indptr = rowptr.cpu().numpy()
    if indptr[-1] < 2 ** 31 - 1:
        indptr = indptr.astype(np.int32)
    else:
        indices = indices.astype(np.int64)
    out = sps.csr_matrix((data, indices, indptr), shape=shape)
    out.has_sorted_indices = True
    return out
