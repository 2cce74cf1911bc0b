"""Backed ("online learning") preprocessing: the chunked stages pyliger runs against its on-disk store
(/root/reference/src/pyliger/_utilities.py:137-154 ``_create_h5_using_adata``; preprocessing/_initialization.py:77-122
``_initialization_online``; _normalization.py:61-88 ``_normalize_online``; _scale.py:62-89 ``_scale_online``;
factorization/_online_iNMF.py:604-620 ``_preprocessing``), on ``pyliger_b200.io.CSRStore`` instead of h5sparse.

Row chunks of ``chunk_size`` cells go store -> device -> store: the arithmetic of every chunk is the same CUDA pass as
the in-memory stages (``inmf_normalize_rows_*`` with the per-gene sums accumulated on the device across chunks,
``inmf_select_count`` / ``inmf_select_scale_*``), so backed and in-memory results are identical.
"""
import os

import numpy as np
import scipy.sparse as sps
import torch

from .. import _lib
from ..engine import _SUFFIX, _stream, require_cuda, resolve_dtype
from ..io import CSRStore, h5_idx_generator, store_path
from ._device import to_device_csr, to_host_csr


def _path(adata):
    return adata.uns.get("backed_path") or store_path(adata)


def create_store_using_adata(adata, chunk_size):
    """raw_data of an in-memory AnnData written chunk by chunk (_utilities.py:137-154)."""
    p = _path(adata)
    os.makedirs(os.path.dirname(p) or ".", exist_ok=True)
    with CSRStore(p, "w") as f:
        for left, right in h5_idx_generator(chunk_size, adata.shape[0]):
            chunk = sps.csr_matrix(adata.X[left:right])
            if "raw_data" not in f.keys():
                f.create_dataset("raw_data", data=chunk, chunks=(chunk_size,), maxshape=(None,))
            else:
                f["raw_data"].append(chunk)
    adata.uns["backed_path"] = p
    return p


def initialization_online(adata, chunk_size, remove_missing):
    """Gene sums / sums of squares and per-cell nUMI / nGene of the raw data, chunk by chunk; genes never expressed
    are dropped from the annotations when ``remove_missing`` (_initialization.py:77-122)."""
    g = adata.shape[1]
    gene_sum, gene_sum_sq = np.zeros(g), np.zeros(g)
    nUMI, nGene = np.zeros(adata.shape[0]), np.zeros(adata.shape[0])
    p = _path(adata)
    if not os.path.isdir(p):
        create_store_using_adata(adata, chunk_size)
    with CSRStore(p, "r") as f:
        for left, right in h5_idx_generator(chunk_size, adata.shape[0]):
            raw = sps.csr_matrix(f["raw_data"][left:right])
            gene_sum += np.ravel(raw.sum(axis=0))
            gene_sum_sq += np.ravel(raw.power(2).sum(axis=0))
            nUMI[left:right] = np.ravel(raw.sum(axis=1))
            nGene[left:right] = raw.getnnz(axis=1)
    idx_missing = gene_sum == 0 if remove_missing else np.repeat(False, g)
    if np.sum(idx_missing) > 0:
        print("Removing {} genes not expressing in {}.".format(np.sum(idx_missing), adata.uns["sample_name"]))
    out = adata[:, ~idx_missing].copy()
    out.isbacked = True
    out.uns["backed_path"] = p
    out.var["gene_sum"] = gene_sum[~idx_missing]
    out.var["gene_sum_sq"] = gene_sum_sq[~idx_missing]
    out.obs["nUMI"] = nUMI
    out.obs["nGene"] = nGene
    out.uns["idx_missing"] = idx_missing
    return out


def normalize_online(adata, chunk_size, dtype="float64", device=None):
    """norm_data appended to the store chunk by chunk; returns (norm_sum, norm_sum_sq) (_normalization.py:61-88)."""
    device = require_cuda(device)
    dt = resolve_dtype(dtype)
    keep = ~np.asarray(adata.uns.get("idx_missing", np.zeros(0, dtype=bool)), dtype=bool)
    with torch.cuda.device(device), CSRStore(_path(adata), "r+") as f:
        g_raw = f["raw_data"].shape[1]
        if keep.size == 0:
            keep = np.ones(g_raw, dtype=bool)
        g = int(keep.sum())
        sums = torch.zeros(2, g, dtype=torch.float64, device=device)
        n_missing = torch.zeros(1, dtype=torch.int64, device=device)
        for left, right in h5_idx_generator(chunk_size, adata.shape[0]):
            raw = f["raw_data"][left:right]
            if g != g_raw:
                raw = raw[:, keep]
            rowptr, colidx, vals, _ = to_device_csr(raw, dt, device)
            norm_vals = torch.empty_like(vals)
            _lib.call("inmf_normalize_rows", _SUFFIX[dt], rowptr, colidx, vals, right - left, norm_vals, sums[0], sums[1],
                      n_missing, _stream())  # the sums accumulate across the chunks
            norm = to_host_csr(rowptr, colidx, norm_vals, (right - left, g))
            if "norm_data" not in f.keys():
                f.create_dataset("norm_data", data=norm, chunks=(chunk_size,), maxshape=(None,))
            else:
                f["norm_data"].append(norm)
        s = sums.cpu().numpy()
    return s[0], s[1]


def scale_online(adata, var_gene_idx, chunk_size, dtype="float64", device=None):
    """scale_data (variable genes only) appended to the store chunk by chunk; returns the AnnData cut down to the
    variable genes (_scale.py:62-89)."""
    device = require_cuda(device)
    dt = resolve_dtype(dtype)
    var_gene_idx = np.asarray(var_gene_idx, dtype=bool)
    n, g = adata.shape
    g_new = int(var_gene_idx.sum())
    with np.errstate(divide="ignore"):
        scaler = 1 / np.sqrt(adata.var["norm_sum_sq"][var_gene_idx].to_numpy() / (n - 1))
    colmap_h = np.full(g, -1, dtype=np.int32)
    colmap_h[var_gene_idx] = np.arange(g_new, dtype=np.int32)
    with torch.cuda.device(device), CSRStore(_path(adata), "r+") as f:
        colmap = torch.from_numpy(colmap_h).to(device)
        scaler_d = torch.from_numpy(np.ascontiguousarray(scaler, dtype=np.float64)).to(device)
        for left, right in h5_idx_generator(chunk_size, n):
            m = right - left
            rowptr, colidx, vals, _ = to_device_csr(f["norm_data"][left:right], dt, device)
            counts = torch.zeros(m, dtype=torch.int64, device=device)
            _lib.call("inmf_select_count", "", rowptr, colidx, m, colmap, counts, _stream())
            new_rowptr = torch.zeros(m + 1, dtype=torch.int64, device=device)
            torch.cumsum(counts, 0, out=new_rowptr[1:])
            nnz_new = int(new_rowptr[-1].item())
            new_colidx = torch.empty(nnz_new, dtype=torch.int32, device=device)
            new_vals = torch.empty(nnz_new, dtype=dt, device=device)
            _lib.call("inmf_select_scale", _SUFFIX[dt], rowptr, colidx, vals, m, colmap, scaler_d, new_rowptr, new_colidx,
                      new_vals, _stream())
            chunk = to_host_csr(new_rowptr, new_colidx, new_vals, (m, g_new))
            if "scale_data" not in f.keys():
                f.create_dataset("scale_data", data=chunk, chunks=(chunk_size,), maxshape=(None,))
            else:
                f["scale_data"].append(chunk)
    out = adata[:, var_gene_idx].copy()
    out.isbacked = True
    out.uns["backed_path"] = _path(adata)
    return out


def preprocessing(adata, h5_chunk_size, var_genes, dtype="float64", device=None):
    """online_iNMF's preprocessing of a raw AnnData handed in as X_new (_online_iNMF.py:604-620): store, gene sums,
    normalisation, scaling on the variable genes."""
    adata = initialization_online(adata, h5_chunk_size, remove_missing=False)
    norm_sum, norm_sum_sq = normalize_online(adata, h5_chunk_size, dtype, device)
    norm_sum[norm_sum < 1e-16] = 1e-16        # nonneg(): avoid zero division
    norm_sum_sq[norm_sum_sq < 1e-16] = 1e-16
    adata.var["norm_sum"] = norm_sum
    adata.var["norm_sum_sq"] = norm_sum_sq
    adata.var["norm_mean"] = norm_sum / adata.shape[0]
    var_gene_idx = adata.var.index.isin(var_genes)
    return scale_online(adata, var_gene_idx, h5_chunk_size, dtype, device)
