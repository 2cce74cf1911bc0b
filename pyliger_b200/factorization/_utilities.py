"""nnlsm_blockpivot on B200: same signature and return tuple as pyliger's operator
(/root/reference/src/pyliger/factorization/_utilities.py:181-301).

Host arrays in, host float64 arrays out; the normal-equation products and the batched block principal
pivoting run in libinmf_b200.so.  ``num_cholesky`` / ``num_eq`` count the per-column factorisations
the GPU solver performed (the reference counts LAPACK calls after grouping columns by passive set,
including a double count at _utilities.py:356-357; nothing in pyliger reads them).
"""
import numpy as np
import scipy.sparse as sps
import torch

from .. import _lib
from ..engine import _SUFFIX, _stream, require_cuda, resolve_dtype


def _solve_on_device(G, Rt, init_mask, k, dtype, device, want_Y=True):
    """G: double [nseg?]... single segment here. Rt: cols x k tensor (dtype). Returns X, Y, stats."""
    cols = Rt.shape[0]
    sfx = _SUFFIX[dtype]
    X = torch.empty_like(Rt)
    Y = torch.empty_like(Rt) if want_Y else None
    stats = torch.zeros(8, dtype=torch.int64, device=device)
    seg_off = torch.tensor([0, cols], dtype=torch.int64, device=device)
    warm = 0
    mask = torch.zeros(max(cols, 1), dtype=torch.int64, device=device)
    if init_mask is not None:
        mask.copy_(init_mask)
        warm = 1
    # G^-1 workspace: all-passive first round of a cold start; complement solves for passive sets > 31
    # and the deferral list of the half-warp solver
    ws = _lib.bpp_workspace(cols, 1, k, device)
    _lib.call("inmf_bpp", sfx, G, Rt, k, X, k, Y, k, mask, warm, seg_off, 1, cols, k, stats, *_lib.ws_args(ws), _stream())
    return X, Y, mask, stats


def pack_passive_mask(init):
    """(n, cols) array -> int64 bit masks of ``init > 0`` per column (_utilities.py:223)."""
    P = np.asarray(init) > 0
    n = P.shape[0]
    w = (np.uint64(1) << np.arange(n, dtype=np.uint64))[:, None]
    return (P.astype(np.uint64) * w).sum(axis=0, dtype=np.uint64).view(np.int64)


def nnlsm_blockpivot(A, B, is_input_prod=False, init=None, *, dtype="float64", device=None):
    """Solves min ||AX-B||_2^2 s.t. X >= 0 column by column with block principal pivoting.

    Returns ``X, (success, Y, num_cholesky, num_eq, num_backup)`` with X, Y of shape (n, cols).
    """
    device = require_cuda(device)
    dt = resolve_dtype(dtype)
    ndt = np.float32 if dt == torch.float32 else np.float64
    sfx = _SUFFIX[dt]
    with torch.cuda.device(device):
        if is_input_prod:
            AtA = np.asarray(A, dtype=np.float64)
            AtB = B.toarray() if sps.issparse(B) else np.asarray(B)
            n, cols = AtB.shape
            G = torch.from_numpy(np.ascontiguousarray(AtA)).to(device)
            Rt = torch.from_numpy(np.ascontiguousarray(AtB.T.astype(ndt, copy=False))).to(device)
        else:
            A = np.ascontiguousarray(np.asarray(A, dtype=ndt))
            m, n = A.shape
            cols = B.shape[1]
            Ad = torch.from_numpy(A).to(device)
            G = torch.zeros(n, n, dtype=torch.float64, device=device)
            Rt = torch.empty(cols, n, dtype=dt, device=device)
            if sps.issparse(B):
                # (A'B)' = B'A: stream B' as CSR against the dense A
                Bt = sps.csc_matrix(B)
                Bt.sort_indices()
                rowptr = torch.from_numpy(Bt.indptr.astype(np.int64)).to(device)
                colidx = torch.from_numpy(Bt.indices.astype(np.int32)).to(device)
                vals = torch.from_numpy(Bt.data.astype(ndt)).to(device)
                seg = torch.tensor([[0, 0, 0, max(cols, 1)], [cols, 0, 0, 0]], dtype=torch.int64, device=device)
                _lib.call("inmf_normal_eq", sfx, Ad, None, m, n, 0, G, None, _stream())
                _lib.call("inmf_spmm", sfx, rowptr, colidx, vals, Ad, n, m, Rt, n, seg, 1, cols, n, _stream())
            else:
                Bd = torch.from_numpy(np.ascontiguousarray(np.asarray(B, dtype=ndt))).to(device)
                _lib.call("inmf_normal_eq", sfx, Ad, Bd, m, n, cols, G, Rt, _stream())
        if n > 64:
            raise ValueError("pyliger_b200.nnlsm_blockpivot supports at most 64 unknowns per column, got %d" % n)
        init_mask = None
        if init is not None:
            init_mask = torch.from_numpy(pack_passive_mask(init)).to(device)
        if cols == 0:
            return np.zeros((n, 0)), (True, np.zeros((n, 0)), 0, 0, 0)
        X, Y, _, stats = _solve_on_device(G, Rt, init_mask, n, dt, device)
        st = stats.cpu().numpy()
        Xh = X.to(torch.float64).cpu().numpy().T
        Yh = Y.to(torch.float64).cpu().numpy().T
    success = bool(st[0] == 0)
    return np.ascontiguousarray(Xh), (success, np.ascontiguousarray(Yh), int(st[1]), int(st[1]), int(st[2]))
