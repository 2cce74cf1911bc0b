"""Synthetic scaled count matrices shaped like pyliger's ``scale_data`` layer (SURVEY.md §8 d).

Two generators, both host numpy with ``default_rng(1000 + i)`` per dataset ``i``:

* ``structured=True``  -- latent ``H* ~ Gamma(0.5, 1)``, ``W*, V_i* ~ Gamma(0.5, 1)`` (``V*`` scaled
  0.3), counts ``~ Poisson(s * H*(W* + V_i*))`` with the depth ``s`` found by bisection so that the
  expected density is ``density``;
* ``structured=False`` -- i.i.d. sparsity pattern with ``Gamma(2, 1)`` values (the survey probes).

Both then apply exactly the reference preprocessing maths: L1-normalise every cell
(/root/reference/src/pyliger/preprocessing/_normalization.py:95) and scale every gene by
``1/sqrt(sum_c norm^2 / (n_i - 1))`` without centring (``_scale.py:100-101``); genes that are all
zero in any dataset and cells that are all zero are dropped.  Output: list of scipy CSR float64
matrices (cells x genes) with sorted column indices, identical gene columns.
"""
import numpy as np
import scipy.sparse as sps


def _depth_for_density(mu_sample, density):
    lo, hi = 0.0, 1.0
    while np.mean(1.0 - np.exp(-hi * mu_sample)) < density:
        hi *= 2.0
        if hi > 1e12:
            break
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if np.mean(1.0 - np.exp(-mid * mu_sample)) < density:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def _structured_counts(n, g, k_true, density, rng, W_star, chunk=8192):
    V_star = 0.3 * rng.gamma(0.5, 1.0, size=(k_true, g))
    M = W_star + V_star
    Hs = rng.gamma(0.5, 1.0, size=(min(n, 2048), k_true))
    s = _depth_for_density(Hs @ M, density)
    blocks = []
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        H = rng.gamma(0.5, 1.0, size=(m, k_true))
        C = rng.poisson(s * (H @ M)).astype(np.float64)
        blocks.append(sps.csr_matrix(C))
    return sps.vstack(blocks).tocsr()


def _iid_counts_fast(n, g, density, rng):
    """Pattern by Bernoulli thresholding in row chunks (fast, no per-row Python loop)."""
    blocks = []
    chunk = max(1, int(4e7 // g))
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        mask = rng.random((m, g), dtype=np.float32) < density
        r, c = np.nonzero(mask)
        vals = rng.gamma(2.0, 1.0, size=r.size)
        blocks.append(sps.csr_matrix((vals, (r, c)), shape=(m, g)))
    return sps.vstack(blocks).tocsr()


def scale_like_pyliger(counts_list):
    """Row-L1 normalise, drop empty cells / genes, scale genes to RMS 1 with the (n-1) denominator."""
    mats = []
    for C in counts_list:
        C = sps.csr_matrix(C, dtype=np.float64)
        rs = np.asarray(C.sum(axis=1)).ravel()
        C = C[rs > 0]
        rs = rs[rs > 0]
        mats.append(sps.diags(1.0 / rs) @ C)
    keep = np.ones(mats[0].shape[1], dtype=bool)
    for Nm in mats:
        keep &= np.asarray(Nm.multiply(Nm).sum(axis=0)).ravel() > 0
    out = []
    for Nm in mats:
        Nm = Nm.tocsc()[:, keep].tocsr()
        ss = np.asarray(Nm.multiply(Nm).sum(axis=0)).ravel()
        scaler = 1.0 / np.sqrt(ss / (Nm.shape[0] - 1))
        X = (Nm @ sps.diags(scaler)).tocsr()
        X.sort_indices()
        out.append(X)
    return out


def make_datasets(ns, g, k_true, density=0.08, structured=True, seed0=1000):
    """List of scaled CSR float64 matrices, one per entry of ``ns`` (cells per dataset)."""
    counts = []
    W_star = np.random.default_rng(seed0 - 1).gamma(0.5, 1.0, size=(k_true, g))
    for i, n in enumerate(ns):
        rng = np.random.default_rng(seed0 + i)
        if structured:
            counts.append(_structured_counts(int(n), g, k_true, density, rng, W_star))
        else:
            counts.append(_iid_counts_fast(int(n), g, density, rng))
    return scale_like_pyliger(counts)


def make_datasets_device(ns, g, k_true, density=0.08, seed0=1000, device="cuda", chunk=16384):
    """Same structured generator as ``make_datasets`` but the Gamma/Poisson draws run on the GPU with
    torch's RNG (seconds instead of a minute at 4 x 50k x 3k); scaling is the same host code.
    Different random stream than the numpy generator, same distribution."""
    import torch

    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed0) - 1)

    def gamma_half(shape):
        # Gamma(1/2, 1) = N(0,1)^2 / 2
        return torch.randn(shape, device=dev, dtype=torch.float32, generator=gen).square_().mul_(0.5)

    W_star = gamma_half((k_true, g))
    counts = []
    for i, n in enumerate(ns):
        gen.manual_seed(int(seed0) + i)
        M = W_star + 0.3 * gamma_half((k_true, g))
        Hs = gamma_half((2048, k_true))
        s = _depth_for_density((Hs @ M).double().cpu().numpy(), density)
        rows, cols, vals = [], [], []
        for lo in range(0, int(n), chunk):
            m = min(chunk, int(n) - lo)
            rate = (gamma_half((m, k_true)) @ M).mul_(float(s))
            C = torch.poisson(rate, generator=gen)
            nz = C.nonzero(as_tuple=True)
            rows.append((nz[0] + lo).cpu().numpy())
            cols.append(nz[1].to(torch.int32).cpu().numpy())
            vals.append(C[nz].double().cpu().numpy())
        counts.append(sps.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                     shape=(int(n), g)))
    return scale_like_pyliger(counts)


def make_datasets_resident(ns, g, k_true, density=0.08, seed0=1000, device="cuda", dtype=None, chunk=32768):
    """The structured generator of ``make_datasets_device`` with everything left on the GPU: counts, the row-L1
    normalisation, the per-gene RMS scaling (same maths as ``scale_like_pyliger``) and the CSR arrays.  Returns a list
    of ``engine.DeviceMatrix`` (what the device-timed legs of bench.py feed to InmfEngine for the workloads that are too
    large to round-trip through host scipy matrices in a bench run, e.g. configs[3]: 2 M cells x 5 k genes).
    Cells without counts keep an empty row; a gene that is all zero in a dataset keeps an empty column."""
    import torch
    from .engine import DeviceMatrix

    dev = torch.device(device)
    dtype = dtype or torch.float32
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed0) - 1)

    def gamma_half(shape):
        return torch.randn(shape, device=dev, dtype=torch.float32, generator=gen).square_().mul_(0.5)

    W_star = gamma_half((k_true, g))
    out = []
    for i, n in enumerate(ns):
        n = int(n)
        gen.manual_seed(int(seed0) + i)
        M = W_star + 0.3 * gamma_half((k_true, g))
        Hs = gamma_half((2048, k_true))
        s = _depth_for_density((Hs @ M).double().cpu().numpy(), density)
        counts_per_row = torch.zeros(n, dtype=torch.int64, device=dev)
        cols, vals = [], []
        ss = torch.zeros(g, dtype=torch.float64, device=dev)
        for lo in range(0, n, chunk):
            m = min(chunk, n - lo)
            C = torch.poisson((gamma_half((m, k_true)) @ M).mul_(float(s)), generator=gen)
            rs = C.sum(dim=1, keepdim=True).clamp_(min=1.0)
            Nm = C / rs                                   # row-L1 normalised (_normalization.py:95)
            ss += Nm.double().square().sum(dim=0)
            nz = C.nonzero(as_tuple=True)
            counts_per_row[lo:lo + m] = torch.bincount(nz[0], minlength=m)
            cols.append(nz[1].to(torch.int32))
            vals.append(Nm[nz])
            del C, Nm, nz
        scaler = torch.where(ss > 0, 1.0 / torch.sqrt(ss / (n - 1)), torch.zeros_like(ss))   # _scale.py:100-101
        colidx = torch.cat(cols)
        v = (torch.cat(vals).double() * scaler[colidx.long()]).to(dtype)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts_per_row, 0, out=rowptr[1:])
        out.append(DeviceMatrix(rowptr, colidx, v, (n, g)))
    return out
