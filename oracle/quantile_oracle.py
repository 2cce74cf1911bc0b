"""TEST INFRASTRUCTURE ONLY -- CPU restatement of pyliger's quantile_norm
(/root/reference/src/pyliger/tools/_quantile_norm.py:9-195, clustering/_utilities.py:11-95).

Restated: the max-factor assignment, the kNN refinement (exact neighbours by brute force instead of sklearn's
kd-tree -- the same neighbour SETS whenever the k-th distance is not tied -- and the sequential, in-place
``cluster_vote``), and the quantile alignment loop with the reference's numpy-global random stream.  The quantile and
interpolation primitives are the third-party functions the reference itself calls (scipy.stats.mstats.mquantiles,
scipy.interpolate.interp1d).  Parity is PINNED by ``tests/golden/quantile.npz``, produced by executing the unmodified
reference (``tests/golden/make_golden.py quantile``; numba / annoy / pynndescent / igraph, which the reference module
imports but this path does not need, are stubbed in ``oracle/ref_shims.py``).
Only tests import this module.
"""
import numpy as np
from scipy import interpolate
from scipy.stats.mstats import mquantiles


def knn_exact(H, k):
    """Indices of the k nearest rows (Euclidean, self included), ties by the smaller index."""
    n = H.shape[0]
    out = np.empty((n, k), dtype=np.int64)
    for lo in range(0, n, 512):
        blk = H[lo:lo + 512]
        d = ((blk[:, None, :] - H[None, :, :]) ** 2).sum(axis=2)
        out[lo:lo + 512] = np.argsort(d, axis=1, kind="stable")[:, :k]
    return out


def cluster_vote(clusts, knn, k):
    """clustering/_utilities.py:57-79: in place, in cell order; most frequent label, ties to the largest label."""
    for i in range(knn.shape[0]):
        counts = {}
        for j in range(k):
            c = clusts[knn[i, j]]
            counts[c] = counts.get(c, 0) + 1
        best, best_count = -1, 0
        for key, val in counts.items():
            if val > best_count or (val == best_count and key > best):
                best, best_count = key, val
        clusts[i] = best
    return clusts


def max_factor(H, do_center=False):
    """_quantile_norm.py:112-121."""
    if do_center:
        S = (H - np.mean(H, axis=0)) / np.std(H, axis=0, ddof=1)
    else:
        S = H / np.sqrt(np.sum(np.square(H), axis=0) / (H.shape[0] - 1))
    return np.argmax(S, axis=1)


def quantile_norm(Hs_in, quantiles=50, ref=None, min_cells=20, dims_use=None, do_center=False, max_sample=1000,
                  refine_knn=True, knn_k=20, rand_seed=1):
    """Returns (clusters list, H_norm list).  Hs_in: list of n_i x k arrays."""
    np.random.seed(rand_seed)
    N = len(Hs_in)
    if ref is None:
        ref = int(np.argmax([h.shape[0] for h in Hs_in]))
    use = list(range(Hs_in[0].shape[1])) if dims_use is None else list(dims_use)
    Hs = [h[:, use] for h in Hs_in]
    num_clusters = Hs[ref].shape[1]
    clusters = []
    for i in range(N):
        cl = max_factor(Hs[i], do_center)
        if refine_knn:
            cl = cluster_vote(cl, knn_exact(Hs[i], knn_k), knn_k)
        clusters.append(cl)
    Hs = [np.copy(h) for h in Hs_in]
    for k in range(N):
        for j in range(num_clusters):
            cells2 = clusters[k] == j
            cells1 = clusters[ref] == j
            for i in use:
                n2, n1 = np.sum(cells2), np.sum(cells1)
                if n1 < min_cells or n2 < min_cells:
                    continue
                if n2 == 1:
                    Hs[k][cells2, i] = np.mean(Hs[ref][cells1, i])
                    continue
                q2 = mquantiles(np.random.permutation(Hs[k][cells2, i])[0:min(n2, max_sample)],
                                np.linspace(0, 1, num=quantiles + 1), alphap=1, betap=1)
                hi, lo = np.max(Hs[k][cells2, i]), np.min(Hs[k][cells2, i])
                if q2[-1] < hi:
                    q2[-1] = hi
                if q2[0] > lo:
                    q2[0] = lo
                q1 = mquantiles(np.random.permutation(Hs[ref][cells1, i])[0:min(n1, max_sample)],
                                np.linspace(0, 1, num=quantiles + 1), alphap=1, betap=1)
                if np.sum(q1) == 0 or np.sum(q2) == 0 or len(np.unique(q1)) < 2 or len(np.unique(q2)) < 2:
                    new = np.repeat(0, n2)
                else:
                    z = q2 == 0
                    if np.sum(z) > 0:
                        q1[z] = np.mean(q1[z])
                    new = interpolate.interp1d(q2, q1)(Hs[k][cells2, i])
                Hs[k][cells2, i] = new
    return clusters, Hs
