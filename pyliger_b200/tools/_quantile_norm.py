"""quantile_norm on B200: same signature, outputs and random stream as pyliger
(/root/reference/src/pyliger/tools/_quantile_norm.py:9-195 with clustering/_utilities.py:11-95).

Stages (all on the device, double precision; H stays resident between them and -- when optimize_ALS / online_iNMF of
this package produced it -- is not even uploaded):

  1. max-factor assignment of every cell on column-scaled H          inmf_qn_max_factor   (:110-125)
  2. refinement by an exact kNN graph in factor space + the           inmf_qn_knn, inmf_qn_vote
     in-place, cell-order-dependent majority vote                     (clustering/_utilities.py:11-17, 57-95)
  3. per dataset, cluster and factor: quantile alignment against      inmf_qn_align        (:133-183)
     the reference dataset (mquantiles alphap = betap = 1, tie handling, np.interp warp)

What stays on the host, by necessity: the reference subsamples clusters larger than ``max_sample`` with
``np.random.permutation`` from numpy's global legacy stream, one draw per (dataset, cluster, factor, side) in loop
order.  Those permutations are drawn here with numpy itself, in the same order (only their first ``max_sample``
entries go to the device); when no cluster exceeds ``max_sample`` the permutations cannot influence the result and are
not drawn at all.

Differences, documented: ``use_ann=True`` (annoy, absent here) is served by the same exact search; kNN ties at the
k-th distance are broken by the smaller cell index (sklearn's kd-tree order is unspecified); datasets are aligned on
one GPU.
Writes ``adata.obs["cluster"]`` and ``adata.obsm["H_norm"]`` like the reference; returns None.
"""
import numpy as np
import torch

from .. import _lib, devcache
from ..engine import _stream, require_cuda


def _device_H(H_host, device):
    """obsm["H"] on the device as a contiguous float64 (n x k) tensor: the twin left by this package's factorisations
    if there is one, an upload otherwise."""
    twin = devcache.lookup_dense(H_host, device)
    if twin is not None:
        return twin.to(torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(H_host, dtype=np.float64))).to(device)


def _members(cluster, num_clusters):
    """Cells sorted by (cluster, cell index) + offsets: memb int32[n], off int32[num_clusters + 1]."""
    order = torch.sort(cluster, stable=True)[1].to(torch.int32)
    counts = torch.bincount(cluster, minlength=num_clusters)[:num_clusters]
    off = torch.zeros(num_clusters + 1, dtype=torch.int32, device=cluster.device)
    off[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return order, off, counts.cpu().numpy()


def quantile_norm(liger_object, quantiles=50, ref_dataset=None, min_cells=20, dims_use=None, do_center=False,
                  max_sample=1000, num_trees=None, refine_knn=True, knn_k=20, use_ann=False, rand_seed=1, *, device=None):
    device = require_cuda(device)
    np.random.seed(rand_seed)
    adatas = liger_object.adata_list
    num_samples = len(adatas)
    if ref_dataset is None:
        ref = int(np.argmax([adata.shape[0] for adata in adatas]))
    else:
        ref = [i for i in range(num_samples) if adatas[i].uns["sample_name"] == ref_dataset][0]
    k_all = adatas[0].obsm["H"].shape[1]
    use = list(range(k_all)) if dims_use is None else [int(d) for d in dims_use]
    num_clusters = len(use)
    if num_clusters > 64 or k_all > 64:
        raise ValueError("pyliger_b200.quantile_norm supports at most 64 factors")
    with torch.cuda.device(device):
        st = _stream()
        dims = torch.tensor(use, dtype=torch.int32, device=device)
        Hd, clusters = [], []
        for i, adata in enumerate(adatas):
            H = _device_H(adata.obsm["H"], device)
            n = H.shape[0]
            cl = torch.empty(n, dtype=torch.int32, device=device)
            ws = torch.empty(2 * num_clusters, dtype=torch.float64, device=device)
            _lib.call("inmf_qn_max_factor", "", H, H.shape[1], n, dims, num_clusters, 1 if do_center else 0, ws, cl, st)
            if refine_knn:
                K = int(knn_k)
                knn = torch.empty(n, K, dtype=torch.int32, device=device)
                _lib.call("inmf_qn_knn", "", H, H.shape[1], n, dims, num_clusters, K, knn, st)
                _lib.call("inmf_qn_vote", "", knn, K, n, cl, st)
                del knn
            Hd.append(H)
            clusters.append(cl.to(torch.int64))
            adata.obs["cluster"] = clusters[-1].cpu().numpy()
        Hn = [h.clone() for h in Hd]  # "all H_matrix used for quantile alignment" (:130)
        memb = [_members(c, num_clusters) for c in clusters]
        m1, off1, cnt1 = memb[ref]
        for kds in range(num_samples):
            m2, off2, cnt2 = memb[kds]
            todo = []  # (cluster j, factor i) groups the reference's loops process, in its order
            for j in range(num_clusters):
                for i in use:
                    if cnt1[j] < min_cells or cnt2[j] < min_cells:
                        continue
                    todo.append((j, i))
            # np.random.permutation draws (two per group) are only needed up to the last group that subsamples
            last = -1
            for t, (j, i) in enumerate(todo):
                if cnt2[j] > max_sample or cnt1[j] > max_sample:
                    last = t
            groups, samp, soff = [], [], 0
            single = []
            for t, (j, i) in enumerate(todo):
                if cnt2[j] == 1:  # :143-145 (only reachable with min_cells <= 1)
                    single.append((j, i))
                    continue
                g = [j, i, 0, 0, 0, 0]
                if t <= last:
                    p2 = np.random.permutation(int(cnt2[j]))
                    p1 = np.random.permutation(int(cnt1[j]))
                    if cnt2[j] > max_sample:
                        g[2], g[3] = soff, max_sample
                        samp.append(p2[:max_sample])
                        soff += max_sample
                    if cnt1[j] > max_sample:
                        g[4], g[5] = soff, max_sample
                        samp.append(p1[:max_sample])
                        soff += max_sample
                groups.append(g)
            for j, i in single:
                cells1 = m1[int(off1[j]):int(off1[j + 1])].long()
                cells2 = m2[int(off2[j]):int(off2[j + 1])].long()
                Hn[kds][cells2, i] = Hn[ref][cells1, i].mean()
            if groups:
                gt = torch.tensor(groups, dtype=torch.int64, device=device)
                sp = (torch.from_numpy(np.concatenate(samp).astype(np.int32)).to(device) if samp
                      else torch.zeros(1, dtype=torch.int32, device=device))
                _lib.call("inmf_qn_align", "", gt, len(groups), Hn[kds], Hn[kds].shape[1], Hn[ref], Hn[ref].shape[1],
                          m2, off2, m1, off1, sp, int(quantiles), int(max_sample), st)
            adatas[kds].obsm["H_norm"] = Hn[kds].cpu().numpy()
    return None
