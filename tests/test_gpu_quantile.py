"""GPU: quantile_norm (max-factor assignment, exact kNN, sequential vote, quantile alignment) against vectors produced
by executing the unmodified reference, and the kernels one by one against numpy."""
import numpy as np
import pytest
import torch

from conftest import QUANTILE_CASES, relerr
from oracle import quantile_oracle as qo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _liger_with_H(Hs):
    import scipy.sparse as sps
    from pyliger_b200 import liger_from_matrices
    lig = liger_from_matrices([sps.csr_matrix((h.shape[0], 5)) for h in Hs])
    for a, h in zip(lig.adata_list, Hs):
        a.obsm["H"] = h.copy()
    return lig


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_quantile_norm_vs_reference(golden_quantile, tag):
    from pyliger_b200 import quantile_norm
    d = golden_quantile
    Hs = [d["H%d" % i] for i in range(3)]
    lig = _liger_with_H(Hs)
    assert quantile_norm(lig, **QUANTILE_CASES[tag]) is None
    for i, a in enumerate(lig.adata_list):
        assert np.array_equal(np.asarray(a.obs["cluster"]), d["%s_cluster%d" % (tag, i)])
        np.testing.assert_allclose(a.obsm["H_norm"], d["%s_Hnorm%d" % (tag, i)], rtol=1e-11, atol=1e-13)
        assert np.array_equal(a.obsm["H"], Hs[i])  # H itself is not touched


def test_knn_and_vote_kernels_against_numpy():
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    rng = np.random.default_rng(5)
    for n, m, K, ld in ((1000, 7, 20, 9), (333, 30, 5, 30), (2100, 12, 40, 16)):
        H = rng.gamma(0.5, 1.0, (n, ld))
        dims = np.sort(rng.choice(ld, m, replace=False)).astype(np.int32)
        Hd = torch.from_numpy(H).cuda()
        knn = torch.empty(n, K, dtype=torch.int32, device="cuda")
        _lib.call("inmf_qn_knn", "", Hd, ld, n, torch.from_numpy(dims).cuda(), m, K, knn, _stream())
        want = qo.knn_exact(H[:, dims], K)
        got = knn.cpu().numpy()
        assert np.array_equal(np.sort(got, axis=1), np.sort(want, axis=1))
        cl = rng.integers(0, m, n)
        cd = torch.from_numpy(cl.astype(np.int32)).cuda()
        _lib.call("inmf_qn_vote", "", knn, K, n, cd, _stream())
        assert np.array_equal(cd.cpu().numpy(), qo.cluster_vote(cl.copy(), got, K))


def test_quantile_norm_after_optimize_ALS_reuses_device_H():
    """H written back by optimize_ALS keeps a device twin: quantile_norm finds it (no upload) and the result equals the
    oracle on the host copies; editing H in place is refused (frozen), a copy takes the upload path."""
    from pyliger_b200 import devcache, liger_from_matrices, optimize_ALS, quantile_norm
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([900, 1200], 200, 8, density=0.12, seed0=2700)
    lig = liger_from_matrices(mats)
    optimize_ALS(lig, 8, 5.0, 1e-6, 6, dtype="float64", progress=False)
    dev = torch.device("cuda", torch.cuda.current_device())
    assert all(devcache.lookup_dense(a.obsm["H"], dev) is not None for a in lig.adata_list)
    with pytest.raises(ValueError):
        lig.adata_list[0].obsm["H"][0, 0] = 1.0
    quantile_norm(lig, min_cells=10, knn_k=10, max_sample=80)
    clusters, Hn = qo.quantile_norm([a.obsm["H"] for a in lig.adata_list], min_cells=10, knn_k=10, max_sample=80)
    for i, a in enumerate(lig.adata_list):
        assert np.array_equal(np.asarray(a.obs["cluster"]), clusters[i])
        assert relerr(a.obsm["H_norm"], Hn[i]) < 1e-11
    lig2 = _liger_with_H([a.obsm["H"] for a in lig.adata_list])  # copies: no twins
    quantile_norm(lig2, min_cells=10, knn_k=10, max_sample=80)
    for a, b in zip(lig.adata_list, lig2.adata_list):
        assert np.array_equal(a.obsm["H_norm"], b.obsm["H_norm"])
