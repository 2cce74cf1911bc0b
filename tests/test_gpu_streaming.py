"""GPU: out-of-core online_iNMF (cells streamed from an on-disk CSR store or from host matrices) and the backed
preprocessing stages against the resident / in-memory paths, which are themselves pinned to the reference."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _mats():
    from pyliger_b200.synth import make_datasets
    return make_datasets([700, 520, 310], 160, 8, density=0.12, seed0=5200)


def _resident(mats, **kw):
    from pyliger_b200 import liger_from_matrices, online_iNMF
    lig = liger_from_matrices(mats)
    online_iNMF(lig, verbose=False, **kw)
    return lig


def _compare(a, b, tol):
    for x, y in zip(a.adata_list, b.adata_list):
        assert relerr(x.obsm["H"], y.obsm["H"]) < tol
        assert relerr(x.varm["V"], y.varm["V"]) < tol
        assert relerr(x.varm["B"], y.varm["B"]) < tol
        assert relerr(x.uns["A"], y.uns["A"]) < tol
    assert relerr(a.W, b.W) < tol


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-3)])
def test_streamed_host_matrices_equal_resident(dtype, tol):
    """stream=True: same mini-batches, same updates, X never resident (CSR row kernels on streamed slots) -- three
    epochs with wrap-around and ragged mini-batches."""
    from pyliger_b200 import liger_from_matrices, online_iNMF
    mats = _mats()
    kw = dict(k=8, value_lambda=5.0, max_epochs=3, miniBatch_size=333, miniBatch_max_iters=2, rand_seed=2, dtype=dtype)
    want = _resident(mats, **kw)
    lig = liger_from_matrices(mats)
    online_iNMF(lig, verbose=False, stream=True, **kw)
    _compare(lig, want, tol)


def test_backed_store_scenarios_equal_in_memory(tmp_path, monkeypatch):
    """Backed AnnData (scale_data in ./results/<sample>.csr, float32 on disk for one of them): scenario 1, then
    scenario 2 (new backed dataset) and scenario 3 (projection of a backed dataset)."""
    from pyliger_b200 import Liger, online_iNMF
    from pyliger_b200.containers import AnnDataLike
    from pyliger_b200.io import store_path, write_store
    monkeypatch.chdir(tmp_path)
    mats = _mats()
    kw = dict(k=8, value_lambda=5.0, max_epochs=2, miniBatch_size=400, rand_seed=1)

    def backed(m, name):
        a = AnnDataLike(None, name, X=None, shape=m.shape)
        a.isbacked = True
        write_store(store_path(a), "scale_data", m, chunk_size=97, mode="w")
        return a

    want = _resident(mats[:2], **kw)
    lig = Liger([backed(mats[0], "s0"), backed(mats[1], "s1")], var_genes=["g%d" % j for j in range(mats[0].shape[1])])
    online_iNMF(lig, verbose=False, **kw)
    _compare(lig, want, 1e-10)
    # scenario 2: a new backed dataset refines W; scenario 3: projection
    from pyliger_b200 import liger_from_matrices
    new_mem = liger_from_matrices([mats[2]]).adata_list[0]
    online_iNMF(want, X_new=[new_mem], verbose=False, **kw)
    online_iNMF(lig, X_new=[backed(mats[2], "s2")], verbose=False, **kw)
    _compare(lig, want, 1e-9)
    proj_mem = liger_from_matrices([mats[1][:200]]).adata_list[0]
    online_iNMF(want, X_new=[proj_mem], projection=True, k=8, verbose=False)
    online_iNMF(lig, X_new=[backed(mats[1][:200], "s3")], projection=True, k=8, verbose=False)
    assert relerr(lig.adata_list[-1].obsm["H"], want.adata_list[-1].obsm["H"]) < 1e-10


def test_backed_preprocessing_equals_in_memory_and_raw_X_new(tmp_path, monkeypatch):
    """Chunked normalize / scale_not_center against the store == the in-memory stages; online_iNMF preprocesses a raw
    X_new (no scale_data) the way the reference does (_online_iNMF.py:604-620)."""
    import scipy.sparse as sps
    from pyliger_b200 import liger_from_raw, normalize, online_iNMF, scale_not_center
    from pyliger_b200.io import CSRStore
    from pyliger_b200.preprocessing._backed import create_store_using_adata, initialization_online
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(3)
    raws = [sps.csr_matrix(rng.poisson(0.3, (n, 90)).astype(np.float64)) for n in (260, 190)]
    for r in raws:  # every cell and gene expressed
        r[:, 0] = 1.0
    raws = [sps.csr_matrix(r) for r in raws]
    var_genes = ["g%d" % j for j in range(5, 70)]
    mem = liger_from_raw(raws, names=["a", "b"])
    mem.var_genes = var_genes
    normalize(mem)
    scale_not_center(mem)
    bk = liger_from_raw(raws, names=["a", "b"])
    bk.var_genes = var_genes
    for i, a in enumerate(bk.adata_list):
        create_store_using_adata(a, 64)
        bk.adata_list[i] = initialization_online(a, 64, remove_missing=True)
    normalize(bk, chunk_size=64)
    scale_not_center(bk, chunk_size=64)
    for a, b in zip(mem.adata_list, bk.adata_list):
        assert b.isbacked and b.shape == a.shape
        with CSRStore(b.uns["backed_path"], "r") as f:
            assert f.keys() == ["norm_data", "raw_data", "scale_data"]
            got = f["scale_data"][:]
        want = a.layers["scale_data"]
        assert got.shape == want.shape and np.array_equal(got.indices, want.indices)
        assert np.array_equal(got.indptr, want.indptr)
        # (the per-gene sums are accumulated in a different order, chunk by chunk: last-ulp differences in the scaler)
        np.testing.assert_allclose(got.data, want.data, rtol=1e-12)
        np.testing.assert_allclose(b.var["norm_sum_sq"].to_numpy(), a.var["norm_sum_sq"].to_numpy(), rtol=1e-12)
    kw = dict(k=6, value_lambda=5.0, max_epochs=2, miniBatch_size=150, rand_seed=1, verbose=False)
    online_iNMF(mem, **kw)
    online_iNMF(bk, **kw)
    assert relerr(bk.W, mem.W) < 1e-10
    # projection of a RAW dataset: online_iNMF preprocesses it into ./results/c.csr
    raw_c = sps.csr_matrix(rng.poisson(0.3, (120, 90)).astype(np.float64))
    raw_c[:, 0] = 1.0
    raw_c = sps.csr_matrix(raw_c)
    new_mem = liger_from_raw([raw_c], names=["c"])
    new_mem.var_genes = var_genes
    normalize(new_mem)
    scale_not_center(new_mem)
    online_iNMF(mem, X_new=[new_mem.adata_list[0]], projection=True, k=6, verbose=False)
    new_raw = liger_from_raw([raw_c], names=["c"]).adata_list[0]
    online_iNMF(bk, X_new=[new_raw], projection=True, k=6, verbose=False)
    assert bk.adata_list[-1].isbacked
    assert relerr(bk.adata_list[-1].obsm["H"], mem.adata_list[-1].obsm["H"]) < 1e-10
