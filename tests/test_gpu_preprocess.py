"""normalize / scale_not_center on the GPU against the executed reference (tests/golden/preprocess.npz) and
the numpy oracle; then the chain normalize -> scale_not_center -> optimize_ALS without re-uploading X."""
import numpy as np
import pytest
import scipy.sparse as sps

from conftest import csr_unpack, relerr
from oracle import inmf_oracle as orc
from oracle import preprocess_oracle as pre

pytestmark = pytest.mark.gpu


def _liger_from_golden(d):
    from pyliger_b200 import liger_from_raw
    lig = liger_from_raw([csr_unpack(d, "raw0"), csr_unpack(d, "raw1")], ["s0", "s1"],
                         var_names=list(d["var_names"]))
    return lig


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-13), ("float32", 2e-6)])
def test_normalize_and_scale_match_reference(golden_pre, dtype, tol):
    import pyliger_b200 as pb
    d = golden_pre
    lig = _liger_from_golden(d)
    assert pb.normalize(lig, dtype=dtype) is None
    for i, a in enumerate(lig.adata_list):
        ref = csr_unpack(d, "norm%d" % i)
        got = a.layers["norm_data"]
        assert got.dtype == np.float64 and got.shape == ref.shape
        assert (got.indices == ref.indices).all() and (got.indptr == ref.indptr).all()
        np.testing.assert_allclose(got.data, ref.data, rtol=tol)
        for col in ("norm_sum", "norm_sum_sq", "norm_mean"):
            np.testing.assert_allclose(a.var[col].to_numpy(), d["%s%d" % (col, i)], rtol=tol, atol=1e-300)
    lig.var_genes = list(d["var_genes"])
    with np.errstate(divide="ignore"):
        assert pb.scale_not_center(lig, dtype=dtype) is None
    for i, a in enumerate(lig.adata_list):
        ref = csr_unpack(d, "scale%d" % i)
        got = a.layers["scale_data"]
        assert got.shape == ref.shape == a.shape
        assert (got.indices == ref.indices).all() and (got.indptr == ref.indptr).all()
        np.testing.assert_allclose(got.data, ref.data, rtol=tol)
        assert list(a.var.index) == list(d["scaled_var_names%d" % i])
        assert tuple(a.raw.shape) == tuple(d["raw_shape%d" % i])
        assert a.layers["norm_data"].shape == a.shape  # sliced along with the genes, as AnnData does


def test_normalize_reports_cells_without_counts(golden_pre, capsys):
    import pyliger_b200 as pb
    raw = csr_unpack(golden_pre, "raw0").tolil()
    raw[5, :] = 0
    lig = pb.liger_from_raw([raw.tocsr()], ["s0"])
    with pytest.raises(ValueError):
        pb.normalize(lig)
    assert "Removing 1 cells not expressing any genes in s0." in capsys.readouterr().out
    pb.normalize(lig, remove_missing=False)  # zero rows stay zero (sklearn l1 normalisation)
    assert lig.adata_list[0].layers["norm_data"][5].nnz == 0


def test_medium_vs_oracle_and_chain_into_als():
    """20k x 1500 counts: both stages against the oracle, then optimize_ALS consumes the device-resident
    scale_data (no second upload) and reproduces the run on an identical host-only matrix."""
    import pyliger_b200 as pb
    rng = np.random.default_rng(9)
    g, k = 1500, 12
    raws = []
    for n in (12000, 8000):
        M = sps.random(n, g, density=0.05, random_state=int(rng.integers(1 << 30)), format="csr")
        M.data = np.ceil(M.data * 6)
        M = M + sps.csr_matrix((np.ones(n), (np.arange(n), rng.integers(0, g, n))), shape=(n, g))
        raws.append(sps.csr_matrix(M))
    names = ["g%d" % j for j in range(g)]
    lig = pb.liger_from_raw(raws, var_names=names)
    pb.normalize(lig)
    keep = np.sort(rng.choice(g, 1000, replace=False))
    lig.var_genes = [names[j] for j in keep]
    mask = np.zeros(g, bool); mask[keep] = True
    norms = [a.layers["norm_data"] for a in lig.adata_list]
    sumsq = [a.var["norm_sum_sq"].to_numpy() for a in lig.adata_list]
    pb.scale_not_center(lig)
    for i, a in enumerate(lig.adata_list):
        no, s1, s2 = pre.normalize_matrix(raws[i])
        np.testing.assert_allclose(norms[i].data, no.data, rtol=1e-13)
        np.testing.assert_allclose(sumsq[i], s2, rtol=1e-12)
        so = pre.scale_matrix(no, s2, mask)
        got = a.layers["scale_data"]
        assert (got.indices == so.indices).all() and (got.indptr == so.indptr).all()
        np.testing.assert_allclose(got.data, so.data, rtol=1e-12)
    info = pb.optimize_ALS(lig, k, 5.0, 1e-6, 4, dtype="float64", progress=False, return_info=True)
    x_bytes = sum(a.layers["scale_data"].data.nbytes + a.layers["scale_data"].indices.nbytes
                  for a in lig.adata_list)
    assert info["h2d_bytes"] < 0.2 * x_bytes  # only row pointers and the initial factors went up
    lig2 = pb.liger_from_matrices([a.layers["scale_data"].copy() for a in lig.adata_list])
    info2 = pb.optimize_ALS(lig2, k, 5.0, 1e-6, 4, dtype="float64", progress=False, return_info=True)
    assert info2["h2d_bytes"] > x_bytes
    np.testing.assert_allclose(info["objectives"], info2["objectives"], rtol=1e-12)
    assert relerr(lig.adata_list[0].varm["W"], lig2.adata_list[0].varm["W"]) < 1e-10
    out = orc.als([a.layers["scale_data"] for a in lig.adata_list], k, 5.0, 1e-6, 4, rand_seed=1)
    assert abs(info["objectives"][0][-1] - out["objectives"][-1]) <= 1e-5 * abs(out["objectives"][-1])
