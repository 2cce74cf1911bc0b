"""CPU: pin the numpy oracle against vectors produced by executing the unmodified reference
(tests/golden/make_golden.py) and against the KATs of SURVEY.md Appendix C."""
import numpy as np
import pytest
import scipy.optimize

from conftest import csr_unpack, relerr
from oracle import inmf_oracle as orc


def test_kat1_literal():
    X, info = orc.nnlsm_blockpivot(np.array([[1.0, 0], [0, 1], [1, 1]]),
                                   np.array([[1.0, 2], [-1, 0.5], [0, 3]]))
    np.testing.assert_allclose(X, [[0.5, 2.1666666666667], [0, 0.6666666666667]], atol=1e-12)
    np.testing.assert_allclose(info[1], [[0, 0], [1.5, 0]], atol=1e-12)
    assert info[0] is True and info[2:] == (2, 4, 0)


def test_kat2_literal():
    np.random.seed(0)
    A = np.random.uniform(0, 2, (8, 4))
    B = np.random.uniform(-1, 2, (8, 3))
    X, info = orc.nnlsm_blockpivot(A, B)
    want = [[0, 0, 0], [0, 0.251137583019, 0], [0.591911405486, 0.298163894555, 0], [0, 0, 0.1677598723]]
    np.testing.assert_allclose(X, want, atol=1e-10)
    assert info[2:] == (4, 11, 0)


@pytest.mark.parametrize("tag", ["kat1", "kat2", "iid", "hard"])
def test_bpp_raw_inputs(golden_bpp, tag):
    d = golden_bpp
    X, info = orc.nnlsm_blockpivot(d[tag + "_A"], d[tag + "_B"])
    np.testing.assert_allclose(X, d[tag + "_X"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(info[1], d[tag + "_Y"], rtol=1e-7, atol=1e-9)
    want = d[tag + "_info"]
    assert [int(info[0]), info[2], info[3], info[4]] == [int(v) for v in want]


def test_bpp_product_inputs(golden_bpp):
    d = golden_bpp
    for tag in d["cases"]:
        tag = str(tag)
        X, info = orc.nnlsm_blockpivot(d[tag + "_G"], d[tag + "_R"], is_input_prod=True)
        np.testing.assert_allclose(X, d[tag + "_X"], rtol=1e-9, atol=1e-11)
        assert (X > 0).tolist() == (d[tag + "_X"] > 0).tolist()
        want = d[tag + "_info"]
        assert [int(info[0]), info[2], info[3], info[4]] == [int(v) for v in want]


def test_bpp_warm_start(golden_bpp):
    d = golden_bpp
    X, info = orc.nnlsm_blockpivot(d["iid_A"], d["iid_B"], init=d["iid_init"])
    np.testing.assert_allclose(X, d["iid_init_X"], rtol=1e-9, atol=1e-11)
    # and the warm start reaches the same (unique) solution as the cold start
    np.testing.assert_allclose(X, d["iid_X"], rtol=1e-8, atol=1e-10)


def test_bpp_matches_scipy_nnls(golden_bpp):
    d = golden_bpp
    A, B = d["iid_A"], d["iid_B"]
    X, _ = orc.nnlsm_blockpivot(A, B)
    for c in range(0, B.shape[1], 7):
        ref, _ = scipy.optimize.nnls(A, B[:, c])
        np.testing.assert_allclose(X[:, c], ref, atol=1e-10)
    assert orc.kkt_violation(A.T @ A, A.T @ B, X) < 1e-10


def _als_inputs(d, n):
    return [csr_unpack(d, "X%d" % i) for i in range(n)]


def test_als_trajectory(golden_als):
    d = golden_als
    k, lam, thresh, T, seed = d["params"]
    out = orc.als(_als_inputs(d, 2), int(k), lam, thresh, int(T), rand_seed=int(seed))
    np.testing.assert_allclose(out["objectives"], d["objectives"], rtol=1e-9)
    assert relerr(out["W"].T, d["W"]) < 1e-7
    for i in range(2):
        assert relerr(out["H"][i], d["H%d" % i]) < 1e-7
        assert relerr(out["V"][i].T, d["V%d" % i]) < 1e-7


def test_als_stats_objective_equals_dense(golden_als):
    d = golden_als
    k, lam, thresh, T, seed = d["params"]
    out = orc.als(_als_inputs(d, 2), int(k), lam, thresh, 4, rand_seed=int(seed), dense_objective=False)
    np.testing.assert_allclose(out["objectives"][1:], d["objectives"][1:5], rtol=1e-10)


def test_als_three_datasets_nrep(golden_als3):
    d = golden_als3
    k, lam, thresh, T, seed, nrep = d["params"]
    out = orc.als(_als_inputs(d, 3), int(k), lam, thresh, int(T), nrep=int(nrep), rand_seed=int(seed))
    assert abs(out["objective"] - float(d["objective"])) / float(d["objective"]) < 1e-9
    assert relerr(out["W"].T, d["W"]) < 1e-6
    for i in range(3):
        assert relerr(out["H"][i], d["H%d" % i]) < 1e-6


def test_als_value_errors(golden_als):
    mats = _als_inputs(golden_als, 2)
    with pytest.raises(ValueError):
        orc.als(mats, 260)
    with pytest.raises(ValueError):
        orc.als(mats, 150)


def test_online_helpers(golden_online):
    d = golden_online
    np.testing.assert_allclose(orc.init_W(3, 2, 1), d["kat4_initW"], rtol=1e-14)
    np.testing.assert_allclose(
        orc.init_W(3, 2, 1),
        [[0.9432939967595, 0.9157005169369], [2.587131550980e-04, 0.3843352483030],
         [0.3319583840560, 0.1173839009124]], rtol=1e-9)
    steps = orc.minibatch_rows(4, 2500, 1, 1000, 7000)
    assert np.array([[list(p) for p in s] for s in steps]).tolist() == d["kat5_idx"].tolist()
    assert steps[2] == [(5000, 6000), (6000, 7000), (0, 500)]
    flat = [(i, lo, hi) for i, s in enumerate(orc.minibatch_rows(7, 130, 2, 100, 333)) for lo, hi in s]
    assert flat == [tuple(r) for r in d["idx2_flat"].tolist()]
    X0 = csr_unpack(d, "X0")
    np.testing.assert_allclose(orc.init_V_online(420, 5, X0, 1), d["s1_initV0"], rtol=1e-13)


def test_update_A_B_kat6(golden_online):
    d = golden_online
    A, B = np.eye(2), np.ones((3, 2))
    Ao, Bo = np.zeros((2, 2)), np.zeros((3, 2))
    Hm = np.array([[1.0, 2], [0, 1]])
    Xm = np.arange(6, dtype=np.float64).reshape(3, 2)
    for s, (it, ep, epp) in enumerate(((0, 0, 0), (1, 0, 0), (2, 1, 0), (3, 1, 1))):
        A, B, Ao, Bo = orc.update_A_B(A, B, Ao, Bo, Hm, Xm, 2, it, ep, epp)
        np.testing.assert_allclose(A, d["kat6_A"][s], rtol=1e-14)
        np.testing.assert_allclose(B, d["kat6_B"][s], rtol=1e-14)
        np.testing.assert_allclose(Ao, d["kat6_Aold"][s], rtol=1e-14)
        np.testing.assert_allclose(Bo, d["kat6_Bold"][s], rtol=1e-14)
    np.testing.assert_allclose(d["kat6_A"][0].ravel(), [2.5, 1, 1, 0.5])


def test_hals_sweep(golden_online):
    d = golden_online
    A = list(d["hals_A"]); B = list(d["hals_B"])
    W = orc.hals_W(A, B, d["hals_W0"].copy(), [v.copy() for v in d["hals_V0"]])
    np.testing.assert_allclose(W, d["hals_W1"], rtol=1e-12)
    V = orc.hals_V(A, B, W.copy(), [v.copy() for v in d["hals_V0"]], 5.0)
    np.testing.assert_allclose(np.array(V), d["hals_V1"], rtol=1e-12)


def test_online_scenario1(golden_online):
    d = golden_online
    k, lam, ep, mbi, mb, chunk, seed = d["s1_params"]
    mats = [csr_unpack(d, "X0"), csr_unpack(d, "X1")]
    out = orc.online(mats, int(k), lam, int(ep), int(mbi), int(mb), int(chunk), int(seed))
    assert relerr(out["W"], d["s1_W"]) < 1e-8
    for i in range(2):
        assert relerr(out["H"][i], d["s1_H%d" % i]) < 1e-7
        assert relerr(out["V"][i], d["s1_V%d" % i]) < 1e-8
        assert relerr(out["A"][i], d["s1_A%d" % i]) < 1e-8
        assert relerr(out["B"][i], d["s1_B%d" % i]) < 1e-8


def test_online_scenario1_two_inner_iters(golden_online):
    d = golden_online
    out = orc.online([csr_unpack(d, "X0")], 4, 3.0, 2, 2, 100, 64, 3)
    assert relerr(out["W"], d["s1b_W"]) < 1e-8
    assert relerr(out["H"][0], d["s1b_H"]) < 1e-7
    assert relerr(out["A"][0], d["s1b_A"]) < 1e-8


def test_online_scenario2_refine(golden_online):
    d = golden_online
    mats = [csr_unpack(d, "X0"), csr_unpack(d, "X1")]
    prev = dict(W=d["s1_W"].copy(), V=[d["s1_V0"].copy(), d["s1_V1"].copy()],
                A=[d["s1_A0"].copy(), d["s1_A1"].copy()], B=[d["s1_B0"].copy(), d["s1_B1"].copy()])
    out = orc.online_refine(mats, prev, [csr_unpack(d, "Xnew")], k=5, value_lambda=5.0, max_epochs=2,
                            miniBatch_size=80, h5_chunk_size=100, rand_seed=1)
    assert relerr(out["W"], d["s2_W"]) < 1e-8
    assert relerr(out["new"][0]["H"], d["s2_Hnew"]) < 1e-7
    assert relerr(out["new"][0]["V"], d["s2_Vnew"]) < 1e-8
    assert relerr(out["new"][0]["A"], d["s2_Anew"]) < 1e-8
    for i in range(2):
        assert relerr(out["H_old"][i], d["s2_Hold%d" % i]) < 1e-7


def test_online_scenario3_projection(golden_online):
    d = golden_online
    out = orc.projection([csr_unpack(d, "Xnew")], d["s1_W"], 5, 100)
    assert relerr(out["H"][0], d["s3_H"]) < 1e-8
    assert np.all(out["V"][0] == 0) and out["V"][0].shape == d["s3_V"].shape


# ------------------------------------------------------------------ normalize / scale_not_center (8(f) rank 1)
def test_preprocess_oracle_matches_reference(golden_pre):
    from oracle import preprocess_oracle as pre
    d = golden_pre
    var_idx = np.isin(d["var_names"], d["var_genes"])
    for i in range(2):
        raw = csr_unpack(d, "raw%d" % i)
        assert not pre.missing_rows(raw).any()
        norm, s1, s2 = pre.normalize_matrix(raw)
        ref = csr_unpack(d, "norm%d" % i)
        assert (norm.indices == ref.indices).all() and (norm.indptr == ref.indptr).all()
        np.testing.assert_allclose(norm.data, ref.data, rtol=1e-15, atol=0)
        np.testing.assert_allclose(s1, d["norm_sum%d" % i], rtol=1e-13)
        np.testing.assert_allclose(s2, d["norm_sum_sq%d" % i], rtol=1e-13)
        sc = pre.scale_matrix(norm, s2, var_idx)
        ref = csr_unpack(d, "scale%d" % i)
        assert sc.shape == ref.shape and (sc.indices == ref.indices).all() and (sc.indptr == ref.indptr).all()
        np.testing.assert_allclose(sc.data, ref.data, rtol=1e-13, atol=0)
        # per-gene RMS of scale_data is 1 for every gene that is expressed (the property ALS relies on)
        ms = np.asarray(sc.power(2).sum(axis=0)).ravel() / (sc.shape[0] - 1)
        expressed = np.asarray((sc != 0).sum(axis=0)).ravel() > 0
        np.testing.assert_allclose(ms[expressed], 1.0, rtol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_inmf_hals_oracle_matches_reference(golden_hals, tag):
    """oracle.inmf_hals against the executed reference iNMF_HALS (_iNMF_HALS.py:60-161, run on a duck-typed liger
    object because the reference's own Liger class makes its final attribute assignment fail)."""
    d = golden_hals
    N, k, lam, thresh, max_iters, nrep, seed = d[tag + "_params"]
    N, k, max_iters, nrep, seed = int(N), int(k), int(max_iters), int(nrep), int(seed)
    mats = [csr_unpack(d, "%s_X%d" % (tag, i)) for i in range(N)]
    got = orc.inmf_hals(mats, k, lam, thresh, max_iters, nrep, seed)
    objs = np.concatenate(got["objectives"])
    np.testing.assert_allclose(objs, d[tag + "_objectives"], rtol=1e-9)
    for i in range(N):
        assert relerr(got["H"][i], d["%s_H%d" % (tag, i)]) < 1e-8
        assert relerr(got["V"][i], d["%s_V%d" % (tag, i)]) < 1e-8
    assert relerr(got["W"], d[tag + "_W"]) < 1e-8


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_quantile_norm_oracle_matches_reference(golden_quantile, tag):
    """oracle.quantile_norm (brute-force kNN, sequential vote, numpy-global permutation stream) against the executed
    reference: clusters exact, H_norm to 1e-12."""
    from conftest import QUANTILE_CASES
    from oracle import quantile_oracle as qo
    d = golden_quantile
    Hs = [d["H%d" % i] for i in range(3)]
    kw = dict(QUANTILE_CASES[tag])
    clusters, Hn = qo.quantile_norm(Hs, **kw)
    for i in range(3):
        assert np.array_equal(clusters[i], d["%s_cluster%d" % (tag, i)])
        np.testing.assert_allclose(Hn[i], d["%s_Hnorm%d" % (tag, i)], rtol=1e-12, atol=1e-14)
