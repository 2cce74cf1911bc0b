"""GPU parity tests proper: the pyliger-shaped API of pyliger_b200 (CUDA, through the C ABI) against
(a) golden vectors produced by the unmodified reference and (b) the numpy oracle on seeded inputs.

Tolerances (north star): per-iteration objective within 1e-5 relative in fp64 mode and 1e-3 in fp32
mode; final H/W/V relative Frobenius error <= 1e-6 (fp64) / 1e-2 (fp32).  The fp64 asserts below are
tighter than required (observed ~1e-12).
"""
import numpy as np
import pytest
import scipy.optimize
import scipy.sparse as sps
import torch

from conftest import csr_unpack, relerr
from oracle import inmf_oracle as orc

pytestmark = pytest.mark.gpu

OBJ_TOL = {"float64": 1e-9, "float32": 1e-3}
FAC_TOL = {"float64": 1e-6, "float32": 1e-2}


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _liger(mats):
    from pyliger_b200 import liger_from_matrices
    return liger_from_matrices(mats)


# ----------------------------------------------------------------------------------- nnlsm_blockpivot
@pytest.mark.parametrize("tag", ["kat1", "kat2", "iid", "hard"])
def test_bpp_raw_inputs_fp64(golden_bpp, tag):
    from pyliger_b200 import nnlsm_blockpivot
    d = golden_bpp
    X, info = nnlsm_blockpivot(d[tag + "_A"], d[tag + "_B"])
    ok, Y, n_chol, n_eq, n_backup = info
    assert X.shape == d[tag + "_X"].shape and Y.shape == X.shape
    tol = 1e-6 if tag == "hard" else 1e-9  # "hard" is ill-conditioned (cond ~ 1e7) on purpose
    np.testing.assert_allclose(X, d[tag + "_X"], rtol=tol, atol=tol * 1e-2 * max(1.0, np.abs(d[tag + "_X"]).max()))
    assert ok == bool(d[tag + "_info"][0])
    if tag != "hard":
        assert n_backup == int(d[tag + "_info"][3])
        np.testing.assert_allclose(Y, d[tag + "_Y"], rtol=1e-7, atol=1e-9)
        assert (X > 0).tolist() == (d[tag + "_X"] > 0).tolist()
    G = d[tag + "_A"].T @ d[tag + "_A"]
    R = d[tag + "_A"].T @ d[tag + "_B"]
    assert orc.kkt_violation(G, R, X) < 1e-8


def test_bpp_hard_case_uses_backup_rule(golden_bpp):
    from pyliger_b200 import nnlsm_blockpivot
    d = golden_bpp
    X, info = nnlsm_blockpivot(d["hard_A"], d["hard_B"])
    assert info[4] > 0 and int(d["hard_info"][3]) > 0  # both walked through the backup rule
    agree = np.mean(np.all((X > 0) == (d["hard_X"] > 0), axis=0))
    print("hard case: backup flips %d (reference %d), active-set agreement %.4f" % (info[4], d["hard_info"][3], agree))
    assert agree > 0.97


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_bpp_product_inputs(golden_bpp, dtype):
    from pyliger_b200 import nnlsm_blockpivot
    d = golden_bpp
    for tag in d["cases"]:
        tag = str(tag)
        X, info = nnlsm_blockpivot(d[tag + "_G"], d[tag + "_R"], is_input_prod=True, dtype=dtype)
        want = d[tag + "_X"]
        err = relerr(X, want)
        agree = np.mean(np.all((X > 0) == (want > 0), axis=0))
        print("%s %s: rel err %.2e, active-set agreement %.4f, backups %d" % (tag, dtype, err, agree, info[4]))
        assert info[0] is True
        if dtype == "float64":
            assert err < 1e-10 and agree == 1.0
            assert info[4] == int(d[tag + "_info"][3])
        else:
            assert err < 1e-4 and agree > 0.98


def test_bpp_sparse_B_and_warm_start(golden_bpp):
    from pyliger_b200 import nnlsm_blockpivot
    d = golden_bpp
    A, B = d["iid_A"], d["iid_B"]
    Xs, _ = nnlsm_blockpivot(A, sps.csr_matrix(B))
    np.testing.assert_allclose(Xs, d["iid_X"], rtol=1e-9, atol=1e-11)
    Xw, infow = nnlsm_blockpivot(A, B, init=d["iid_init"])
    np.testing.assert_allclose(Xw, d["iid_init_X"], rtol=1e-8, atol=1e-10)
    Xc, infoc = nnlsm_blockpivot(A, B)
    Xw2, infow2 = nnlsm_blockpivot(A, B, init=Xc)
    np.testing.assert_allclose(Xw2, Xc, rtol=1e-10, atol=1e-12)
    assert infow2[2] <= B.shape[1]  # a correct warm start needs exactly one factorisation per column
    assert infoc[2] > infow2[2]


@pytest.mark.parametrize("k", [1, 2, 20, 32, 33, 47, 64])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_bpp_random_vs_oracle_and_kkt(k, dtype):
    from pyliger_b200 import nnlsm_blockpivot
    rng = np.random.default_rng(100 + k)
    m, cols = 3 * k + 5, 777
    A = rng.uniform(0, 1, (m, k))
    B = rng.uniform(-0.3, 1.0, (m, cols))
    G, R = A.T @ A, A.T @ B
    X, info = nnlsm_blockpivot(G, R, is_input_prod=True, dtype=dtype)
    Xo, infoo = orc.nnlsm_blockpivot(G, R, is_input_prod=True)
    assert info[0] is True and X.shape == (k, cols)
    assert (X >= 0).all()
    if dtype == "float64":
        assert relerr(X, Xo) < 1e-9
        assert orc.kkt_violation(G, R, X) < 1e-9
        assert info[4] == infoo[4]
    else:
        assert relerr(X, Xo) < 5e-3
    for c in range(0, cols, 97):
        ref, _ = scipy.optimize.nnls(A, B[:, c])
        assert np.linalg.norm(X[:, c] - ref) <= (1e-8 if dtype == "float64" else 5e-3) * max(1.0, np.linalg.norm(ref))


@pytest.mark.parametrize("k", [33, 40, 47, 62, 63, 64])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("warm", [False, True])
def test_bpp_large_passive_sets(k, dtype, warm):
    """Mostly-positive solutions: passive sets of 32..k indices (complement-set / general solver paths)."""
    from pyliger_b200 import nnlsm_blockpivot
    rng = np.random.default_rng(300 + k)
    m, cols = 4 * k, 600
    A = rng.uniform(0, 1, (m, k)) + np.eye(m, k)
    Xt = rng.uniform(0.2, 1.0, (k, cols))
    Xt[rng.uniform(size=Xt.shape) < 0.08] = -0.5  # a few components want to be negative -> clamped
    R_noise = 0.05 * rng.standard_normal((m, cols))
    B = A @ Xt + R_noise
    G, R = A.T @ A, A.T @ B
    Xo, infoo = orc.nnlsm_blockpivot(G, R, is_input_prod=True)
    psize = (Xo > 0).sum(0)
    assert (psize > 31).any() and (psize < k).any()
    assert k < 40 or (psize > 31).mean() > 0.5
    init = None
    if warm:  # start from a perturbed passive set: one wrong index per column
        init = (Xo > 0).astype(float)
        flip = rng.integers(0, k, cols)
        init[flip, np.arange(cols)] = 1.0 - init[flip, np.arange(cols)]
    X, info = nnlsm_blockpivot(G, R, is_input_prod=True, init=init, dtype=dtype)
    assert info[0] is True and (X >= 0).all()
    if dtype == "float64":
        assert relerr(X, Xo) < 1e-9
        assert orc.kkt_violation(G, R, X) < 1e-8
        assert ((X > 0) == (Xo > 0)).all()
    else:
        assert relerr(X, Xo) < 5e-3


def test_bpp_edge_cases():
    from pyliger_b200 import nnlsm_blockpivot
    G = np.array([[2.0, 0.5], [0.5, 1.0]])
    X, info = nnlsm_blockpivot(G, np.zeros((2, 0)), is_input_prod=True)
    assert X.shape == (2, 0) and info[0] is True
    R = -np.abs(np.random.default_rng(0).uniform(0.1, 1, (2, 5)))  # all-negative rhs -> x = 0, y = -r
    X, info = nnlsm_blockpivot(G, R, is_input_prod=True)
    assert np.all(X == 0) and np.allclose(info[1], -R) and info[2] == 0
    R = G @ np.abs(R)  # interior solution -> equals the unconstrained one
    X, info = nnlsm_blockpivot(G, R, is_input_prod=True)
    np.testing.assert_allclose(X, np.linalg.solve(G, R), rtol=1e-12)
    with pytest.raises(ValueError):
        nnlsm_blockpivot(np.eye(65), np.ones((65, 2)), is_input_prod=True)


# ----------------------------------------------------------------------------------- optimize_ALS
def _check_als(lig, d, N, dtype):
    assert relerr(lig.adata_list[0].varm["W"], d["W"]) < FAC_TOL[dtype]
    for i in range(N):
        a = lig.adata_list[i]
        assert a.obsm["H"].dtype == np.float64 and a.obsm["H"].shape == d["H%d" % i].shape
        assert relerr(a.obsm["H"], d["H%d" % i]) < FAC_TOL[dtype]
        assert relerr(a.varm["V"], d["V%d" % i]) < FAC_TOL[dtype]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("warm", [False, True])
def test_als_trajectory_vs_reference(golden_als, dtype, warm):
    from pyliger_b200 import optimize_ALS
    d = golden_als
    k, lam, thresh, T, seed = d["params"]
    lig = _liger([csr_unpack(d, "X0"), csr_unpack(d, "X1")])
    info = optimize_ALS(lig, int(k), lam, thresh, int(T), rand_seed=int(seed), dtype=dtype, warm_start=warm,
                        return_info=True, progress=False)
    objs = np.array(info["objectives"][0])
    rel = np.abs(objs - d["objectives"]) / d["objectives"]
    print("ALS %s warm=%s: max rel objective error %.3e; bpp %s" % (dtype, warm, rel.max(), info["bpp"]["H"]))
    assert rel.max() < OBJ_TOL[dtype]
    _check_als(lig, d, 2, dtype)
    assert info["bpp"]["H"]["not_converged"] == 0 and info["bpp"]["H"]["pivot_failures"] == 0
    assert lig.W.shape == (csr_unpack(d, "X0").shape[1], int(k))


def test_als_returns_none_and_prints(golden_als3, capsys):
    from pyliger_b200 import optimize_ALS
    d = golden_als3
    k, lam, thresh, T, seed, nrep = d["params"]
    lig = _liger([csr_unpack(d, "X%d" % i) for i in range(3)])
    ret = optimize_ALS(lig, int(k), value_lambda=lam, thresh=thresh, max_iters=int(T), nrep=int(nrep),
                       rand_seed=int(seed), print_obj=True, progress=False)
    assert ret is None
    out = capsys.readouterr().out
    obj = float(out.strip().split("Objective:")[-1])
    assert abs(obj - float(d["objective"])) / float(d["objective"]) < 1e-9
    _check_als(lig, d, 3, "float64")


def test_als_value_errors(golden_als):
    from pyliger_b200 import optimize_ALS
    lig = _liger([csr_unpack(golden_als, "X0"), csr_unpack(golden_als, "X1")])
    with pytest.raises(ValueError, match="number of cells in smallest dataset"):
        optimize_ALS(lig, 260)
    with pytest.raises(ValueError, match="number of variable genes"):
        optimize_ALS(lig, 150)


def test_als_user_inits_and_no_early_work(golden_als):
    """W_init/V_init/H_init take the reference layouts (k x g, n x k); a huge thresh stops work after
    one iteration exactly like the reference's `continue` branch (_iNMF_ANLS.py:166-167)."""
    from pyliger_b200 import optimize_ALS
    d = golden_als
    mats = [csr_unpack(d, "X0"), csr_unpack(d, "X1")]
    k = 5
    rng = np.random.default_rng(4)
    W0 = rng.uniform(0, 2, (k, mats[0].shape[1]))
    V0 = [rng.uniform(0, 2, (k, mats[0].shape[1])) for _ in mats]
    H0 = [rng.uniform(0, 2, (m.shape[0], k)) for m in mats]
    want = orc.als(mats, k, 4.0, 1e-6, 3, W_init=W0, V_init=V0, H_init=H0)
    lig = _liger(mats)
    info = optimize_ALS(lig, k, 4.0, 1e-6, 3, W_init=W0, V_init=V0, H_init=H0, return_info=True, progress=False)
    np.testing.assert_allclose(info["objectives"][0], want["objectives"], rtol=1e-9)
    assert relerr(lig.H[0], want["H"][0]) < 1e-7
    stop = orc.als(mats, k, 4.0, 0.5, 6)
    lig = _liger(mats)
    info = optimize_ALS(lig, k, 4.0, 0.5, 6, return_info=True, progress=False)
    assert info["iters_run"] == [stop["iters_run"]] and 0 < stop["iters_run"] < 6
    # the pass 1 of the iteration that was not run (enqueued speculatively into the spare H buffer) is dropped
    for i in range(len(mats)):
        assert relerr(lig.H[i], stop["H"][i]) < 1e-7
    assert relerr(lig.W, stop["W"].T) < 1e-7


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_als_vs_oracle_medium(dtype):
    """Seeded medium problem (k=20, C1-like proportions) against the oracle incl. active sets."""
    from pyliger_b200 import optimize_ALS
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([900, 700], 400, 20, density=0.10, seed0=3000)
    k = 20
    want = orc.als(mats, k, 5.0, 1e-6, 6, rand_seed=1)
    lig = _liger(mats)
    info = optimize_ALS(lig, k, 5.0, 1e-6, 6, rand_seed=1, dtype=dtype, return_info=True, progress=False)
    rel = np.abs(np.array(info["objectives"][0]) - want["objectives"]) / np.array(want["objectives"])
    assert rel.max() < OBJ_TOL[dtype]
    for i in range(2):
        assert relerr(lig.H[i], want["H"][i]) < FAC_TOL[dtype]
        agree = np.mean(np.all((lig.H[i] > 0) == (want["H"][i] > 0), axis=1))
        print("ALS medium %s dataset %d: H active-set agreement per cell %.4f" % (dtype, i, agree))
        assert agree > (0.999 if dtype == "float64" else 0.9)
    assert relerr(lig.W.T, want["W"]) < FAC_TOL[dtype]


def test_als_large_properties():
    """Size-independent properties at a larger shape (fp32): monotone objective, KKT of the H solve,
    statistics-form objective == dense-residual objective of the returned factors."""
    from pyliger_b200 import optimize_ALS
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([6000, 5000, 4000], 1000, 30, density=0.08, structured=False, seed0=3100)
    k, lam = 30, 5.0
    lig = _liger(mats)
    info = optimize_ALS(lig, k, lam, 1e-6, 5, dtype="float32", return_info=True, progress=False)
    objs = np.array(info["objectives"][0])
    assert np.all(np.diff(objs) < 0)
    W = lig.W.T
    V = [v.T for v in lig.V]
    dense = orc.objective_dense(mats, lig.H, W, V, lam)
    assert abs(dense - objs[-1]) / dense < 1e-4
    assert all((h >= 0).all() for h in lig.H) and (W >= 0).all()
    assert info["bpp"]["H"]["not_converged"] == 0


# ----------------------------------------------------------------------------------- online_iNMF
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_online_scenario1_vs_reference(golden_online, dtype):
    from pyliger_b200 import online_iNMF
    d = golden_online
    k, lam, ep, mbi, mb, chunk, seed = d["s1_params"]
    lig = _liger([csr_unpack(d, "X0"), csr_unpack(d, "X1")])
    ret = online_iNMF(lig, k=int(k), value_lambda=lam, max_epochs=int(ep), miniBatch_max_iters=int(mbi),
                      miniBatch_size=int(mb), h5_chunk_size=int(chunk), rand_seed=int(seed), verbose=False,
                      dtype=dtype)
    assert ret is None
    tol = 1e-7 if dtype == "float64" else 2e-2
    assert relerr(lig.W, d["s1_W"]) < tol
    for i, a in enumerate(lig.adata_list):
        assert a.obsm["H"].shape == d["s1_H%d" % i].shape
        assert relerr(a.obsm["H"], d["s1_H%d" % i]) < tol
        assert relerr(a.varm["V"], d["s1_V%d" % i]) < tol
        assert relerr(a.varm["B"], d["s1_B%d" % i]) < tol
        assert relerr(a.uns["A"], d["s1_A%d" % i]) < tol


def test_online_two_inner_iters(golden_online):
    from pyliger_b200 import online_iNMF
    d = golden_online
    lig = _liger([csr_unpack(d, "X0")])
    online_iNMF(lig, k=4, value_lambda=3.0, max_epochs=2, miniBatch_max_iters=2, miniBatch_size=100,
                h5_chunk_size=64, rand_seed=3, verbose=False)
    assert relerr(lig.W, d["s1b_W"]) < 1e-7
    assert relerr(lig.H[0], d["s1b_H"]) < 1e-7
    assert relerr(lig.adata_list[0].uns["A"], d["s1b_A"]) < 1e-7
    assert relerr(lig.adata_list[0].varm["B"], d["s1b_B"]) < 1e-7


def _scenario1_state(lig, d):
    for i, a in enumerate(lig.adata_list):
        a.obsm["H"] = d["s1_H%d" % i].copy(); a.varm["W"] = d["s1_W"].copy()
        a.varm["V"] = d["s1_V%d" % i].copy(); a.varm["B"] = d["s1_B%d" % i].copy(); a.uns["A"] = d["s1_A%d" % i].copy()


def test_online_scenario2_refine(golden_online):
    from pyliger_b200 import AnnDataLike, online_iNMF
    d = golden_online
    lig = _liger([csr_unpack(d, "X0"), csr_unpack(d, "X1")])
    _scenario1_state(lig, d)
    new = AnnDataLike(csr_unpack(d, "Xnew"), "new")
    online_iNMF(lig, X_new=[new], k=5, value_lambda=5.0, max_epochs=2, miniBatch_size=80, h5_chunk_size=100,
                rand_seed=1, verbose=False)
    assert len(lig.adata_list) == int(d["s2_n_datasets"]) == 3
    assert relerr(new.varm["W"], d["s2_W"]) < 1e-7
    assert relerr(new.obsm["H"], d["s2_Hnew"]) < 1e-7
    assert relerr(new.varm["V"], d["s2_Vnew"]) < 1e-7
    assert relerr(new.uns["A"], d["s2_Anew"]) < 1e-7
    assert relerr(new.varm["B"], d["s2_Bnew"]) < 1e-7
    for i in range(2):
        assert relerr(lig.adata_list[i].obsm["H"], d["s2_Hold%d" % i]) < 1e-7
        assert relerr(lig.adata_list[i].varm["W"], d["s2_W"]) < 1e-7


def test_online_scenario3_projection(golden_online):
    from pyliger_b200 import AnnDataLike, online_iNMF
    d = golden_online
    lig = _liger([csr_unpack(d, "X0"), csr_unpack(d, "X1")])
    _scenario1_state(lig, d)
    new = AnnDataLike(csr_unpack(d, "Xnew"), "proj")
    online_iNMF(lig, X_new=[new], projection=True, k=5, miniBatch_size=100, verbose=False)
    assert relerr(new.obsm["H"], d["s3_H"]) < 1e-8
    assert np.all(new.varm["V"] == 0) and new.varm["V"].shape == d["s3_V"].shape
    assert len(lig.adata_list) == 3


def test_online_vs_oracle_medium_fp32():
    from pyliger_b200 import online_iNMF
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([3000, 2500], 500, 20, density=0.1, seed0=3200)
    want = orc.online(mats, 20, 5.0, 2, 1, 500, 1000, 1)
    lig = _liger(mats)
    online_iNMF(lig, k=20, value_lambda=5.0, max_epochs=2, miniBatch_size=500, verbose=False, dtype="float32")
    assert relerr(lig.W, want["W"]) < 2e-2
    assert relerr(lig.H[0], want["H"][0]) < 5e-2


# ----------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_als_ragged_inputs_with_empty_rows_and_columns(dtype):
    """Empty cells, genes that are all-zero in one dataset, very unequal dataset sizes, k = 1 and k = 33."""
    from pyliger_b200 import optimize_ALS
    rng = np.random.default_rng(9)
    g = 70
    mats = []
    for n, dens in ((400, 0.2), (37, 0.3), (1500, 0.05)):
        m = sps.random(n, g, density=dens, random_state=int(rng.integers(1 << 30)), format="lil",
                       data_rvs=lambda s: rng.gamma(2.0, 1.0, s))
        m[::7] = 0           # empty cells
        m[:, 5] = 0          # a gene that is all zero in this dataset
        m[:, g - 1] = 0
        mats.append(sps.csr_matrix(m))
    for k in (1, 33):
        want = orc.als(mats, k, 2.0, 1e-6, 4, rand_seed=3)
        lig = _liger(mats)
        info = optimize_ALS(lig, k, 2.0, 1e-6, 4, rand_seed=3, dtype=dtype, return_info=True, progress=False)
        rel = np.abs(np.array(info["objectives"][0]) - want["objectives"]) / np.array(want["objectives"])
        assert rel.max() < OBJ_TOL[dtype]
        for i in range(3):
            assert lig.H[i].shape == (mats[i].shape[0], k)
            assert relerr(lig.H[i], want["H"][i]) < (1e-6 if dtype == "float64" else 5e-2)
            assert np.all(lig.H[i][::7] == 0)  # empty cells stay at zero loading


def test_online_ragged_minibatches_wrap_many_epochs():
    """Mini-batch sizes that do not divide the dataset sizes, several epochs (window wrap-around)."""
    from pyliger_b200 import online_iNMF
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([333, 211], 90, 6, density=0.15, seed0=3300)
    want = orc.online(mats, 6, 4.0, 4, 1, 97, 50, 2)
    lig = _liger(mats)
    online_iNMF(lig, k=6, value_lambda=4.0, max_epochs=4, miniBatch_size=97, h5_chunk_size=50, rand_seed=2,
                verbose=False)
    assert relerr(lig.W, want["W"]) < 1e-7
    for i in range(2):
        assert relerr(lig.H[i], want["H"][i]) < 1e-6
        assert relerr(lig.adata_list[i].uns["A"], want["A"][i]) < 1e-7
        assert relerr(lig.adata_list[i].varm["B"], want["B"][i]) < 1e-7


def test_bpp_large_batch_properties():
    """Size-independent properties at a large column count (fp32): KKT conditions, idempotence of the warm
    start, linearity of the right-hand side scaling (x(c r) = c x(r))."""
    from pyliger_b200 import nnlsm_blockpivot
    rng = np.random.default_rng(12)
    k, cols = 40, 200000
    A = rng.uniform(0, 1, (120, k))
    G = A.T @ A
    R = G @ np.maximum(rng.normal(0.2, 1.0, (k, cols)), 0) + rng.normal(0, 0.5, (k, cols))
    X, info = nnlsm_blockpivot(G, R, is_input_prod=True, dtype="float32")
    assert info[0] is True and (X >= 0).all()
    Y = G @ X - R
    scale = np.abs(R).max()
    assert np.maximum(-Y, 0)[X == 0].max() / scale < 2e-4          # dual feasibility on the active set
    assert np.abs(Y[X > 0]).max() / scale < 2e-4                    # stationarity on the passive set
    X2, info2 = nnlsm_blockpivot(G, R, is_input_prod=True, dtype="float32", init=X)
    assert relerr(X2, X) < 1e-5 and info2[2] <= cols * 1.01         # already optimal: one solve per column
    X3, _ = nnlsm_blockpivot(G, 3.0 * R, is_input_prod=True, dtype="float32")
    assert relerr(X3, 3.0 * X) < 1e-4


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_als_with_nearly_dead_factor_falls_back_from_inverse(dtype):
    """A factor whose loadings are ~1e-3 of the others (G_55 ~ 1e-6 of the rest) makes G of the H-update nearly singular: the G^-1 based
    complement path must be refused (conditioning guard) and the Cholesky paths give the oracle's answer."""
    from pyliger_b200 import liger_from_matrices, optimize_ALS
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([700, 500], 260, 8, density=0.12, seed0=7100)
    k = 8
    rng = np.random.default_rng(3)
    W0 = rng.uniform(0, 2, (k, 260)); W0[5] *= 1e-3
    V0 = [rng.uniform(0, 2, (k, 260)) for _ in mats]
    for v in V0:
        v[5] *= 1e-3
    H0 = [rng.uniform(0, 2, (m.shape[0], k)) for m in mats]
    want = orc.als(mats, k, 5.0, 0.0, 2, W_init=W0.copy(), V_init=[v.copy() for v in V0],
                   H_init=[h.copy() for h in H0])
    lig = liger_from_matrices(mats)
    info = optimize_ALS(lig, k, 5.0, 0.0, 2, W_init=W0.copy(), V_init=[v.copy() for v in V0],
                        H_init=[h.copy() for h in H0], dtype=dtype, progress=False, return_info=True)
    tol = 1e-9 if dtype == "float64" else 1e-3
    got, ref = np.array(info["objectives"][0]), np.array(want["objectives"])
    assert np.max(np.abs(got - ref) / ref) < tol
    assert info["bpp"]["H"]["not_converged"] == 0


# ----------------------------------------------------------------------------------- BASELINE.json configs
def _trajectory_report(tag, lig, info, want, N):
    objs = np.array(info["objectives"][0])
    rel = np.abs(objs - np.array(want["objectives"])) / np.array(want["objectives"])
    herr = [relerr(lig.H[i], want["H"][i]) for i in range(N)]
    verr = [relerr(lig.V[i].T, want["V"][i]) for i in range(N)]
    werr = relerr(lig.W.T, want["W"])
    agree = [float(np.mean(np.all((lig.H[i] > 0) == (want["H"][i] > 0), axis=1))) for i in range(N)]
    print("%s: max rel objective err %.2e, H %s, W %.2e, V %s, H active-set agreement per cell %s, bpp(H) %s" % (
        tag, rel.max(), ["%.1e" % e for e in herr], werr, ["%.1e" % e for e in verr], ["%.4f" % a for a in agree],
        info["bpp"]["H"]))
    return rel.max(), max(herr), werr, max(verr), min(agree)


@pytest.fixture(scope="module")
def c1_problem():
    """configs[0] of BASELINE.json: 2 datasets x 7 000 cells x 2 000 genes, k = 20, lambda = 5, 30 iterations,
    rand_seed = 1 -- the oracle (the pinned restatement of the reference) runs it once for all C1 tests (~15 s)."""
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([7000, 7000], 2000, 20, density=0.10, seed0=1000)
    want = orc.als(mats, 20, 5.0, 1e-6, 30, rand_seed=1, dense_objective=False)
    return mats, want


@pytest.mark.parametrize("dtype,warm", [("float64", True), ("float64", False), ("float32", True), ("float32", False)])
def test_c1_full_trajectory_vs_oracle(c1_problem, dtype, warm):
    """C1 at full size: every one of the 30 objectives within the north star's tolerance (1e-5 fp64 / 1e-3 fp32;
    measured 1e-15 / 1e-8), final H, W, V in relative Frobenius norm, per-cell active-set agreement of H."""
    from pyliger_b200 import optimize_ALS
    mats, want = c1_problem
    lig = _liger(mats)
    info = optimize_ALS(lig, 20, 5.0, 1e-6, 30, rand_seed=1, dtype=dtype, warm_start=warm, return_info=True,
                        progress=False)
    assert info["iters_run"] == [want["iters_run"]]
    obj, h, w, v, agree = _trajectory_report("C1 %s warm=%s" % (dtype, warm), lig, info, want, 2)
    if dtype == "float64":
        assert obj < 1e-5 and obj < 1e-10    # north-star tolerance, and what the kernels actually deliver
        assert h < 1e-7 and w < 1e-7 and v < 1e-7 and agree > 0.999
    else:
        assert obj < 1e-3 and obj < 1e-5
        assert h < 2e-2 and w < 2e-2 and v < 2e-2 and agree > 0.98
    assert info["bpp"]["H"]["not_converged"] == 0


@pytest.mark.parametrize("name,ns,g,k,density", [
    ("C2 reduced", [2500, 2500, 2500, 2500], 600, 30, 0.08),          # 4 datasets, k = 30: sliced passes, half-warp BPP
    ("C4 reduced", [1300] * 8, 800, 40, 0.08),                         # 8 datasets, k = 40: split tile, complement BPP
])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_named_config_shapes_reduced_vs_oracle(name, ns, g, k, density, dtype):
    """The k = 30 / k = 40 configurations of BASELINE.json at a size the oracle finishes in seconds: 8 iterations of
    the whole trajectory against the oracle (same seeded init), default options."""
    from pyliger_b200 import optimize_ALS
    from pyliger_b200.synth import make_datasets
    mats = make_datasets(ns, g, k, density=density, seed0=3300)
    want = orc.als(mats, k, 5.0, 1e-6, 8, rand_seed=1, dense_objective=False)
    lig = _liger(mats)
    info = optimize_ALS(lig, k, 5.0, 1e-6, 8, rand_seed=1, dtype=dtype, return_info=True, progress=False)
    obj, h, w, v, agree = _trajectory_report("%s %s" % (name, dtype), lig, info, want, len(ns))
    if dtype == "float64":
        assert obj < 1e-9 and h < 1e-6 and w < 1e-6 and v < 1e-6 and agree > 0.999
    else:
        assert obj < 1e-3 and h < 5e-2 and w < 5e-2 and v < 5e-2 and agree > 0.95
    assert info["bpp"]["H"]["not_converged"] == 0


@pytest.mark.parametrize("k", [12, 30, 40, 64])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_half_warp_bpp_equals_full_warp_bpp(k, dtype):
    """inmf_bpp with the full workspace (half-warp solver + deferral list) against the same call with warm bit 3 (full-
    warp solver only): same X, same passive sets, same counters except `deferred`; every column is accounted for."""
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    dt = torch.float64 if dtype == "float64" else torch.float32
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(900 + k)
    g, cols = 300, 70011   # (batches below 64k columns go to the full-warp kernel alone)
    M = rng.uniform(0, 2, (k, g)) * (rng.random((k, g)) < 0.7)
    V = rng.uniform(0, 1, (k, g)) * (rng.random((k, g)) < 0.5)
    G = torch.from_numpy(M @ M.T + 5.0 * V @ V.T).to(dev)
    Xs = rng.gamma(0.3, 1.0, (cols, g)) * (rng.random((cols, g)) < 0.1)
    R = torch.from_numpy(Xs @ M.T - 0.5 * rng.random((cols, k))).to(dev, dt)   # mixed signs: partitions of all sizes
    seg = torch.tensor([0, cols // 3, cols], dtype=torch.int64, device=dev)     # two segments sharing G twice
    G2 = torch.stack([G, G * 1.5])
    res = {}
    for flags in (0, 8):
        for prefer in (0, 2):
            X = torch.empty_like(R)
            mask = torch.zeros(cols, dtype=torch.int64, device=dev)
            stats = torch.zeros(8, dtype=torch.int64, device=dev)
            ws = _lib.bpp_workspace(cols, 2, k, dev)
            for warm in (0, 1):  # cold, then warm from the cold run's sets
                stats.zero_()
                _lib.call("inmf_bpp", "f64" if dtype == "float64" else "f32", G2, R, k, X, k, None, 0, mask,
                          warm | prefer | flags, seg, 2, cols, k, stats, *_lib.ws_args(ws), _stream())
                res[(flags, prefer, warm)] = (X.clone(), mask.clone(), stats.cpu().numpy().copy())
    for prefer in (0, 2):
        for warm in (0, 1):
            a, b = res[(0, prefer, warm)], res[(8, prefer, warm)]
            assert b[2][7] == 0 and a[2][6] == cols and b[2][6] == cols and a[2][0] == 0 and b[2][0] == 0
            assert relerr(a[0].cpu().numpy(), b[0].cpu().numpy()) < (1e-12 if dtype == "float64" else 1e-4)
            same = float((a[1] == b[1]).double().mean().item())
            print("k=%d %s prefer=%d warm=%d: deferred %d of %d, identical passive sets %.5f, factorisations %d vs %d" % (
                k, dtype, prefer, warm, a[2][7], cols, same, a[2][1], b[2][1]))
            assert same > (0.9999 if dtype == "float64" else 0.99)
            if k <= 30 and prefer:
                assert a[2][7] == 0   # min(|P|, |F|) <= 15 always: nothing to defer


# ----------------------------------------------------------------------------------- iNMF_HALS
@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_inmf_hals_vs_reference(golden_hals, tag, dtype, capsys):
    """iNMF_HALS (pass 1 + per-cell HALS sweep + pass 2 + W/V HALS sweeps + statistics-form objective) against the
    executed reference: every printed objective, the written-back H, W, V; restarts (case b: nrep = 2)."""
    from pyliger_b200 import iNMF_HALS
    d = golden_hals
    N, k, lam, thresh, max_iters, nrep, seed = d[tag + "_params"]
    N, k, max_iters, nrep, seed = int(N), int(k), int(max_iters), int(nrep), int(seed)
    mats = [csr_unpack(d, "%s_X%d" % (tag, i)) for i in range(N)]
    lig = _liger(mats)
    info = iNMF_HALS(lig, k=k, value_lambda=lam, thresh=thresh, max_iters=max_iters, nrep=nrep, rand_seed=seed,
                     dtype=dtype, return_info=True)
    out = capsys.readouterr().out
    assert out.count("Initial Training Obj:") == nrep and "Iter: 1, Total time:" in out   # the reference's prints
    objs = np.concatenate(info["objectives"])
    rel = np.abs(objs - d[tag + "_objectives"]) / d[tag + "_objectives"]
    print("iNMF_HALS %s %s: max rel objective err %.2e" % (tag, dtype, rel.max()))
    assert objs.shape == d[tag + "_objectives"].shape and rel.max() < OBJ_TOL[dtype]
    for i in range(N):
        assert relerr(lig.adata_list[i].obsm["H"], d["%s_H%d" % (tag, i)]) < FAC_TOL[dtype]
        assert relerr(lig.adata_list[i].varm["V"], d["%s_V%d" % (tag, i)]) < FAC_TOL[dtype]
        assert relerr(lig.adata_list[i].varm["W"], d[tag + "_W"]) < FAC_TOL[dtype]


def test_c1_mixed_precision_mode(c1_problem):
    """dtype="mixed" (bf16 tile in the two passes, everything else fp32) on C1 at full size, 30 iterations: every
    objective within the north star's fp32 tolerance (1e-3); the factor errors and the per-cell active-set agreement
    of this reduced-precision mode are reported (and bounded loosely)."""
    from pyliger_b200 import optimize_ALS
    mats, want = c1_problem
    lig = _liger(mats)
    info = optimize_ALS(lig, 20, 5.0, 1e-6, 30, rand_seed=1, dtype="mixed", return_info=True, progress=False)
    obj, h, w, v, agree = _trajectory_report("C1 mixed", lig, info, want, 2)
    assert obj < 1e-3
    assert h < 0.2 and w < 0.2 and v < 0.2 and agree > 0.5
    assert info["bpp"]["H"]["not_converged"] == 0


@pytest.mark.parametrize("k,budget_rows", [(30, 100000), (40, 60), (12, 25)])
def test_deterministic_mode_is_bit_reproducible(k, budget_rows):
    """optimize_ALS(dtype='float32', deterministic=True): two runs give bitwise identical objectives and factors (the
    default mode accumulates with floating-point atomics: equal only to round-off), and both agree with the default
    mode to fp32 tolerance.  Small tiles force several gene / cell blocks (ordered sums over > 2 partials)."""
    from pyliger_b200 import blocked, optimize_ALS
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([2100, 1700, 900], 500, k, density=0.1, seed0=3500 + k)
    old = blocked.SMEM_BUDGET
    kp = max(32, (k + 3) // 4 * 4)
    blocked.SMEM_BUDGET = min(old, budget_rows * blocked.smem_row_bytes(kp, 4))
    try:
        runs = []
        for det in (True, True, False):
            lig = _liger(mats)
            info = optimize_ALS(lig, k, 5.0, 1e-6, 6, rand_seed=1, dtype="float32", deterministic=det, return_info=True,
                                progress=False)
            runs.append((np.array(info["objectives"][0]), [h.copy() for h in lig.H], lig.W.copy(),
                         [v.copy() for v in lig.V]))
    finally:
        blocked.SMEM_BUDGET = old
    a, b, c = runs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert all(np.array_equal(x, y) for x, y in zip(a[1], b[1])) and all(np.array_equal(x, y) for x, y in zip(a[3], b[3]))
    np.testing.assert_allclose(a[0], c[0], rtol=1e-5)
    assert relerr(a[2], c[2]) < 1e-2
    with pytest.raises(ValueError):
        optimize_ALS(_liger(mats), k, dtype="float64", deterministic=True, progress=False)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_graphed_iterations_equal_eager_iterations(dtype):
    """Warm-started iterations replayed as CUDA graphs (engine.use_graph) against the eager launch sequence: same
    objectives and factors (fp64: to round-off of the atomics; the kernels and their order are identical), early stop
    with a dropped speculative pass 1 included."""
    from pyliger_b200 import optimize_ALS
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([1500, 1100, 700], 300, 12, density=0.1, seed0=3700)
    res = {}
    old = InmfEngine.use_graph
    try:
        for graph in (True, False):
            InmfEngine.use_graph = graph
            for thresh, iters in ((1e-9, 9), (3e-3, 30)):
                lig = _liger(mats)
                info = optimize_ALS(lig, 12, 5.0, thresh, iters, rand_seed=1, dtype=dtype, return_info=True, progress=False)
                res[(graph, thresh)] = (np.array(info["objectives"][0]), [h.copy() for h in lig.H], lig.W.copy())
    finally:
        InmfEngine.use_graph = old
    tol = 1e-12 if dtype == "float64" else 1e-5
    for thresh in (1e-9, 3e-3):
        a, b = res[(True, thresh)], res[(False, thresh)]
        assert a[0].shape == b[0].shape and (thresh == 1e-9 or 3 < len(a[0]) < 30)
        np.testing.assert_allclose(a[0], b[0], rtol=tol)
        assert relerr(a[2], b[2]) < (1e-10 if dtype == "float64" else 1e-3)
        assert all(relerr(x, y) < (1e-10 if dtype == "float64" else 1e-3) for x, y in zip(a[1], b[1]))


@pytest.mark.parametrize("dtype,k,det", [("float64", 12, False), ("float32", 12, False), ("float32", 36, False),
                                         ("float32", 50, False), ("mixed", 30, False), ("float32", 30, True)])
def test_kernels_write_inside_their_buffers(monkeypatch, dtype, k, det):
    """Every buffer the engine hands to a kernel is bracketed by sentinel bands (InmfEngine.guard_elems): after
    optimize_ALS, iNMF_HALS and online_iNMF with small tiles (several gene / cell blocks, ragged last slices) no band
    has been written.  (Stands in for compute-sanitizer memcheck, which is not available on the GPU pool.)"""
    from pyliger_b200 import blocked, iNMF_HALS, online_iNMF, optimize_ALS
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.synth import make_datasets
    engines = []
    orig = InmfEngine._alloc

    def recording_alloc(self):
        engines.append(self)
        return orig(self)
    monkeypatch.setattr(InmfEngine, "_alloc", recording_alloc)
    monkeypatch.setattr(InmfEngine, "guard_elems", 256)
    kp = max(32, (k + 3) // 4 * 4)
    monkeypatch.setattr(blocked, "SMEM_BUDGET", min(blocked.SMEM_BUDGET, 200 * blocked.smem_row_bytes(kp, 4)))
    mats = make_datasets([1333, 1071, 517], 411, k, density=0.12, seed0=3900 + k)
    optimize_ALS(_liger(mats), k, 5.0, 1e-6, 5, rand_seed=1, dtype=dtype, deterministic=det, progress=False)
    if dtype != "mixed" and not det:
        iNMF_HALS(_liger(mats), k, max_iters=4, rand_seed=1, dtype=dtype, verbose=False)
        online_iNMF(_liger(mats), k=k, miniBatch_size=600, max_epochs=2, rand_seed=1, dtype=dtype, verbose=False)
    assert engines and all(e._guards for e in engines)
    for e in engines:
        assert e.check_guards() == []
    # the check itself sees a stray write
    e = engines[0]
    name, raw, ge, n = e._guards[0]
    raw[ge + n] = 0
    assert e.check_guards() == [name]


@pytest.mark.parametrize("k,ns,mb,epochs", [(6, [333, 211], 97, 4), (36, [2100, 1300, 900], 1000, 3),
                                            (20, [1200, 1200], 600, 2), (40, [2000, 9100], 5500, 2)])
def test_online_window_statistics_equal_row_statistics(k, ns, mb, epochs):
    """fp32 online_iNMF: the mini-batch statistics X_mb' H_mb, H_mb' H_mb through the block-aligned sliced pass
    (InmfEngine.window_stats) against the per-row scatter kernel -- ragged windows that straddle two blocks, wrap around
    the end of a dataset over several epochs, an aligned case (one block per window) and windows larger than a tile."""
    from pyliger_b200 import online_iNMF
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.synth import make_datasets
    mats = make_datasets(ns, 300, k, density=0.1, seed0=4100 + k)
    res = {}
    old = InmfEngine.window_stats
    used = []
    orig = InmfEngine.stats_window

    def counting(self, step):
        used.append(step)
        return orig(self, step)
    try:
        InmfEngine.stats_window = counting
        for flag in (True, False):
            InmfEngine.window_stats = flag
            lig = _liger(mats)
            online_iNMF(lig, k=k, value_lambda=5.0, max_epochs=epochs, miniBatch_size=mb, rand_seed=3, dtype="float32",
                        verbose=False)
            res[flag] = (lig.W.copy(), [v.copy() for v in lig.V], [h.copy() for h in lig.H],
                         [a.uns["A"].copy() for a in lig.adata_list], [a.varm["B"].copy() for a in lig.adata_list])
            if flag:
                assert used, "the window pass was not taken"
                n_used = len(used)
        assert len(used) == n_used
    finally:
        InmfEngine.window_stats = old
        InmfEngine.stats_window = orig
    a, b = res[True], res[False]
    assert relerr(a[0], b[0]) < 2e-4
    for i in range(len(ns)):
        assert relerr(a[1][i], b[1][i]) < 2e-3
        assert relerr(a[3][i], b[3][i]) < 1e-4
        assert relerr(a[4][i], b[4][i]) < 1e-4
        assert relerr(a[2][i], b[2][i]) < 5e-3
