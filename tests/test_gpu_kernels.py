"""GPU: every C-ABI kernel against numpy / the oracle on seeded inputs (run with -m gpu on a B200)."""
import numpy as np
import pytest
import scipy.sparse as sps
import torch

from conftest import relerr
from oracle import inmf_oracle as orc

pytestmark = pytest.mark.gpu

DTYPES = [("float64", 1e-11), ("float32", 2e-5)]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _engine(mats, k, lam, dtype):
    from pyliger_b200.engine import InmfEngine
    return InmfEngine(mats, k, lam, dtype=dtype)


def _rand_problem(seed, ns=(230, 170), g=90, k=7, density=0.15):
    rng = np.random.default_rng(seed)
    mats = [sps.random(n, g, density=density, random_state=seed + i, format="csr",
                       data_rvs=lambda s: rng.gamma(2.0, 1.0, s)) for i, n in enumerate(ns)]
    W = rng.uniform(0, 2, (k, g))
    V = [rng.uniform(0, 2, (k, g)) * (rng.random((k, g)) < 0.7) for _ in ns]
    H = [rng.uniform(0, 2, (n, k)) * (rng.random((n, k)) < 0.6) for n in ns]
    return mats, W, V, H


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("k", [7, 33, 64])
def test_prep_spmm_stats_objective(dev, dtype, tol, k):
    lam = 3.5
    mats, W, V, H = _rand_problem(k, k=k)
    eng = _engine(mats, k, lam, dtype)
    eng.set_als_state(W, V, H)
    eng.prep()
    N, g = len(mats), mats[0].shape[1]
    for i in range(N):
        M = W + V[i]
        assert relerr(eng.Mt[i, :, :k].cpu().numpy(), M.T) < tol
        assert float(eng.Mt[i, :, k:].abs().max().item() if eng.ldm > k else 0.0) == 0.0
        assert relerr(eng.Gwv[i].cpu().numpy(), M @ M.T) < tol
        assert relerr(eng.Gvv[i].cpu().numpy(), V[i] @ V[i].T) < tol
        assert relerr(eng.G[i].cpu().numpy(), M @ M.T + lam * V[i] @ V[i].T) < tol
    # statistics with the H we set
    eng.stats()
    for i in range(N):
        assert relerr(eng.St[i].cpu().numpy(), np.asarray(mats[i].T.dot(H[i]))) < tol
        assert relerr(eng.HtH[i].cpu().numpy(), H[i].T @ H[i]) < tol
    obj = float(eng.objective().item())
    want = orc.objective_dense(mats, H, W, V, lam)
    assert abs(obj - want) / want < max(tol, 1e-12) * 10
    xn = eng.xnorm2.cpu().numpy()
    np.testing.assert_allclose(xn, [m.multiply(m).sum() for m in mats], rtol=max(tol, 1e-12))
    # SpMM alone: R = X (W+V)'
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    R = torch.zeros_like(eng.H)
    _lib.call("inmf_spmm", eng.sfx, eng.X.rowptr, eng.X.colidx, eng.X.vals, eng.Mt, eng.ldm, eng.g, R, eng.ldm,
              eng.seg_full, N, eng.max_rows_full, k, _stream())
    want_R = np.vstack([np.asarray(mats[i].dot((W + V[i]).T)) for i in range(N)])
    assert relerr(R[:, :k].cpu().numpy(), want_R) < tol
    # blocked pass 1 (TMA tile + packed pairs) computes the same right-hand sides
    eng.rhs_H()
    assert eng.pass1 is not None and eng.pass2 is not None
    assert relerr(eng.H[:, :k].cpu().numpy(), want_R) < tol
    assert float(eng.H[:, k:].abs().max().item() if eng.ldm > k else 0.0) == 0.0
    # and the row-kernel statistics agree with the blocked pass 2 used by eng.stats() above
    eng.set_als_state(W, V, H)
    St2 = torch.zeros_like(eng.St); HtH2 = torch.zeros_like(eng.HtH)
    eng.stats(seg_desc=eng.seg_full, H=eng.H, St=St2, HtH=HtH2)
    assert relerr(St2.cpu().numpy(), eng.St.cpu().numpy()) < tol
    assert relerr(HtH2.cpu().numpy(), eng.HtH.cpu().numpy()) < tol


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_rhs_v_w(dev, dtype, tol):
    k, lam = 9, 5.0
    mats, W, V, H = _rand_problem(3, k=k)
    eng = _engine(mats, k, lam, dtype)
    eng.set_als_state(W, V, H)
    eng.stats()
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    N, g = eng.N, eng.g
    _lib.call("inmf_rhs_v", eng.sfx, eng.St, eng.Wt, eng.HtH, eng.Rv, eng.Gv, lam, N, g, k, _stream())
    _lib.call("inmf_rhs_w", eng.sfx, eng.St, eng.Vt, eng.HtH, eng.Rw, eng.Gw, N, g, k, _stream())
    S = [np.asarray(mats[i].T.dot(H[i])).T for i in range(N)]
    C = [H[i].T @ H[i] for i in range(N)]
    for i in range(N):
        assert relerr(eng.Rv[i].cpu().numpy(), (S[i] - C[i] @ W).T) < tol * 10
        assert relerr(eng.Gv[i].cpu().numpy(), (1 + lam) * C[i]) < tol
    assert relerr(eng.Rw.cpu().numpy(), sum(S[i] - C[i] @ V[i] for i in range(N)).T) < tol * 10
    assert relerr(eng.Gw[0].cpu().numpy(), sum(C)) < tol


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_window_segments_wrap(dev, dtype, tol):
    """Mini-batch windows incl. wrap-around at the end of a dataset (seg_desc semantics)."""
    k, lam = 6, 2.0
    mats, W, V, H = _rand_problem(5, ns=(50, 37), g=40, k=k)
    eng = _engine(mats, k, lam, dtype)
    eng.set_als_state(W, V, None)
    eng.prep()
    # dataset 0: rows 44..49,0..5 (12 rows); dataset 1: rows 30..36,0..1 (9 rows)
    seg = torch.tensor([[0, 0, 44, 50], [12, 50, 30, 37], [21, 0, 0, 0]], dtype=torch.int64, device=dev)
    out = torch.zeros(21, eng.ldm, dtype=eng.dtype, device=dev)
    eng.solve_H(seg_desc=seg, max_rows=12, out=out)
    rows0 = [(44 + j) % 50 for j in range(12)]
    rows1 = [(30 + j) % 37 for j in range(9)]
    for i, rows, sl in ((0, rows0, slice(0, 12)), (1, rows1, slice(12, 21))):
        M = W + V[i]
        G = M @ M.T + lam * V[i] @ V[i].T
        R = np.asarray(mats[i][rows].dot(M.T)).T
        want = orc.bpp(G, R)[0].T
        assert relerr(out[sl, :k].cpu().numpy(), want) < max(tol * 50, 1e-9)
    St = torch.zeros_like(eng.St)
    HtH = torch.zeros_like(eng.HtH)
    eng.stats(seg_desc=seg, max_rows=12, H=out, St=St, HtH=HtH)
    Hh = out[:, :k].cpu().numpy().astype(np.float64)
    assert relerr(St[0].cpu().numpy(), np.asarray(mats[0][rows0].T.dot(Hh[:12]))) < tol * 10
    assert relerr(St[1].cpu().numpy(), np.asarray(mats[1][rows1].T.dot(Hh[12:]))) < tol * 10
    assert relerr(HtH[1].cpu().numpy(), Hh[12:].T @ Hh[12:]) < tol * 10


def test_update_ab_kat6(dev, golden_online):
    """KAT-6 of SURVEY.md App. C through the CUDA epilogue (fp64)."""
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    d = golden_online
    f64 = torch.float64
    A = torch.eye(2, dtype=f64, device=dev).reshape(1, 2, 2).clone()
    B = torch.ones(1, 3, 2, dtype=f64, device=dev)
    Ao = torch.zeros_like(A)
    Bo = torch.zeros_like(B)
    Hm = np.array([[1.0, 2], [0, 1]])
    Xm = np.arange(6, dtype=np.float64).reshape(3, 2)
    HHt = torch.from_numpy(Hm @ Hm.T).to(dev).reshape(1, 2, 2)
    XHt = torch.from_numpy(Xm @ Hm.T).to(dev).reshape(1, 3, 2)
    inv_mb = torch.tensor([0.5], dtype=f64, device=dev)
    for s, (it, ep, epp) in enumerate(((0, 0, 0), (1, 0, 0), (2, 1, 0), (3, 1, 1))):
        scale = 0.0 if it == 0 else (0.5 if it == 1 else (it - 1) / it)
        roll = int(ep > 0 and ep - epp == 1)
        _lib.call("inmf_update_ab", "f64", A, Ao, B, Bo, HHt, XHt, inv_mb, scale, roll, 1, 3, 2, _stream())
        np.testing.assert_allclose(A[0].cpu().numpy(), d["kat6_A"][s], rtol=1e-14)
        np.testing.assert_allclose(B[0].cpu().numpy(), d["kat6_B"][s], rtol=1e-14)
        np.testing.assert_allclose(Ao[0].cpu().numpy(), d["kat6_Aold"][s], rtol=1e-14)
        np.testing.assert_allclose(Bo[0].cpu().numpy(), d["kat6_Bold"][s], rtol=1e-14)
    # exact-zero diagonal -> 1e-15
    A.zero_(); Ao.zero_()
    _lib.call("inmf_update_ab", "f64", A, Ao, B, Bo, torch.zeros_like(HHt), XHt, inv_mb, 0.0, 0, 1, 3, 2, _stream())
    np.testing.assert_array_equal(A[0].cpu().numpy(), [[1e-15, 0], [0, 1e-15]])


@pytest.mark.parametrize("N,k,n_prev,n_inner", [(1, 5, 0, 1), (1, 40, 0, 2), (2, 20, 0, 1), (3, 33, 1, 1), (4, 30, 0, 1),
                                                (5, 64, 2, 2), (8, 40, 0, 1), (8, 12, 3, 1), (9, 40, 0, 1), (13, 17, 4, 1),
                                                (20, 24, 0, 1), (8, 64, 0, 1), (2, 64, 0, 1)])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-4)])
def test_hals_sweeps_all_kernel_shapes_vs_oracle(dev, N, k, n_prev, n_inner, dtype, tol):
    """inmf_hals picks its kernel by the number of current datasets and factors (lane groups per dataset for <= 8
    current datasets, column-per-lane forms beyond); every form against the oracle's _update_W_HALS / _update_V_HALS
    (_utilities.py:102-129), with previous datasets (scenario 2) and repeated inner iterations."""
    from oracle import inmf_oracle as orc
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream, resolve_dtype
    rng = np.random.RandomState(100 * N + k)
    g = 301
    nall = N
    Hs = [rng.rand(3 * k + 7, k) * (rng.rand(3 * k + 7, k) < 0.6) for _ in range(nall)]
    A = np.stack([h.T @ h / len(h) for h in Hs])
    for i in range(nall):
        A[i][np.diag_indices(k)] += 1e-3
    B = rng.rand(nall, g, k)
    W0 = rng.rand(g, k)
    V0 = rng.rand(nall, g, k) * (rng.rand(nall, g, k) < 0.7)
    W, V = W0.copy(), [v.copy() for v in V0]
    for _ in range(n_inner):  # _online_iNMF.py:444-460: W over all datasets, V over the current ones
        W = orc.hals_W(list(A), list(B), W, V)
        V[n_prev:] = orc.hals_V(list(A[n_prev:]), list(B[n_prev:]), W, V[n_prev:], 5.0)
    dt = resolve_dtype(dtype)
    sfx = "f64" if dt == torch.float64 else "f32"
    Ad = torch.from_numpy(A).to(dev)
    Bd = torch.from_numpy(B).to(dev, dt)
    Wd = torch.from_numpy(W0).to(dev, dt)
    Vd = torch.from_numpy(V0).to(dev, dt)
    _lib.call("inmf_hals", sfx, Ad, Bd, Wd, Vd, 5.0, n_prev, nall, g, k, n_inner, _stream())
    assert relerr(Wd.cpu().numpy(), W) < tol
    assert relerr(Vd.cpu().numpy(), np.stack(V)) < tol
    if n_prev:
        np.testing.assert_array_equal(Vd[:n_prev].cpu().numpy(), torch.from_numpy(V0[:n_prev]).to(dt).numpy())


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-11), ("float32", 5e-5)])
def test_hals_sweep_golden(dev, golden_online, dtype, tol):
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream, resolve_dtype
    d = golden_online
    dt = resolve_dtype(dtype)
    A = torch.from_numpy(d["hals_A"]).to(dev)
    B = torch.from_numpy(d["hals_B"]).to(dev, dt)
    W = torch.from_numpy(d["hals_W0"]).to(dev, dt)
    V = torch.from_numpy(d["hals_V0"]).to(dev, dt)
    N, g, k = d["hals_V0"].shape
    sfx = "f64" if dt == torch.float64 else "f32"
    _lib.call("inmf_hals", sfx, A, B, W, V, 5.0, 0, N, g, k, 1, _stream())
    assert relerr(W.cpu().numpy(), d["hals_W1"]) < tol
    assert relerr(V.cpu().numpy(), d["hals_V1"]) < tol
    # scenario-2 form: first dataset is "previous" -> its V must stay untouched, W sums still use it
    W2 = torch.from_numpy(d["hals_W0"]).to(dev, dt)
    V2 = torch.from_numpy(d["hals_V0"]).to(dev, dt)
    _lib.call("inmf_hals", sfx, A, B, W2, V2, 5.0, 1, N, g, k, 1, _stream())
    assert relerr(W2.cpu().numpy(), d["hals_W1"]) < tol
    np.testing.assert_array_equal(V2[0].cpu().numpy(), torch.from_numpy(d["hals_V0"][0]).to(dt).numpy())
    assert relerr(V2[1:].cpu().numpy(), d["hals_V1"][1:]) < tol


def test_error_reporting(dev):
    from pyliger_b200 import _lib
    with pytest.raises(_lib.InmfError, match="k must be"):
        _lib.call("inmf_bpp", "f32", 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 1, 1, 65, 0, 0, 0, 0)


def test_tcgen05_selftest_tile(dev):
    """One 128 x 64 x n tcgen05 (kind::tf32) tile with the accumulator in TMEM: TF32 and 3xTF32 accuracy."""
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in (16, 32, 48, 64):
        A = rng.uniform(-1, 1, (128, 64)).astype(np.float32)
        B = rng.uniform(-1, 1, (64, n)).astype(np.float32)
        want = A.astype(np.float64) @ B.astype(np.float64)
        for three, tol in ((0, 3e-3), (1, 1e-5)):
            C = torch.full((128, n), float("nan"), device=dev)
            Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
            rc = lib.inmf_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), n, three, _stream())
            torch.cuda.synchronize()
            assert rc == 0
            assert np.abs(C.cpu().numpy() - want).max() / np.abs(want).max() < tol


@pytest.mark.parametrize("mode,tol", [("tf32x3", 5e-5), ("tf32", 5e-3)])
def test_tcgen05_passes_match_cuda_core_passes(dev, mode, tol):
    """Dense-tile tcgen05 form of both passes against the CUDA-core blocked form (ragged sizes, k = 9, 33)."""
    from pyliger_b200.engine import InmfEngine
    for k in (9, 33):
        mats, W, V, H = _rand_problem(11 + k, ns=(700, 333, 130), g=210, k=k, density=0.12)
        ref = InmfEngine(mats, k, 5.0, dtype=torch.float32)
        tc = InmfEngine(mats, k, 5.0, dtype=torch.float32, use_tc=True, tc_three_pass=(mode == "tf32x3"))
        assert tc.tc1 is not None and tc.tc2 is not None
        for e in (ref, tc):
            e.set_als_state(W, V, H)
            e.prep()
            e.stats()
        assert relerr(tc.St.cpu().numpy(), ref.St.cpu().numpy()) < tol
        assert relerr(tc.HtH.cpu().numpy(), ref.HtH.cpu().numpy()) < 1e-6
        ref.rhs_H()
        tc.rhs_H()
        assert relerr(tc.H.cpu().numpy(), ref.H.cpu().numpy()) < tol


def test_als_with_tensor_core_passes(dev):
    from pyliger_b200 import liger_from_matrices, optimize_ALS
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([900, 700], 400, 20, density=0.10, seed0=3000)
    want = orc.als(mats, 20, 5.0, 1e-6, 5, rand_seed=1)
    lig = liger_from_matrices(mats)
    info = optimize_ALS(lig, 20, 5.0, 1e-6, 5, rand_seed=1, dtype="float32", tensor_core="tf32x3", return_info=True,
                        progress=False)
    rel = np.abs(np.array(info["objectives"][0]) - want["objectives"]) / np.array(want["objectives"])
    assert rel.max() < 1e-3
    assert relerr(lig.H[0], want["H"][0]) < 1e-2


@pytest.mark.parametrize("seed,n", [(0, 1), (0, 311), (1, 312), (12345, 313), (4294967295, 1000003)])
def test_device_mt19937_stream_is_numpys(dev, seed, n):
    """inmf_mt19937_uniform == np.random.seed(seed); np.random.uniform(0, 2, n), bit for bit."""
    from pyliger_b200 import _lib
    out = torch.empty(n, dtype=torch.float64, device=dev)
    _lib.call("inmf_mt19937_uniform", "", seed, n, 2.0, out, torch.cuda.current_stream().cuda_stream)
    np.random.seed(seed)
    ref = np.random.uniform(0, 2, n)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("world,rank,replicate", [(1, 0, False), (3, 1, False), (2, 1, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_device_format_builders_equal_sort_path(dev, dtype, world, rank, replicate):
    """inmf_blockfmt1/2 (count / scan / fill) produce exactly the arrays of the generic stable-sort path,
    including rows that are empty and a matrix with duplicate (unsummed) column entries."""
    from pyliger_b200 import blocked
    from pyliger_b200.engine import DeviceCSR
    rng = np.random.default_rng(17)
    g, kp = 523, 12
    mats = [sps.random(n, g, density=0.12, random_state=40 + i, format="csr") for i, n in enumerate((2011, 640, 3))]
    dup = sps.random(300, g, density=0.1, random_state=7, format="coo")
    dup = sps.csr_matrix((np.r_[dup.data, dup.data[:500]], (np.r_[dup.row, dup.row[:500]], np.r_[dup.col, dup.col[:500]])),
                         shape=dup.shape)
    dup_raw = sps.csr_matrix((dup.data, dup.indices, dup.indptr), shape=dup.shape)
    mats.append(dup_raw)
    mats[0] = sps.csr_matrix(mats[0].multiply(rng.uniform(size=(2011, 1)) > 0.1))  # some empty cells
    X = DeviceCSR(mats, dtype, dev, rank, world, replicate=replicate)
    esize = 4 if dtype == torch.float32 else 8
    old = blocked.SMEM_BUDGET
    blocked.SMEM_BUDGET = 200 * kp * esize  # several gene / cell blocks
    try:
        out = {}
        for flag in (True, False):
            blocked.USE_DEVICE_BUILD = flag
            out[flag] = (blocked.build_pass1(X, kp, esize, 148), blocked.build_pass2(X, kp, esize, 148))
    finally:
        blocked.SMEM_BUDGET = old
        blocked.USE_DEVICE_BUILD = True
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a.ptr, b.ptr)
        assert a.pairs.shape == b.pairs.shape and torch.equal(a.pairs, b.pairs)
        assert torch.equal(a.work, b.work) and a.max_tile_rows == b.max_tile_rows


@pytest.mark.parametrize("k", [33, 40, 47, 56, 64])
def test_split_tile_passes_match_plain_passes(dev, k):
    """inmf_blocked_split_f32 ([rows x 32] + remainder tile) against numpy and against the plain tiled kernel:
    pass 1 (X (W+V)'), pass 2 (X'H and H'H), several gene / cell blocks, and one whole ALS run."""
    from pyliger_b200 import blocked
    lam = 2.0
    mats, W, V, H = _rand_problem(k, k=k)
    old = (blocked.SMEM_BUDGET, blocked.SPLIT_TILE, blocked.SELL)
    blocked.SMEM_BUDGET = 40 * 256  # 40-64 tile rows: several gene and cell blocks at this toy size
    blocked.SELL = False            # the row-ordered kernels themselves (the sliced ones: test_sell_passes_...)
    res = {}
    try:
        for split in (False, True):
            blocked.SPLIT_TILE = split
            eng = _engine(mats, k, lam, "float32")
            assert eng.pass1.split == split and eng.pass2.split == split
            eng.set_als_state(W, V, H)
            eng.prep()
            eng.stats()
            St, HtH = eng.St.clone(), eng.HtH.clone()
            eng.rhs_H()
            R = eng.H.clone()
            eng.set_als_state(W, V, H)
            objs = [float(eng.initial_objective())] + [float(eng.als_iteration(warm=i > 0).item()) for i in range(3)]
            res[split] = (St, HtH, R, objs, eng.H.clone())
    finally:
        blocked.SMEM_BUDGET, blocked.SPLIT_TILE, blocked.SELL = old
    N = len(mats)
    St, HtH, R, objs, Hf = res[True]
    off = 0
    for i in range(N):
        n = mats[i].shape[0]
        assert relerr(St[i].cpu().numpy(), np.asarray(mats[i].T.dot(H[i]))) < 5e-6
        assert relerr(HtH[i].cpu().numpy(), H[i].T @ H[i]) < 5e-6
        assert relerr(R[off:off + n, :k].cpu().numpy(), np.asarray(mats[i].dot((W + V[i]).T))) < 5e-6
        off += n
    assert float(R[:, k:].abs().max().item() if R.shape[1] > k else 0.0) == 0.0  # pad columns stay zero
    np.testing.assert_allclose(objs, res[False][3], rtol=2e-6)
    assert relerr(Hf.cpu().numpy(), res[False][4].cpu().numpy()) < 1e-3


@pytest.mark.parametrize("k", [29, 30, 32])
def test_wide_tile_passes_match_plain_passes(dev, k):
    """inmf_blocked_wide_f32 (KP = 32: 8 factors per lane, bank-swizzled halves) against numpy and the plain kernel."""
    from pyliger_b200 import blocked
    lam = 2.0
    mats, W, V, H = _rand_problem(100 + k, ns=(333, 170), g=120, k=k, density=0.3)
    old = (blocked.SMEM_BUDGET, blocked.WIDE_TILE, blocked.SELL)
    blocked.SMEM_BUDGET = 50 * 128  # 50 tile rows: several gene and cell blocks
    blocked.SELL = False
    res = {}
    try:
        for wide in (False, True):
            blocked.WIDE_TILE = wide
            eng = _engine(mats, k, lam, "float32")
            assert eng.ldm == 32
            eng.set_als_state(W, V, H)
            eng.prep()
            eng.stats()
            St, HtH = eng.St.clone(), eng.HtH.clone()
            eng.rhs_H()
            R = eng.H.clone()
            eng.set_als_state(W, V, H)
            objs = [float(eng.initial_objective())] + [float(eng.als_iteration(warm=i > 0).item()) for i in range(3)]
            res[wide] = (St, HtH, R, objs, eng.H.clone())
    finally:
        blocked.SMEM_BUDGET, blocked.WIDE_TILE, blocked.SELL = old
    St, HtH, R, objs, Hf = res[True]
    off = 0
    for i in range(len(mats)):
        n = mats[i].shape[0]
        assert relerr(St[i].cpu().numpy(), np.asarray(mats[i].T.dot(H[i]))) < 5e-6
        assert relerr(HtH[i].cpu().numpy(), H[i].T @ H[i]) < 5e-6
        assert relerr(R[off:off + n, :k].cpu().numpy(), np.asarray(mats[i].dot((W + V[i]).T))) < 5e-6
        off += n
    assert float(R[:, k:].abs().max().item() if R.shape[1] > k else 0.0) == 0.0
    np.testing.assert_allclose(objs, res[False][3], rtol=2e-6)
    assert relerr(Hf.cpu().numpy(), res[False][4].cpu().numpy()) < 1e-3


@pytest.mark.parametrize("k,height", [(3, 32), (20, 32), (30, 32), (32, 32), (20, 8), (30, 8), (32, 8), (33, 8), (40, 8),
                                      (33, 33), (36, 33), (37, 33), (40, 33), (44, 8), (48, 8), (50, 8), (64, 8)])
@pytest.mark.parametrize("budget_rows", [37, 100000])
def test_sell_passes_match_numpy_and_row_ordered_passes(dev, k, height, budget_rows):
    """inmf_sell_pass_f32 (rows sorted by length, slices of 8 rows, 4 lanes per row) against numpy and against the
    row-ordered tiled kernels: pass 1, pass 2 + H'H, with several gene / cell blocks and with one, ragged row lengths
    (empty rows and columns included), and a short ALS run."""
    from pyliger_b200 import blocked
    lam = 2.0
    mats, W, V, H = _rand_problem(200 + k, ns=(333, 171, 70), g=130, k=k, density=0.25)  # every n_i > k
    mats[0] = sps.csr_matrix(mats[0].multiply(np.arange(mats[0].shape[0])[:, None] % 7 != 0))  # empty rows
    mats[1] = sps.csr_matrix(mats[1].multiply(np.arange(130)[None, :] % 5 != 0))               # empty columns
    kp = max(32, (k + 3) // 4 * 4)
    old = (blocked.SMEM_BUDGET, blocked.SELL, blocked.SELL_HEIGHT_KP32, blocked.SELL_HEIGHT_KP40)
    blocked.SMEM_BUDGET = min(old[0], budget_rows * blocked.smem_row_bytes(kp, 4))
    blocked.SELL_HEIGHT_KP32 = height
    blocked.SELL_HEIGHT_KP40 = height
    res = {}
    try:
        for sell in (False, True):
            blocked.SELL = sell
            eng = _engine(mats, k, lam, "float32")
            assert isinstance(eng.pass1, blocked.SellPass) == sell and isinstance(eng.pass2, blocked.SellPass) == sell
            if sell:
                assert eng.pass1.height == height and eng.pass2.height == height
                assert eng.pass1.padded >= eng.pass1.nnz and eng.pass2.padded >= eng.pass2.nnz
                assert eng.pass1.nnz == eng.pass2.nnz == sum(m.nnz for m in mats)
            eng.set_als_state(W, V, H)
            eng.prep()
            eng.stats()
            St, HtH = eng.St.clone(), eng.HtH.clone()
            eng.rhs_H()
            R = eng.H.clone()
            eng.set_als_state(W, V, H)
            objs = [float(eng.initial_objective())] + [float(eng.als_iteration(warm=i > 0).item()) for i in range(3)]
            res[sell] = (St, HtH, R, objs, eng.H.clone())
    finally:
        blocked.SMEM_BUDGET, blocked.SELL, blocked.SELL_HEIGHT_KP32, blocked.SELL_HEIGHT_KP40 = old
    St, HtH, R, objs, Hf = res[True]
    off = 0
    for i in range(len(mats)):
        n = mats[i].shape[0]
        assert relerr(St[i].cpu().numpy(), np.asarray(mats[i].T.dot(H[i]))) < 5e-6
        assert relerr(HtH[i].cpu().numpy(), H[i].T @ H[i]) < 5e-6
        assert relerr(R[off:off + n, :k].cpu().numpy(), np.asarray(mats[i].dot((W + V[i]).T))) < 5e-6
        off += n
    assert float(R[:, k:].abs().max().item() if R.shape[1] > k else 0.0) == 0.0  # pad columns stay zero
    # different summation order than the row-ordered kernels: the fp32 objective of this toy problem (4e7 -> 6e4,
    # i.e. three digits of cancellation) agrees to the north star's fp32 tolerance
    np.testing.assert_allclose(objs, res[False][3], rtol=1e-3)
    # (no comparison of the final H of the two fp32 runs: on this random toy problem the ALS map amplifies summation-
    # order differences; trajectories are pinned against the oracle in test_gpu_parity.py)


def _bf16_round(a):
    """float32 -> bf16 (round to nearest even) -> float32, in numpy."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("k", [3, 20, 30, 32])
@pytest.mark.parametrize("budget_rows", [50, 100000])
def test_mixed_precision_passes_equal_numpy_with_bf16_operand(dev, k, budget_rows):
    """dtype="mixed": the two passes with the dense operand rounded to bf16 (and nothing else): R = X bf16(W+V)',
    S = X' bf16(H) to fp32 round-off, H'H from the unrounded H; parity-ordered slices, several tiles and one."""
    from pyliger_b200 import blocked
    lam = 2.0
    mats, W, V, H = _rand_problem(300 + k, ns=(333, 171, 70), g=130, k=k, density=0.25)
    old = blocked.SMEM_BUDGET
    blocked.SMEM_BUDGET = min(old, budget_rows * 64)
    try:
        eng = _engine(mats, k, lam, "mixed")
        assert eng.mixed and eng.pass1.height == -32 and eng.pass2.height == -32 and eng.dtype == torch.float32
        eng.set_als_state(W, V, H)
        eng.prep()
        eng.stats()
        St, HtH = eng.St.clone(), eng.HtH.clone()
        eng.rhs_H()
        R = eng.H.clone()
    finally:
        blocked.SMEM_BUDGET = old
    off = 0
    for i in range(len(mats)):
        n = mats[i].shape[0]
        Mi = _bf16_round((W + V[i]).astype(np.float32)).astype(np.float64)
        Hi32 = H[i].astype(np.float32)
        assert relerr(R[off:off + n, :k].cpu().numpy(), np.asarray(mats[i].dot(Mi.T))) < 5e-6
        assert relerr(St[i].cpu().numpy(), np.asarray(mats[i].T.dot(_bf16_round(Hi32).astype(np.float64)))) < 5e-6
        assert relerr(HtH[i].cpu().numpy(), Hi32.astype(np.float64).T @ Hi32.astype(np.float64)) < 5e-6
        # and the rounding is the only difference to the exact products: 2^-9 per operand entry
        assert relerr(R[off:off + n, :k].cpu().numpy(), np.asarray(mats[i].dot((W + V[i]).T))) < 4e-3
        off += n
    assert float(R[:, k:].abs().max().item() if R.shape[1] > k else 0.0) == 0.0
    with pytest.raises(ValueError):
        _engine(mats, 33, lam, "mixed")
