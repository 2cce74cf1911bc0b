"""CPU: host-side logic, the C-ABI surface (load + exported symbols, no compute), 2-rank gloo plumbing."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sps
import torch

from conftest import ROOT, csr_unpack
from oracle import inmf_oracle as orc


def test_library_exports_every_declared_symbol():
    from pyliger_b200 import _lib
    from pyliger_b200.build import build_library
    build_library()  # no-op when up to date; nvcc cross-compiles without a GPU
    header = open(os.path.join(ROOT, "include", "inmf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(inmf_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 24
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.load().inmf_version() >= 100


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pyliger_b200
    from pyliger_b200._lib import InmfError
    with pytest.raises(InmfError, match="no CPU fallback"):
        pyliger_b200.nnlsm_blockpivot(np.eye(2), np.ones((2, 1)), is_input_prod=True)
    lig = pyliger_b200.liger_from_matrices([sps.random(30, 10, 0.3, format="csr") for _ in range(2)])
    with pytest.raises(InmfError):
        pyliger_b200.optimize_ALS(lig, 3)
    with pytest.raises(ValueError):  # argument checks come first, as in the reference
        pyliger_b200.optimize_ALS(lig, 10)


def test_product_path_never_imports_oracle():
    """Only tests/ (incl. tests/tools/), __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for top in ("pyliger_b200", "benchmarks"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".sh", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in src, "%s/%s mentions the oracle" % (top, f)


def test_shard_bounds_cover():
    from pyliger_b200.engine import shard_bounds
    for n in (0, 1, 7, 100, 250001):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_device_csr_slicing_cpu(golden_als):
    from pyliger_b200.engine import DeviceCSR
    mats = [csr_unpack(golden_als, "X0"), csr_unpack(golden_als, "X1")]
    world = 3
    parts = [DeviceCSR(mats, torch.float64, torch.device("cpu"), r, world) for r in range(world)]
    for i, m in enumerate(mats):
        rows = []
        for p in parts:
            b0, b1 = p.row_base[i], p.row_base[i + 1]
            ip = p.rowptr.numpy()
            sub = sps.csr_matrix((p.vals.numpy()[ip[b0]:ip[b1]], p.colidx.numpy()[ip[b0]:ip[b1]],
                                  ip[b0:b1 + 1] - ip[b0]), shape=(b1 - b0, m.shape[1]))
            rows.append(sub)
        assert abs(sps.vstack(rows) - m).max() == 0
    rep = DeviceCSR(mats, torch.float32, torch.device("cpu"), 1, world, replicate=True)
    assert rep.n_total == sum(m.shape[0] for m in mats) and rep.own == parts[1].own
    assert rep.vals.dtype == torch.float32


def test_online_windows_match_reference_row_order():
    from pyliger_b200.engine import online_windows, shard_bounds
    n, mb, epochs, chunk = [333, 200], [130, 78], 2, 100
    iters = 5
    for world in (1, 2):
        covered = [[[] for _ in range(iters)] for _ in n]
        for rank in range(world):
            sub = [shard_bounds(m, rank, world) for m in mb]
            plan = online_windows(iters, mb, sub, n, [0, n[0]])
            for t in range(iters):
                for i in range(2):
                    cnt = plan[t, i + 1, 0] - plan[t, i, 0]
                    assert cnt == sub[i][1] - sub[i][0]
                    covered[i][t] += [int((plan[t, i, 2] + j) % plan[t, i, 3]) for j in range(cnt)]
        for i in range(2):
            want = orc.minibatch_rows(iters, mb[i], epochs, chunk, n[i])
            for t in range(iters):
                rows = [r for lo, hi in want[t] for r in range(lo, hi)]
                assert covered[i][t] == rows


def test_als_init_stream_matches_kat3():
    from pyliger_b200.factorization._iNMF_ANLS import _als_init_host
    W, V, H = _als_init_host([4, 5], 3, 2, 1 + 0 - 1)  # rand_seed=1 seeds 0 (App. B-1)
    np.testing.assert_allclose(W, [[1.097627007855, 1.430378732745, 1.205526752143],
                                   [1.089766365994, 0.847309598678, 1.291788226133]], rtol=1e-12)
    Wo, Vo, Ho = orc.als_init([4, 5], 3, 2, 1)
    np.testing.assert_array_equal(H[1], Ho[1])
    np.testing.assert_array_equal(V[0], Vo[0])


def test_online_init_matches_oracle(golden_online):
    from pyliger_b200.factorization._online_iNMF import _init_V_online, _init_W
    np.testing.assert_array_equal(_init_W(3, 2, 1), golden_online["kat4_initW"])
    X0 = csr_unpack(golden_online, "X0")
    np.testing.assert_allclose(_init_V_online(420, 5, X0, 100, 1), golden_online["s1_initV0"], rtol=1e-13)


def test_pack_passive_mask():
    from pyliger_b200.factorization._utilities import pack_passive_mask
    init = np.zeros((64, 3))
    init[0, 0] = 1; init[63, 0] = 2; init[5, 1] = 0.1; init[6, 1] = -1
    m = pack_passive_mask(init).view(np.uint64)
    assert m[0] == (1 | (1 << 63)) and m[1] == (1 << 5) and m[2] == 0


# ------------------------------------------------------------------------------------------ gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from pyliger_b200.engine import DeviceCSR, all_gather_rows, shard_bounds
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        mats = [sps.random(n, 23, density=0.3, random_state=7 + i, format="csr") for i, n in enumerate((41, 30, 5))]
        k = 4
        H = [rng.uniform(0, 1, (m.shape[0], k)) for m in mats]
        part = DeviceCSR(mats, torch.float64, torch.device("cpu"), rank, world)
        ip = part.rowptr.numpy()
        St = torch.zeros(len(mats), 23, k, dtype=torch.float64)
        HtH = torch.zeros(len(mats), k, k, dtype=torch.float64)
        gathered = []
        for i, m in enumerate(mats):
            b0, b1 = part.row_base[i], part.row_base[i + 1]
            lo, hi = part.own[i]
            sub = sps.csr_matrix((part.vals.numpy()[ip[b0]:ip[b1]], part.colidx.numpy()[ip[b0]:ip[b1]],
                                  ip[b0:b1 + 1] - ip[b0]), shape=(b1 - b0, 23))
            Hl = H[i][lo:hi]
            St[i] = torch.from_numpy(np.asarray(sub.T.dot(Hl)))
            HtH[i] = torch.from_numpy(Hl.T @ Hl)
            counts = [shard_bounds(m.shape[0], r, world)[1] - shard_bounds(m.shape[0], r, world)[0] for r in range(world)]
            gathered.append(all_gather_rows(torch.from_numpy(Hl.copy()), counts).numpy())
        dist.all_reduce(St)
        dist.all_reduce(HtH)
        ok = True
        for i, m in enumerate(mats):
            ok &= np.allclose(St[i].numpy(), np.asarray(m.T.dot(H[i])), rtol=1e-12)
            ok &= np.allclose(HtH[i].numpy(), H[i].T @ H[i], rtol=1e-12)
            ok &= np.array_equal(gathered[i], H[i])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_statistics_and_gather():
    """world_size 2 over gloo: row-sharded statistics all-reduce to the full ones; ragged H gather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------------------------------ blocked formats
def _emulate_blocked(bp, dense, kp, out, k, gram=None, row_bytes=None):
    """numpy statement of what inmf_blocked_spmm (or, with row_bytes = 128, inmf_blocked_split_f32) computes from
    (work, ptr, pairs)."""
    work = bp.work.numpy()[:bp.nwork]
    ptr = bp.ptr.numpy()
    pairs = bp.pairs.numpy()
    esize = 4 if pairs.dtype == np.int32 else 8
    row_bytes = kp * esize if row_bytes is None else row_bytes
    assert np.all(pairs[:, 0] % row_bytes == 0)  # premultiplied byte offsets of tile rows
    idx = pairs[:, 0].astype(np.int64) // row_bytes
    vals = pairs[:, 1].copy().view(np.float32 if pairs.dtype == np.int32 else np.float64).astype(np.float64)
    for ptr_base, dense_row0, tile_rows, r0, r1, out_base, gram_seg, _ in work:
        tile = dense[dense_row0:dense_row0 + tile_rows]
        assert tile.shape[0] == tile_rows
        for r in range(r0, r1):
            b, e = ptr[ptr_base + r], ptr[ptr_base + r + 1]
            if e > b:
                assert idx[b:e].max() < tile_rows and idx[b:e].min() >= 0
                out[out_base + r, :k] += vals[b:e] @ tile[idx[b:e], :k]
        if gram is not None and gram_seg >= 0:
            gram[gram_seg] += tile[:, :k].T @ tile[:, :k]


@pytest.mark.parametrize("world,rank", [(1, 0), (3, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_blocked_formats_reproduce_products(world, rank, dtype):
    from pyliger_b200 import blocked
    from pyliger_b200.engine import DeviceCSR
    rng = np.random.default_rng(5)
    g, k, kp = 157, 6, 8
    mats = [sps.random(n, g, density=0.2, random_state=3 + i, format="csr") for i, n in enumerate((301, 77, 5))]
    X = DeviceCSR(mats, dtype, torch.device("cpu"), rank, world)
    esize = 4 if dtype == torch.float32 else 8
    old = blocked.SMEM_BUDGET
    blocked.SMEM_BUDGET = 50 * kp * esize  # force several gene / cell blocks at this toy size
    try:
        p1 = blocked.build_pass1(X, kp, esize, n_sms=7)
        p2 = blocked.build_pass2(X, kp, esize, n_sms=7)
    finally:
        blocked.SMEM_BUDGET = old
    assert p1.max_tile_rows <= 50 and p2.max_tile_rows <= 50
    N = len(mats)
    Mt = rng.uniform(0, 1, (N * g, kp)); Mt[:, k:] = 0
    H = rng.uniform(0, 1, (X.n_own_total, kp)); H[:, k:] = 0
    R = np.zeros((X.n_own_total, kp))
    _emulate_blocked(p1, Mt, kp, R, k)
    St = np.zeros((N * g, k)); gram = np.zeros((N, k, k))
    _emulate_blocked(p2, H, kp, St, k, gram)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    for s, m in enumerate(mats):
        lo, hi = X.own[s]
        o0, o1 = X.own_off[s], X.own_off[s + 1]
        sub = m[lo:hi]
        np.testing.assert_allclose(R[o0:o1, :k], sub @ Mt[s * g:(s + 1) * g, :k], rtol=tol, atol=tol)
        np.testing.assert_allclose(St[s * g:(s + 1) * g], sub.T @ H[o0:o1, :k], rtol=tol, atol=tol)
        np.testing.assert_allclose(gram[s], H[o0:o1, :k].T @ H[o0:o1, :k], rtol=1e-12, atol=1e-12)
    # every owned non-zero appears exactly once in each pass
    own_nnz = sum(m[lo:hi].nnz for m, (lo, hi) in zip(mats, X.own))
    assert p1.pairs.shape[0] == own_nnz == p2.pairs.shape[0]


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1)])
def test_window_work_lists_reproduce_minibatch_rhs(world, rank):
    """Blocked format over all resident cells + per-step work lists == X[window rows] @ M of online_iNMF."""
    from pyliger_b200 import blocked
    from pyliger_b200.engine import DeviceCSR, online_windows, shard_bounds
    rng = np.random.default_rng(11)
    g, k, kp = 93, 5, 8
    mats = [sps.random(n, g, density=0.25, random_state=20 + i, format="csr") for i, n in enumerate((131, 58))]
    X = DeviceCSR(mats, torch.float64, torch.device("cpu"), rank, world, replicate=True)
    mbs = [40, 18]
    sub = [shard_bounds(m, rank, world) for m in mbs]  # rank r handles [lo, hi) of every mini-batch
    steps = 9  # > one epoch of the first dataset: windows wrap
    plan = online_windows(steps, mbs, sub, X.n_global, X.row_base)
    old = blocked.SMEM_BUDGET
    blocked.SMEM_BUDGET = 40 * kp * 8
    try:
        bp, NB, gbs = blocked.build_pass1_resident(X, kp, 8)
    finally:
        blocked.SMEM_BUDGET = old
    assert NB > 1
    work, counts = blocked.window_work(X, NB, gbs, plan, pieces=3)
    Mt = rng.uniform(0, 1, (len(mats) * g, kp)); Mt[:, k:] = 0
    wrapped = 0
    for t in range(steps):
        bp.work, bp.nwork = torch.from_numpy(work[t]), int(counts[t])
        rows = sum(hi - lo for lo, hi in sub)
        out = np.zeros((rows, kp))
        _emulate_blocked(bp, Mt, kp, out, k)
        for s, m in enumerate(mats):
            out_off, win_start, n = int(plan[t, s, 0]), int(plan[t, s, 2]), int(plan[t, s, 3])
            cnt = int(plan[t, s + 1, 0]) - out_off
            idx = (win_start + np.arange(cnt)) % n
            wrapped += int(cnt > 0 and idx[-1] < idx[0])
            np.testing.assert_allclose(out[out_off:out_off + cnt, :k], m[idx] @ Mt[s * g:(s + 1) * g, :k],
                                       rtol=1e-12, atol=1e-12)
    assert wrapped > 0


def test_split_tile_formats_use_128_byte_rows():
    """With the split-tile kernel selected (fp32, 32 < KP <= 64) the formats carry row * 128 offsets and the tile
    row capacity follows the [rows x 32] + [rows x rs] shared-memory footprint."""
    from pyliger_b200 import blocked
    from pyliger_b200.engine import DeviceCSR
    rng = np.random.default_rng(8)
    g, k, kp = 211, 37, 40
    mats = [sps.random(n, g, density=0.15, random_state=60 + i, format="csr") for i, n in enumerate((190, 44))]
    X = DeviceCSR(mats, torch.float32, torch.device("cpu"))
    old = (blocked.SMEM_BUDGET, blocked.SPLIT_TILE)
    blocked.SMEM_BUDGET, blocked.SPLIT_TILE = 60 * 160, True
    try:
        assert blocked.split_applies(kp, 4) and not blocked.split_applies(kp, 8) and not blocked.split_applies(32, 4)
        assert blocked.row_bytes(kp, 4) == 128 and blocked.smem_row_bytes(kp, 4) == 160
        assert blocked.smem_row_bytes(48, 4) == 192 and blocked.smem_row_bytes(64, 4) == 256
        p1 = blocked.build_pass1(X, kp, 4, n_sms=5)
        p2 = blocked.build_pass2(X, kp, 4, n_sms=5)
    finally:
        blocked.SMEM_BUDGET, blocked.SPLIT_TILE = old
    assert p1.split and p2.split and p1.max_tile_rows <= 60 and p2.max_tile_rows <= 60
    Mt = rng.uniform(0, 1, (2 * g, kp)); Mt[:, k:] = 0
    H = rng.uniform(0, 1, (X.n_own_total, kp)); H[:, k:] = 0
    R = np.zeros((X.n_own_total, kp)); St = np.zeros((2 * g, k)); gram = np.zeros((2, k, k))
    _emulate_blocked(p1, Mt, kp, R, k, row_bytes=128)
    _emulate_blocked(p2, H, kp, St, k, gram, row_bytes=128)
    for s_, m in enumerate(mats):
        o0, o1 = X.own_off[s_], X.own_off[s_ + 1]
        np.testing.assert_allclose(R[o0:o1, :k], m @ Mt[s_ * g:(s_ + 1) * g, :k], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(St[s_ * g:(s_ + 1) * g], m.T @ H[o0:o1, :k], rtol=1e-5, atol=1e-5)


def _emulate_tc(tp, dense, k, out):
    """numpy statement of inmf_tc_relayout + inmf_tc_pass from (work, ptr, pairs, segd)."""
    from pyliger_b200 import blocked
    work = tp.work.numpy()[:tp.nwork]
    ptr = tp.ptr.numpy()
    pairs = tp.pairs.numpy()
    off = pairs[:, 0].astype(np.int64)
    vals = pairs[:, 1].copy().view(np.float32).astype(np.float64)
    # invert the A-tile offset: (m, kk)
    m = (off // 2048) * 8 + (off % 128) // 16
    kk = ((off % 2048) // 128) * 4 + (off % 16) // 4
    assert np.all(blocked._tc_a_offset(m, kk) == off)
    segd = tp.segd.numpy()
    chunk_rows = {}
    for row0, rows, chunk0 in segd:
        for c in range(-(-rows // 64)):
            chunk_rows[chunk0 + c] = (row0 + c * 64, min(64, rows - c * 64))
    for pidx0, dchunk0, nch, out_row0, tile_rows, atomic, _, _ in work:
        for c in range(nch):
            b, e = ptr[(pidx0 + c) * 8], ptr[(pidx0 + c) * 8 + 8]
            r0, nr = chunk_rows[dchunk0 + c]
            if e > b:
                assert m[b:e].max() < tile_rows and kk[b:e].max() < nr
                bands = np.repeat(np.arange(8), np.diff(ptr[(pidx0 + c) * 8:(pidx0 + c) * 8 + 9]))
                assert np.all(m[b:e] // 16 == bands)
                np.add.at(out, out_row0 + m[b:e], vals[b:e, None] * dense[r0 + kk[b:e], :k])


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1)])
def test_tc_formats_reproduce_products(world, rank):
    from pyliger_b200 import blocked
    from pyliger_b200.engine import DeviceCSR
    rng = np.random.default_rng(6)
    g, k = 300, 6
    mats = [sps.random(n, g, density=0.1, random_state=13 + i, format="csr") for i, n in enumerate((401, 130, 64))]
    X = DeviceCSR(mats, torch.float32, torch.device("cpu"), rank, world)
    p1 = blocked.build_tc_pass1(X, k)
    p2 = blocked.build_tc_pass2(X, k, n_sms=4)
    assert p1.npad == 16 and p2.npad == 16
    N = len(mats)
    Mt = rng.uniform(0, 1, (N * g, 8))
    H = rng.uniform(0, 1, (X.n_own_total, 8))
    R = np.zeros((X.n_own_total, k))
    _emulate_tc(p1, Mt, k, R)
    St = np.zeros((N * g, k))
    _emulate_tc(p2, H, k, St)
    for s, mm in enumerate(mats):
        lo, hi = X.own[s]
        o0, o1 = X.own_off[s], X.own_off[s + 1]
        sub = mm[lo:hi]
        np.testing.assert_allclose(R[o0:o1], sub @ Mt[s * g:(s + 1) * g, :k], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(St[s * g:(s + 1) * g], sub.T @ H[o0:o1, :k], rtol=1e-5, atol=1e-5)


def test_device_twin_cache_detects_edits_and_dies_with_the_matrix():
    import gc
    from pyliger_b200 import devcache
    m = sps.random(50, 30, density=0.3, random_state=1, format="csr")
    dev = torch.device("cpu")
    t = (torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int32)),
         torch.from_numpy(m.data.copy()))
    devcache.register(m, dev, *t)
    assert devcache.lookup(m, dev) is not None
    assert devcache.lookup(m, torch.device("meta")) is None       # other device
    assert devcache.lookup(m.copy(), dev) is None                 # identity, not equality
    # registered matrices are frozen: an in-place edit (even of a single entry no sampled fingerprint would see)
    # raises instead of silently diverging from the device copy
    with pytest.raises(ValueError):
        m.data[1] = 123.0
    with pytest.raises(ValueError):
        m.data *= 2.0
    assert devcache.lookup(m, dev) is not None
    # whoever re-enables writing gets the upload path: the twin is dropped on the next lookup
    m.data.flags.writeable = True
    m.data[1] = 123.0                                             # one (non-sampled) entry
    assert devcache.lookup(m, dev) is None and devcache.lookup(m, dev) is None
    big = sps.random(400, 300, density=0.3, random_state=3, format="csr")   # > 4096 non-zeros: strided sampling
    devcache.register(big, dev, *t)
    big.data.flags.writeable = True
    big.data[1] += 1.0
    assert devcache.lookup(big, dev) is None
    m2 = sps.random(20, 30, density=0.3, random_state=2, format="csr")
    devcache.register(m2, dev, *t)
    n_before = len(devcache._cache)
    del m2
    gc.collect()
    assert len(devcache._cache) == n_before - 1


def test_anndata_like_slicing_and_raw():
    from pyliger_b200.containers import AnnDataLike, liger_from_raw
    X = sps.random(12, 7, density=0.5, random_state=0, format="csr")
    a = AnnDataLike.from_raw(X, "s", var_names=["g%d" % j for j in range(7)])
    a.layers["norm_data"] = X * 2.0
    a.var["norm_sum_sq"] = np.arange(7.0)
    a.obsm["H"] = np.arange(24.0).reshape(12, 2)
    keep = np.array([True, False, True, True, False, False, True])
    a.raw = a
    b = a[:, keep].copy()
    assert b.shape == (12, 4) and b.layers["norm_data"].shape == (12, 4)
    assert list(b.var.index) == ["g0", "g2", "g3", "g6"] and list(b.var["norm_sum_sq"]) == [0.0, 2.0, 3.0, 6.0]
    assert b.raw.shape == (12, 7) and b.obsm["H"].shape == (12, 2)
    c = a[np.arange(12) % 2 == 0, :]
    assert c.shape == (6, 7) and c.obsm["H"].shape == (6, 2) and len(c.obs) == 6
    lig = liger_from_raw([X, X[:5]], ["a", "b"])
    assert lig.num_samples == 2 and lig.adata_list[1].shape == (5, 7) and lig.adata_list[0].layers == {}


def _window_work_loop(X, NB, gbs, plan, pieces):
    """Straightforward statement of blocked.window_work (one Python iteration per work item)."""
    g, N = X.g, X.N
    out = []
    for t in range(plan.shape[0]):
        items = []
        for s in range(N):
            out_off, win_start, n = int(plan[t, s, 0]), int(plan[t, s, 2]), int(plan[t, s, 3])
            count = int(plan[t, s + 1, 0]) - out_off
            per = max(-(-count // pieces), 1)
            for b in range(NB):
                rows_b = min(gbs, g - b * gbs)
                pbase = NB * int(X.row_base[s]) + b * X.n_local[s]
                for q0 in range(0, count, per):
                    q1 = min(count, q0 + per)
                    lo = (win_start + q0) % n
                    hi = lo + (q1 - q0)
                    for a, e, obase in ((lo, min(hi, n), out_off + q0 - lo), (0, max(hi - n, 0), out_off + q0 + (n - lo))):
                        if e > a:
                            items.append([pbase, s * g + b * gbs, rows_b, a, e, obase, -1, 0])
        out.append(items)
    return out


@pytest.mark.parametrize("pieces", [1, 3, 4])
def test_window_work_vectorised_equals_loop(pieces):
    from pyliger_b200 import blocked
    from pyliger_b200.engine import DeviceCSR, online_windows
    mats = [sps.random(n, 61, density=0.2, random_state=5 + i, format="csr") for i, n in enumerate((131, 58, 17))]
    X = DeviceCSR(mats, torch.float64, torch.device("cpu"), replicate=True)
    mbs = [40, 18, 5]
    for sub in ([(0, 40), (0, 18), (0, 5)], [(13, 27), (6, 12), (2, 4)], [(0, 40), (18, 18), (0, 5)]):
        plan = online_windows(11, mbs, sub, X.n_global, X.row_base)
        work, counts = blocked.window_work(X, 3, 21, plan, pieces)
        ref = _window_work_loop(X, 3, 21, plan, pieces)
        for t in range(plan.shape[0]):
            assert counts[t] == len(ref[t])
            assert work[t, :counts[t]].tolist() == ref[t]
            assert (work[t, counts[t]:, 3] == work[t, counts[t]:, 4]).all()  # padding items are empty


def test_csr_store_roundtrip_append_and_slicing(tmp_path):
    """pyliger_b200.io.CSRStore: the h5sparse-like surface of the backed mode (create_dataset / append / row slicing),
    zero-copy raw_rows, fancy row selection, read-only protection."""
    from pyliger_b200.io import CSRStore, h5_idx_generator, write_store
    m = sps.random(537, 41, density=0.2, random_state=4, format="csr")
    p = str(tmp_path / "s.csr")
    with CSRStore(p, "w") as f:
        for lo, hi in h5_idx_generator(100, m.shape[0]):
            if "scale_data" not in f.keys():
                f.create_dataset("scale_data", data=m[lo:hi], chunks=(100,), maxshape=(None,))
            else:
                f["scale_data"].append(m[lo:hi])
    assert list(h5_idx_generator(100, 250)) == [(0, 100), (100, 200), (200, 250)]
    f = CSRStore(p, "r")
    ds = f["scale_data"]
    assert ds.shape == m.shape and ds.nnz == m.nnz and f.keys() == ["scale_data"] and "scale_data" in f
    assert (ds[:] != m).nnz == 0 and (ds[100:333] != m[100:333]).nnz == 0 and ds[530:9999].shape == (7, 41)
    ip, ix, dv = ds.raw_rows(17, 29)
    sub = m[17:29]
    assert np.array_equal(ip, sub.indptr) and np.array_equal(ix, sub.indices) and np.array_equal(dv, sub.data)
    idx = np.array([5, 5, 400, 2])
    assert (ds[idx, :] != m[idx, :]).nnz == 0
    with pytest.raises(IOError):
        ds.append(m[:3])
    with pytest.raises(KeyError):
        f["norm_data"]
    with pytest.raises(IOError):
        CSRStore(str(tmp_path / "missing.csr"), "r")
    # float32 storage and adding a second dataset to an existing store
    write_store(p, "norm_data", m, chunk_size=64, dtype="float32", mode="a")
    g = CSRStore(p, "r")
    assert g.keys() == ["norm_data", "scale_data"] and g["norm_data"].dtype == np.float32
    np.testing.assert_allclose(g["norm_data"][:].toarray(), m.toarray(), rtol=1e-6)


def test_streaming_window_ranges_wrap_repeatedly():
    from pyliger_b200.streaming import _window_ranges
    assert _window_ranges(3, 4, 10) == [(3, 7)]
    assert _window_ranges(8, 5, 10) == [(8, 10), (0, 3)]
    assert _window_ranges(27, 25, 10) == [(7, 10), (0, 10), (0, 10), (0, 2)]   # mini-batch larger than the dataset
    assert _window_ranges(0, 0, 10) == []


def test_backed_anndata_like(tmp_path):
    from pyliger_b200.containers import AnnDataLike
    from pyliger_b200.io import write_store
    m = sps.random(30, 12, density=0.3, random_state=1, format="csr")
    p = write_store(str(tmp_path / "a.csr"), "raw_data", m, chunk_size=7, mode="w")
    a = AnnDataLike.backed(p, "a")
    assert a.isbacked and a.shape == (30, 12) and a.X is None and "backed_path" in a.uns_keys()
    b = a[:, np.arange(12) % 2 == 0].copy()
    assert b.isbacked and b.shape == (30, 6) and list(b.var.index) == ["g%d" % j for j in range(0, 12, 2)]


@pytest.mark.parametrize("world,N,g", [(2, 4, 3000), (8, 4, 3000), (8, 8, 5000), (3, 2, 7), (8, 1, 5)])
def test_sharded_gene_solve_chunks_partition_the_columns(world, N, g):
    """engine._solve_chunk: the ranks' chunks tile the N*g columns of the V solve (and the g columns of W) exactly once,
    and every chunk is expressed as consecutive partial segments with the right first dataset."""
    from pyliger_b200.engine import InmfEngine
    for total, per_seg in ((N * g, g), (g, g)):
        covered = np.zeros(total, dtype=np.int64)
        for rank in range(world):
            e = InmfEngine.__new__(InmfEngine)
            e.world, e.rank, e._chunks, e.device, e.shard_min_columns = world, rank, {}, torch.device("cpu"), 0
            c0, c1, s0, offs, per = e._solve_chunk(total, per_seg)
            offs = offs.numpy()
            assert per * world >= total and 0 <= c0 <= c1 <= total
            covered[c0:c1] += 1
            if c1 > c0:
                assert offs[0] == c0 and offs[-1] == c1 and np.all(np.diff(offs) > 0) and s0 == c0 // per_seg
                for j in range(len(offs) - 1):   # partial segment j lies inside dataset s0 + j
                    assert offs[j] // per_seg == s0 + j and (offs[j + 1] - 1) // per_seg == s0 + j
        assert np.all(covered == 1)
    e = InmfEngine.__new__(InmfEngine)
    e.world, e.rank, e._chunks = 1, 0, {}
    assert e._solve_chunk(10, 5) is None


def test_window_statistics_work_lists_place_every_window_row_in_its_block():
    """blocked.window_stats_work (the online step's block-aligned statistics pass): for random dataset / mini-batch sizes,
    ranks and step counts, every window row lands at the position of its cell inside the slot of its block, no two rows
    share a position, every touched block is listed with all its gene ranges exactly once, and nothing else is listed."""
    from pyliger_b200 import blocked
    from pyliger_b200.engine import online_windows
    rng = np.random.RandomState(0)
    checked = 0
    for trial in range(120):
        N = rng.randint(1, 4)
        ns = [int(rng.randint(40, 400)) for _ in range(N)]
        mbs = [int(rng.randint(5, min(ns[i], 90))) for i in range(N)]
        world = rng.randint(1, 3)
        rank = rng.randint(0, world)
        sub = [((m * rank) // world, (m * (rank + 1)) // world) for m in mbs]
        T = rng.randint(1, 40)
        row_base = np.concatenate([[0], np.cumsum(ns)])
        plan = online_windows(T, mbs, sub, ns, row_base)
        cnt = [hi - lo for lo, hi in sub]
        cbs = [-(-max(c, 1) // (-(-max(c, 1) // 64))) for c in cnt]   # blocks of at most 64 cells
        nb = [-(-n // c) for n, c in zip(ns, cbs)]
        blk_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        nch, tb = 3, int(blk_off[-1])
        W = np.zeros((tb * nch, 8), dtype=np.int64)
        W[:, 0] = np.arange(tb * nch)                                 # item id in the slice_begin field
        W[:, 3] = np.repeat([min(cbs[s], ns[s] - b * cbs[s]) for s in range(N) for b in range(nb[s])], nch)
        res = blocked.window_stats_work(W, (nb, cbs, blk_off, nch), plan, ns)
        if res is None:   # a window wrapping onto its own first block: the row kernels take the run
            continue
        work, counts, place, region = res
        for t in range(T):
            slots = {}
            for r in range(counts[t]):
                blk = work[t, r, 0] // nch
                s = int(np.searchsorted(blk_off[1:], blk, side="right"))
                slots.setdefault(int(work[t, r, 2]), []).append((s, int(blk - blk_off[s]), int(work[t, r, 3])))
            assert all(len(v) == nch and len(set(v)) == 1 for v in slots.values())
            seen, touched = set(), set()
            for i in range(N):
                for q in range(cnt[i]):
                    row = (plan[t, i, 2] + q) % ns[i]
                    pos = int(place[t, plan[t, i, 0] + q])
                    assert pos not in seen and region[i] <= pos < region[i + 1]
                    seen.add(pos)
                    base = [b0 for b0, v in slots.items() if b0 <= pos < b0 + cbs[i] and v[0][0] == i]
                    assert len(base) == 1
                    s_, b_, rows_ = slots[base[0]][0]
                    assert b_ == row // cbs[i] and pos - base[0] == row - b_ * cbs[i] and pos - base[0] < rows_
                    touched.add(base[0])
            assert touched == set(slots)
            checked += 1
    assert checked > 500


def test_fused_reduction_counters_partition_the_work_items():
    """blocked.mc_work (pass 2 fused with its all-reduce): every work item belongs to exactly one (dataset, gene range)
    counter, all items of a counter cover the same rows of the same dataset, the counters' ranges tile every dataset's
    genes exactly once, expected[] counts the items, and the items that add a tile Gram (gram_seg >= 0) are exactly those
    of the first range of their dataset."""
    from pyliger_b200 import blocked

    class SP(object):
        pass
    rng = np.random.RandomState(3)
    for trial in range(50):
        N, g = rng.randint(1, 5), int(rng.randint(40, 900))
        nb = [int(rng.randint(0, 6)) for _ in range(N)]
        nch = int(rng.randint(1, 5))
        per = -(-g // nch)
        glo = np.arange(0, g, per)
        rows = []
        for s in range(N):
            for b in range(nb[s]):
                for ci, lo in enumerate(glo):
                    rows.append([0, 0, 0, 0, s * g, s if ci == 0 else -1, b, 0, lo, min(g, lo + per)])
        if not rows:
            continue
        rows = np.asarray(rows, dtype=np.int64)
        rng.shuffle(rows)
        sp = SP()
        sp.work_host, sp.row_lo, sp.row_hi = rows[:, :8].copy(), rows[:, 8].copy(), rows[:, 9].copy()
        W, expected = blocked.mc_work(sp, g)
        assert expected.sum() == len(W) and W[:, 6].min() == 0 and W[:, 6].max() == len(expected) - 1
        cover = {}
        for c in range(len(expected)):
            items = W[W[:, 6] == c]
            assert len(items) == expected[c]
            assert len(set(map(tuple, items[:, [4, 7]]))) == 1          # one dataset, one row range
            lo, hi = int(items[0, 7] >> 32), int(items[0, 7] & 0xffffffff)
            s = int(items[0, 4]) // g
            cover.setdefault(s, []).append((lo, hi))
            assert ((items[:, 5] >= 0) == (lo == 0)).all() and (items[items[:, 5] >= 0, 5] == s).all()
            assert expected[c] == nb[s]
        for s, ranges in cover.items():
            ranges.sort()
            assert ranges[0][0] == 0 and ranges[-1][1] == g
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
