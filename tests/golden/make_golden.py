"""Generate the golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

Run only in the development container (needs /root/reference):

    python tests/golden/make_golden.py

The reference (welch-lab/pyliger, pure Python) is imported through ``oracle/ref_shims.py``; no
reference source is copied.  Inputs are small seeded matrices that are stored in the .npz files next
to the reference's outputs, so the tests never need the reference again.
"""
import contextlib
import io
import os
import sys

import numpy as np
import pandas  # noqa: F401  (before the shims: they stub pandas' optional numexpr)
import scipy.sparse as sps
import sklearn.preprocessing  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from pyliger_b200.synth import make_datasets  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref_als, ref_online, ref_util, ref_onl_mod, RefLiger = ref_shims.load()
AnnData = ref_shims.AnnData


def csr_pack(prefix, X, d):
    X = sps.csr_matrix(X)
    d[prefix + "_data"] = X.data
    d[prefix + "_indices"] = X.indices.astype(np.int32)
    d[prefix + "_indptr"] = X.indptr.astype(np.int64)
    d[prefix + "_shape"] = np.asarray(X.shape, dtype=np.int64)


def make_liger(mats):
    adatas = []
    for i, m in enumerate(mats):
        a = AnnData(X=m, uns={"sample_name": "s%d" % i}, layers={"scale_data": m})
        adatas.append(a)
    lig = RefLiger(adatas)
    lig.var_genes = ["g%d" % j for j in range(mats[0].shape[1])]
    return lig


def gen_bpp():
    d = {}
    # KAT-1 / KAT-2 of SURVEY.md Appendix C
    A = np.array([[1.0, 0], [0, 1], [1, 1]])
    B = np.array([[1.0, 2], [-1, 0.5], [0, 3]])
    X, info = ref_util.nnlsm_blockpivot(A, B)
    d.update(kat1_A=A, kat1_B=B, kat1_X=X, kat1_Y=info[1], kat1_info=np.array([info[0], info[2], info[3], info[4]]))
    np.random.seed(0)
    A = np.random.uniform(0, 2, (8, 4))
    B = np.random.uniform(-1, 2, (8, 3))
    X, info = ref_util.nnlsm_blockpivot(A, B)
    d.update(kat2_A=A, kat2_B=B, kat2_X=X, kat2_Y=info[1], kat2_info=np.array([info[0], info[2], info[3], info[4]]))
    # random product-form problems at the factor counts of the configs (H-update shaped)
    cases = []
    for k, g, cols, seed in ((5, 40, 64, 1), (20, 200, 300, 2), (30, 300, 300, 3), (40, 300, 300, 4),
                             (60, 400, 200, 5), (64, 400, 100, 6), (33, 100, 97, 7)):
        rng = np.random.default_rng(seed)
        W = np.abs(rng.uniform(0, 2, (k, g)))
        V = np.abs(rng.uniform(0, 2, (k, g))) * 0.5
        Xc = sps.random(cols, g, density=0.1, random_state=seed, data_rvs=lambda s: rng.gamma(2.0, 1.0, s)).tocsr()
        G = (W + V) @ (W + V).T + 5.0 * V @ V.T
        R = np.asarray(Xc.dot((W + V).T)).T
        X, info = ref_util.nnlsm_blockpivot(G, R, is_input_prod=True)
        tag = "prod%d" % k
        d[tag + "_G"] = G
        d[tag + "_R"] = R
        d[tag + "_X"] = X
        d[tag + "_Y"] = info[1]
        d[tag + "_info"] = np.array([info[0], info[2], info[3], info[4]])
        cases.append(tag)
    # i.i.d. variant with negative right-hand sides and an init= warm start
    rng = np.random.default_rng(11)
    A = rng.uniform(0, 1, (300, 20))
    B = rng.uniform(-0.2, 0.8, (300, 150))
    X, info = ref_util.nnlsm_blockpivot(A, B)
    d.update(iid_A=A, iid_B=B, iid_X=X, iid_Y=info[1], iid_info=np.array([info[0], info[2], info[3], info[4]]))
    init = np.where(rng.random(X.shape) < 0.9, X, 1.0 - X)  # mostly right, some wrong
    X2, info2 = ref_util.nnlsm_blockpivot(A, B, init=init)
    d.update(iid_init=init, iid_init_X=X2, iid_init_Y=info2[1],
             iid_init_info=np.array([info2[0], info2[2], info2[3], info2[4]]))
    # a hard, ill-conditioned problem that exercises the backup rule
    rng = np.random.default_rng(5)
    base = rng.uniform(0, 1, (40, 3))
    A = np.hstack([base @ rng.uniform(0, 1, (3, 12)) + 1e-3 * rng.uniform(0, 1, (40, 12))])
    B = rng.uniform(-1, 1, (40, 200))
    X, info = ref_util.nnlsm_blockpivot(A, B)
    d.update(hard_A=A, hard_B=B, hard_X=X, hard_Y=info[1], hard_info=np.array([info[0], info[2], info[3], info[4]]))
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "bpp.npz"), **d)
    print("bpp: backup counts", {c: int(d[c + "_info"][3]) for c in cases}, "hard", d["hard_info"], "iid", d["iid_info"])


def run_ref_als(mats, **kw):
    lig = make_liger(mats)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        ref_als(lig, print_obj=True, **kw)
    obj = float(buf.getvalue().strip().split("Objective:")[-1])
    H = [np.array(a.obsm["H"]) for a in lig.adata_list]
    W = np.array(lig.adata_list[0].varm["W"])
    V = [np.array(a.varm["V"]) for a in lig.adata_list]
    return obj, H, W, V


def gen_als():
    d = {}
    mats = make_datasets([310, 260], 150, 6, density=0.12, structured=True, seed0=2000)
    T = 12
    k = 6
    objs = []
    for t in range(1, T + 1):
        obj, H, W, V = run_ref_als(mats, k=k, value_lambda=5.0, thresh=1e-6, max_iters=t, rand_seed=1)
        objs.append(obj)
    obj0, _, _, _ = run_ref_als(mats, k=k, value_lambda=5.0, thresh=1e-6, max_iters=0, rand_seed=1)
    for i, m in enumerate(mats):
        csr_pack("X%d" % i, m, d)
        d["H%d" % i] = H[i]
        d["V%d" % i] = V[i]
    d["W"] = W
    d["objectives"] = np.array([obj0] + objs)
    d["params"] = np.array([k, 5.0, 1e-6, T, 1], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "als_small.npz"), **d)
    print("als objectives", d["objectives"])

    # three datasets, other lambda / seed, nrep=2, with user-supplied V_init
    d = {}
    mats = make_datasets([120, 150, 90], 80, 4, density=0.15, structured=True, seed0=2100)
    k = 4
    obj, H, W, V = run_ref_als(mats, k=k, value_lambda=2.5, thresh=1e-4, max_iters=8, rand_seed=7, nrep=2)
    for i, m in enumerate(mats):
        csr_pack("X%d" % i, m, d)
        d["H%d" % i] = H[i]
        d["V%d" % i] = V[i]
    d["W"] = W
    d["objective"] = np.array(obj)
    d["params"] = np.array([k, 2.5, 1e-4, 8, 7, 2], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "als_three.npz"), **d)
    print("als_three objective", obj)


def gen_online():
    d = {}
    # helper KATs
    d["kat4_initW"] = ref_util._init_W(3, 2, 1)
    idx = ref_onl_mod._generate_idx(4, 2500, 1, 1000, 7000)
    d["kat5_idx"] = np.array([[list(map(int, p)) for p in idx[i]] for i in range(4)], dtype=np.int64)
    idx2 = ref_onl_mod._generate_idx(7, 130, 2, 100, 333)
    flat = []
    for i in range(7):
        for (lo, hi) in idx2[i]:
            flat.append((i, int(lo), int(hi)))
    d["idx2_flat"] = np.array(flat, dtype=np.int64)
    A = np.eye(2)
    B = np.ones((3, 2))
    A_old = np.zeros((2, 2))
    B_old = np.zeros((3, 2))
    Hm = np.array([[1.0, 2], [0, 1]])
    Xm = np.arange(6, dtype=np.float64).reshape(3, 2)
    seqA, seqB, seqAo, seqBo = [], [], [], []
    for (it, ep, epp) in ((0, 0, 0), (1, 0, 0), (2, 1, 0), (3, 1, 1)):
        A, B, A_old, B_old = ref_onl_mod._update_A_B(A, B, A_old, B_old, Hm, Xm, 2, it, ep, epp)
        seqA.append(A.copy()); seqB.append(B.copy()); seqAo.append(A_old.copy()); seqBo.append(B_old.copy())
    d.update(kat6_A=np.array(seqA), kat6_B=np.array(seqB), kat6_Aold=np.array(seqAo), kat6_Bold=np.array(seqBo))
    # one HALS sweep on random inputs
    rng = np.random.default_rng(3)
    g, k, N = 37, 5, 3
    W = rng.uniform(0, 1, (g, k)); Vs = [rng.uniform(0, 1, (g, k)) for _ in range(N)]
    As = []
    for _ in range(N):
        Hh = rng.uniform(0, 1, (k, 50)); As.append(Hh @ Hh.T / 50)
    Bs = [rng.uniform(0, 1, (g, k)) for _ in range(N)]
    d.update(hals_W0=W.copy(), hals_V0=np.array(Vs), hals_A=np.array(As), hals_B=np.array(Bs))
    W1 = ref_util._update_W_HALS(As, Bs, W.copy(), [v.copy() for v in Vs])
    V1 = ref_util._update_V_HALS(As, Bs, W1.copy(), [v.copy() for v in Vs], 5.0)
    d.update(hals_W1=W1, hals_V1=np.array(V1))

    # scenario 1 end to end
    mats = make_datasets([420, 300], 120, 5, density=0.12, structured=True, seed0=2200)
    lig = make_liger(mats)
    with contextlib.redirect_stderr(io.StringIO()):
        ref_online(lig, k=5, value_lambda=5.0, max_epochs=3, miniBatch_max_iters=1, miniBatch_size=144,
                   h5_chunk_size=100, rand_seed=1, verbose=False)
    for i, m in enumerate(mats):
        csr_pack("X%d" % i, m, d)
        a = lig.adata_list[i]
        d["s1_H%d" % i] = np.array(a.obsm["H"]); d["s1_V%d" % i] = np.array(a.varm["V"])
        d["s1_B%d" % i] = np.array(a.varm["B"]); d["s1_A%d" % i] = np.array(a.uns["A"])
    d["s1_W"] = np.array(lig.adata_list[0].varm["W"])
    d["s1_params"] = np.array([5, 5.0, 3, 1, 144, 100, 1], dtype=np.float64)
    d["s1_initV0"] = ref_util._init_V_online(420, 5, {"scale_data": mats[0]}, 100, 1)

    # scenario 1 with miniBatch_max_iters=2, one dataset
    lig1 = make_liger([mats[0]])
    with contextlib.redirect_stderr(io.StringIO()):
        ref_online(lig1, k=4, value_lambda=3.0, max_epochs=2, miniBatch_max_iters=2, miniBatch_size=100,
                   h5_chunk_size=64, rand_seed=3, verbose=False)
    a = lig1.adata_list[0]
    d.update(s1b_H=np.array(a.obsm["H"]), s1b_W=np.array(a.varm["W"]), s1b_V=np.array(a.varm["V"]),
             s1b_A=np.array(a.uns["A"]), s1b_B=np.array(a.varm["B"]))

    # scenario 2 (refine with a new dataset) and scenario 3 (projection) on top of scenario 1
    new = make_datasets([260, 420, 300], 120, 5, density=0.12, structured=True, seed0=2199)[0]
    # NB: gene filtering may differ; regenerate consistently with the same gene set
    if new.shape[1] != mats[0].shape[1]:
        new = sps.random(260, mats[0].shape[1], density=0.12, random_state=9,
                         data_rvs=lambda s: np.random.default_rng(9).gamma(2.0, 1.0, s)).tocsr()
    csr_pack("Xnew", new, d)
    a_new = AnnData(X=new, uns={"sample_name": "new"}, layers={"scale_data": new})
    import copy
    lig2 = make_liger(mats)
    for i, a in enumerate(lig2.adata_list):
        a.obsm["H"] = d["s1_H%d" % i].copy(); a.varm["W"] = d["s1_W"].copy()
        a.varm["V"] = d["s1_V%d" % i].copy(); a.varm["B"] = d["s1_B%d" % i].copy(); a.uns["A"] = d["s1_A%d" % i].copy()
    with contextlib.redirect_stderr(io.StringIO()):
        ref_online(lig2, X_new=[a_new], k=5, value_lambda=5.0, max_epochs=2, miniBatch_size=80,
                   h5_chunk_size=100, rand_seed=1, verbose=False)
    d.update(s2_Hnew=np.array(a_new.obsm["H"]), s2_W=np.array(a_new.varm["W"]), s2_Vnew=np.array(a_new.varm["V"]),
             s2_Anew=np.array(a_new.uns["A"]), s2_Bnew=np.array(a_new.varm["B"]))
    for i in range(2):
        d["s2_Hold%d" % i] = np.array(lig2.adata_list[i].obsm["H"])
    d["s2_n_datasets"] = np.array(len(lig2.adata_list))

    lig3 = make_liger(mats)
    for i, a in enumerate(lig3.adata_list):
        a.varm["W"] = d["s1_W"].copy()
    a_proj = AnnData(X=new, uns={"sample_name": "proj"}, layers={"scale_data": new})
    ref_online(lig3, X_new=[a_proj], projection=True, k=5, miniBatch_size=100, verbose=False)
    d.update(s3_H=np.array(a_proj.obsm["H"]), s3_V=np.array(a_proj.varm["V"]), s3_W=np.array(a_proj.varm["W"]))
    np.savez_compressed(os.path.join(OUT, "online.npz"), **d)
    print("online done; W sum", d["s1_W"].sum(), "s2 W sum", d["s2_W"].sum())


def gen_preprocess():
    """normalize + scale_not_center of the reference (preprocessing/_normalization.py, _scale.py), in-memory
    mode, on raw counts held by the duck-typed AnnData of pyliger_b200.containers."""
    from pyliger.preprocessing._normalization import normalize as ref_normalize
    from pyliger.preprocessing._scale import scale_not_center as ref_scale
    from pyliger_b200.containers import AnnDataLike

    rng = np.random.default_rng(77)
    g = 40
    var_names = ["gene%d" % j for j in range(g)]
    raws = []
    for i, n in enumerate((60, 45)):
        C = rng.poisson(0.35, size=(n, g)).astype(np.float64)
        C[np.arange(n), rng.integers(0, g, n)] += 1.0  # no empty cells
        C[:, 3 + i] = 0.0                              # a gene absent from this dataset
        raws.append(sps.csr_matrix(C))
    var_genes = [var_names[j] for j in sorted(rng.choice(g, 25, replace=False))]
    lig = RefLiger([AnnDataLike.from_raw(m, "s%d" % i, var_names=var_names) for i, m in enumerate(raws)])
    with contextlib.redirect_stdout(io.StringIO()):
        ref_normalize(lig)
    d = {"var_genes": np.asarray(var_genes), "var_names": np.asarray(var_names)}
    for i, a in enumerate(lig.adata_list):
        csr_pack("raw%d" % i, raws[i], d)
        csr_pack("norm%d" % i, a.layers["norm_data"], d)
        for col in ("norm_sum", "norm_sum_sq", "norm_mean"):
            d["%s%d" % (col, i)] = a.var[col].to_numpy()
    lig.var_genes = var_genes
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(divide="ignore"):
        ref_scale(lig)
    for i, a in enumerate(lig.adata_list):
        csr_pack("scale%d" % i, a.layers["scale_data"], d)
        d["scaled_var_names%d" % i] = np.asarray(list(a.var.index), dtype=str)
        d["raw_shape%d" % i] = np.asarray(a.raw.shape)
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **d)
    print("preprocess done; scale_data sums", [float(a.layers["scale_data"].sum()) for a in lig.adata_list])


class _DuckLiger(object):
    """iNMF_HALS only touches ``adata_list`` and assigns ``W``: the reference's own Liger class makes that assignment
    fail (read-only property, _iNMF_HALS.py:156), so the unmodified function is run on a plain object instead."""


def gen_hals():
    from pyliger.factorization._iNMF_HALS import iNMF_HALS as ref_hals
    d = {}
    cases = {"a": dict(ns=[260, 210], g=120, k=5, lam=5.0, thresh=1e-4, max_iters=12, nrep=1, seed=1),
             "b": dict(ns=[150, 170, 90], g=90, k=7, lam=2.0, thresh=1e-6, max_iters=6, nrep=2, seed=3)}
    for tag, c in cases.items():
        mats = make_datasets(c["ns"], c["g"], c["k"], density=0.15, structured=True, seed0=2300 + len(c["ns"]))
        lig = _DuckLiger()
        lig.adata_list = [AnnData(X=m, layers={"scale_data": m}) for m in mats]
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref_hals(lig, k=c["k"], value_lambda=c["lam"], thresh=c["thresh"], max_iters=c["max_iters"], nrep=c["nrep"],
                     rand_seed=c["seed"])
        objs = [float(line.split("Obj:")[-1]) for line in buf.getvalue().splitlines() if "Obj:" in line]
        for i, m in enumerate(mats):
            csr_pack("%s_X%d" % (tag, i), m, d)
            d["%s_H%d" % (tag, i)] = np.array(lig.adata_list[i].obsm["H"])
            d["%s_V%d" % (tag, i)] = np.array(lig.adata_list[i].varm["V"])
        d[tag + "_W"] = np.array(lig.adata_list[0].varm["W"])
        d[tag + "_objectives"] = np.array(objs)
        d[tag + "_params"] = np.array([len(mats), c["k"], c["lam"], c["thresh"], c["max_iters"], c["nrep"], c["seed"]],
                                      dtype=np.float64)
        print("hals", tag, "objectives", objs[:3], "...", objs[-1], "n =", len(objs))
    np.savez_compressed(os.path.join(OUT, "hals.npz"), **d)


def gen_quantile():
    """quantile_norm of the unmodified reference on H matrices of a short optimize_ALS run (oracle), three cases:
    default options on small clusters (no subsampling), max_sample below the cluster sizes (the permutation stream
    matters), and do_center / dims_use / no kNN refinement."""
    from pyliger.tools._quantile_norm import quantile_norm as ref_qn
    from oracle import inmf_oracle as orc
    d = {}
    mats = make_datasets([700, 900, 500], 150, 6, density=0.15, structured=True, seed0=2500)
    res = orc.als(mats, 6, 5.0, 1e-6, 8, rand_seed=1)
    Hs = [np.ascontiguousarray(h) for h in res["H"]]
    for i, h in enumerate(Hs):
        d["H%d" % i] = h
    cases = {"a": dict(quantiles=50, min_cells=20, max_sample=1000, knn_k=20, refine_knn=True, do_center=False,
                       dims_use=None, rand_seed=1),
             "b": dict(quantiles=20, min_cells=5, max_sample=60, knn_k=15, refine_knn=True, do_center=False,
                       dims_use=None, rand_seed=7),
             "c": dict(quantiles=30, min_cells=10, max_sample=100, knn_k=20, refine_knn=False, do_center=True,
                       dims_use=[0, 2, 3, 5], rand_seed=3)}
    for tag, kw in cases.items():
        lig = _DuckLiger()
        lig.adata_list = []
        for i, h in enumerate(Hs):
            a = AnnData(X=mats[i], obs={}, uns={"sample_name": "s%d" % i})
            a.obsm["H"] = h.copy()
            lig.adata_list.append(a)
        ref_qn(lig, **kw)
        for i, a in enumerate(lig.adata_list):
            d["%s_cluster%d" % (tag, i)] = np.asarray(a.obs["cluster"])
            d["%s_Hnorm%d" % (tag, i)] = np.asarray(a.obsm["H_norm"])
        print("quantile", tag, [np.bincount(np.asarray(a.obs["cluster"]), minlength=6).tolist() for a in lig.adata_list])
    np.savez_compressed(os.path.join(OUT, "quantile.npz"), **d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["bpp", "als", "online", "preprocess", "hals", "quantile"]
    for name in which:
        {"bpp": gen_bpp, "als": gen_als, "online": gen_online, "preprocess": gen_preprocess, "hals": gen_hals,
         "quantile": gen_quantile}[name]()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
