"""online_iNMF on B200: same signature, scenarios and write-back as pyliger
(/root/reference/src/pyliger/factorization/_online_iNMF.py:23-601), CUDA kernels underneath.

Scenario 1 (X_new is None)            -- factorise liger_object's datasets from scratch
Scenario 2 (X_new, projection=False)  -- refine W with new datasets, previous A/B/V enter the W update
Scenario 3 (X_new, projection=True)   -- project new datasets onto the existing W (V = 0)

Backed AnnData (``adata.isbacked``): the reference reads its mini-batches from ``./results/<sample>.hdf5`` through
h5sparse (_online_iNMF.py:410-412, 480-490, 555-560).  Here the store is ``pyliger_b200.io.CSRStore`` (same slicing
API; a real h5sparse file also works when h5sparse is installed) and the cells are STREAMED: X never becomes resident,
the rows of the next mini-batch are gathered and copied to the device while the current one is processed
(``pyliger_b200/streaming.py``); ``stream=True`` does the same for in-memory matrices that should not be uploaded
whole.  X_new without a ``scale_data`` layer is preprocessed like the reference does (_online_iNMF.py:604-620) into a
store (``preprocessing/_backed.py``).
Keyword-only extensions: ``dtype`` ("float64" default / "float32"), ``device``, ``stream``.
"""
import numpy as np
from tqdm import tqdm

import torch

from ..engine import InmfEngine, resolve_dtype
from ..streaming import StreamingEngine


def _h5_idx_generator(chunk_size, matrix_size):
    """Row chunks [0,c), [c,2c), ... clipped at matrix_size (pyliger/_utilities.py:100-116)."""
    lo = 0
    while lo < matrix_size:
        yield lo, min(lo + chunk_size, matrix_size)
        lo += chunk_size


def _init_W(num_genes, k, rand_seed):
    """Host RNG for bit-identical streams (_utilities.py:16-26)."""
    np.random.seed(seed=rand_seed)
    W = np.abs(np.random.uniform(0, 2, (num_genes, k)))
    return W / np.sqrt(np.sum(np.square(W), axis=0))


def _init_V_online(num_cell, k, X, chunk_size, rand_seed):
    """k cells sampled with replacement (sorted) as unit-norm columns (_utilities.py:69-88)."""
    np.random.seed(seed=rand_seed)
    idx = np.sort(np.random.choice(list(range(num_cell)), k))
    V = X[idx, :].transpose().toarray()
    return V / np.sqrt(np.sum(np.square(V), axis=0))


def _scale_data(adata):
    """The reference's ``_get_scale_data`` (_online_iNMF.py:555-560): the on-disk store of a backed AnnData, the layer
    of an in-memory one."""
    if getattr(adata, "isbacked", False):
        from ..io import open_backed
        return open_backed(adata, "r")["scale_data"]
    return adata.layers["scale_data"]


def _is_streamed(X):
    return not hasattr(X, "indptr")  # anything that is not an in-memory scipy matrix


def _gather_own_rows(engine, outs):
    """Per-rank row blocks of every dataset (host arrays) -> full per-dataset arrays on every rank."""
    if engine.world == 1:
        return outs
    from ..engine import all_gather_rows, shard_bounds
    full = []
    for i, o in enumerate(outs):
        sizes = [shard_bounds(engine.n_cells[i], r, engine.world) for r in range(engine.world)]
        t = torch.from_numpy(o).to(engine.device)
        full.append(all_gather_rows(t, [hi - lo for lo, hi in sizes], engine.group).cpu().numpy())
    return full


def _final_H(engine, with_V=True):
    """H of every cell of every dataset as host float64 arrays, from the engine's final W, V."""
    if isinstance(engine, StreamingEngine):
        return _gather_own_rows(engine, engine.stream_final_H(with_V=with_V))
    engine.final_H(with_V=with_V)
    return engine.gather_H()


def _to_dev(engine, arr):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=np.float64))).to(
        engine.device, engine.dtype)


def _cal_W_V(Xs, num_genes, num_cells, prev_info, W_init, V_init, A_init, B_init, k, value_lambda, max_epochs,
             miniBatch_max_iters, miniBatch_size, h5_chunk_size, rand_seed, dtype, device, progress, stream=False):
    """The mini-batch loop of _online_iNMF_cal_W_V (_online_iNMF.py:334-462). Returns the engine,
    holding W, V, A, B on the device."""
    num_files = len(Xs)
    np.random.seed(seed=rand_seed)
    W = _init_W(num_genes, k, rand_seed) if W_init is None else W_init
    Vs = ([_init_V_online(num_cells[i], k, Xs[i], h5_chunk_size, rand_seed) for i in range(num_files)]
          if V_init is None else V_init)
    n_prev = 0 if prev_info is None else len(prev_info["A_prev"])
    stream = stream or any(_is_streamed(X) for X in Xs)
    if stream:  # out-of-core: the cells stay in the store / on the host, mini-batches are streamed
        engine = StreamingEngine(Xs, k, value_lambda, dtype=dtype, device=device, n_prev=n_prev)
    else:
        engine = InmfEngine(Xs, k, value_lambda, dtype=dtype, device=device, replicate=True, n_prev=n_prev,
                            need_stats=False)  # mini-batch statistics go through the window kernels, not pass 2

    miniBatch_sizes = np.asarray(
        [np.round((num_cells[i] / np.sum(num_cells)) * miniBatch_size).astype(int) for i in range(num_files)])
    total_iters = np.floor(num_cells[0] * max_epochs / miniBatch_sizes[0]).astype(int)
    engine.alloc_online(miniBatch_sizes)
    engine.Wt.copy_(_to_dev(engine, W))
    for i in range(num_files):
        engine.Vt[i].copy_(_to_dev(engine, Vs[i]))
        if A_init is not None:
            engine.A[i].copy_(torch.from_numpy(np.asarray(A_init[i], dtype=np.float64)).to(engine.device))
        if B_init is not None:
            engine.B[i].copy_(_to_dev(engine, B_init[i]))
    for p in range(n_prev):
        engine.A_all[p].copy_(torch.from_numpy(np.asarray(prev_info["A_prev"][p], dtype=np.float64)).to(engine.device))
        engine.B_all[p].copy_(_to_dev(engine, prev_info["B_prev"][p]))
        engine.Vt_all[p].copy_(_to_dev(engine, prev_info["V_prev"][p]))

    def step_args(it):
        """scale_param of _update_A_B (_online_iNMF.py:576-581), including its `1/miniBatch_size` at iter 1, and
        whether this step starts a new epoch."""
        epoch_prev = (it * miniBatch_sizes[0]) // num_cells[0]
        epoch = ((it + 1) * miniBatch_sizes[0]) // num_cells[0]
        if it == 0:
            scales = [0.0] * num_files
        elif it == 1:
            scales = [1.0 / m for m in miniBatch_sizes]
        else:
            scales = [(it - 1) / it] * num_files
        return scales, bool(epoch > 0 and epoch - epoch_prev == 1)

    if stream:
        engine.stream_minibatches(int(total_iters), step_args, miniBatch_max_iters)
        return engine
    plan = engine.online_plan(int(total_iters))
    it_range = tqdm(range(total_iters)) if progress else range(total_iters)
    for it in it_range:
        scales, rollover = step_args(it)
        engine.online_step(plan[it], scales, rollover, miniBatch_max_iters, step=it)
    return engine


def _write_back_W_V(engine, W_init, V_init):
    """HALS mutates caller-supplied W_init / V_init in place in the reference (_utilities.py:113,123):
    mirror that, and return the host arrays that are written back."""
    W = engine.host_W()
    Vs = engine.host_V()
    if W_init is not None:
        W_init[...] = W
        W = W_init
    if V_init is not None:
        for i in range(len(Vs)):
            V_init[i][...] = Vs[i]
            Vs[i] = V_init[i]
    return W, Vs


def _host_A_B(engine):
    A = [engine.A[i].cpu().numpy() for i in range(engine.N)]
    B = [engine.B[i].to(torch.float64).cpu().numpy() for i in range(engine.N)]
    return A, B


def online_iNMF(
    liger_object,
    X_new=None,
    projection=False,
    W_init=None,
    V_init=None,
    H_init=None,
    A_init=None,
    B_init=None,
    k=20,
    value_lambda=5.0,
    max_epochs=5,
    miniBatch_max_iters=1,
    miniBatch_size=5000,
    h5_chunk_size=1000,
    rand_seed=1,
    verbose=True,
    *,
    dtype="float64",
    device=None,
    stream=False,
):
    dt = resolve_dtype(dtype)
    num_genes = len(liger_object.var_genes)
    params = dict(k=k, value_lambda=value_lambda, max_epochs=max_epochs, miniBatch_max_iters=miniBatch_max_iters,
                  miniBatch_size=miniBatch_size, h5_chunk_size=h5_chunk_size, rand_seed=rand_seed, dtype=dt,
                  device=device, progress=verbose, stream=stream)

    def h_engine(X, lam):
        """Engine for an H-only pass over one dataset (projection / refresh): streamed if X is (or should be)."""
        if stream or _is_streamed(X):
            return StreamingEngine([X], k, lam, dtype=dt, device=device)
        return InmfEngine([X], k, lam, dtype=dt, device=device, need_stats=False)

    if X_new is None:  # ---------------- scenario 1 (_online_iNMF.py:125-178)
        if verbose:
            print("Starting Online iNMF...")
        Xs = [_scale_data(adata) for adata in liger_object.adata_list]
        num_cells = [adata.shape[0] for adata in liger_object.adata_list]
        engine = _cal_W_V(Xs, num_genes, num_cells, None, W_init, V_init, A_init, B_init, **params)
        if verbose:
            print("Calculate metagene loadings...")
        Hs = _final_H(engine)
        W, Vs = _write_back_W_V(engine, W_init, V_init)
        A, B = _host_A_B(engine)
        for i, adata in enumerate(liger_object.adata_list):
            adata.obsm["H"] = Hs[i]
            adata.varm["W"] = W
            adata.varm["V"] = Vs[i]
            adata.varm["B"] = B[i]
            adata.uns["A"] = A[i]
        return None

    # raw datasets are preprocessed into an on-disk store first (_online_iNMF.py:184-190, 275-281, 604-620)
    X_new = list(X_new)
    for i, adata in enumerate(X_new):
        has_scale = "scale_data" in adata.layers.keys()
        if not has_scale and getattr(adata, "isbacked", False):
            from ..io import open_backed
            has_scale = "scale_data" in open_backed(adata, "r").keys()
        if not has_scale:
            from ..preprocessing._backed import preprocessing
            X_new[i] = preprocessing(adata, h5_chunk_size, liger_object.var_genes, dtype="float64", device=device)

    if projection:  # ---------------- scenario 3 (_online_iNMF.py:271-331)
        if verbose:
            print("Metagene projection")
        W = liger_object.W if W_init is None else W_init
        for adata in X_new:
            engine = h_engine(_scale_data(adata), 0.0)
            engine.Wt.copy_(_to_dev(engine, W))
            engine.Vt.zero_()
            adata.obsm["H"] = _final_H(engine, with_V=False)[0]
            adata.varm["W"] = W
            adata.varm["V"] = np.zeros((num_genes, k))
            liger_object.adata_list.append(adata)
        return None

    # ---------------- scenario 2 (_online_iNMF.py:181-268)
    if verbose:
        print("{} new datasets detected.".format(len(X_new)))
    if W_init is None:
        W_init = liger_object.W
    prev_info = {
        "A_prev": [adata.uns["A"] for adata in liger_object.adata_list],
        "B_prev": [adata.varm["B"] for adata in liger_object.adata_list],
        "V_prev": [adata.varm["V"] for adata in liger_object.adata_list],
    }
    W = W_init
    for adata in X_new:
        engine = _cal_W_V([_scale_data(adata)], num_genes, [adata.shape[0]], prev_info, W, V_init, A_init, B_init,
                          **params)
        if verbose:
            print("Calculate metagene loadings...")
        Hnew = _final_H(engine)[0]
        W, Vs = _write_back_W_V(engine, W, V_init)
        A, B = _host_A_B(engine)
        adata.obsm["H"] = Hnew
        adata.varm["W"] = W
        adata.varm["V"] = Vs[0]
        adata.varm["B"] = B[0]
        adata.uns["A"] = A[0]
    # refresh H of the existing datasets with the refined W and their own V
    for adata in liger_object.adata_list:
        engine = h_engine(_scale_data(adata), value_lambda)
        engine.Wt.copy_(_to_dev(engine, W))
        engine.Vt[0].copy_(_to_dev(engine, adata.varm["V"]))
        adata.obsm["H"] = _final_H(engine)[0]
        adata.varm["W"] = W
    for adata in X_new:
        liger_object.adata_list.append(adata)
    return None
