"""iNMF_HALS on B200: same signature, algorithm and write-back as pyliger's batch HALS factorisation
(/root/reference/src/pyliger/factorization/_iNMF_HALS.py:16-163), CUDA kernels underneath.

Per iteration (reference :103-137): one HALS sweep over H (``_update_H_HALS``, _utilities.py:132-148), the statistics
X H', H H' (:116-117), one HALS sweep over W then over every V_i (``_update_W_HALS`` / ``_update_V_HALS``,
_utilities.py:102-129), the objective ||X - (W+V)H||^2 + lambda ||VH||^2 (:126-131).  On the device that is
pass 1 (right-hand sides) + ``inmf_hals_h`` + pass 2 (statistics) + ``inmf_hals`` + the statistics-form objective;
no n x g temporary exists.

Reference quirks kept: the seeded initialisation order (``np.random.seed(rand_seed)`` once, then per restart V from
``np.random.choice`` columns, W from a re-seeded stream, H from the continued stream, _utilities.py:16-38, 91-94); the
result of an iteration is only kept if the objective did not increase (:147-150); restarts do not compare objectives
(the last restart's factors are written back).
Reference bug side-stepped: ``liger_object.W = final_W`` (:156) assigns to a read-only property of the ``Liger``
class and raises AttributeError before anything is written back; here H, W, V are written to every dataset's AnnData
(:158-161) and the attribute assignment is skipped.
Keyword-only extensions: ``dtype`` ("float64" default / "float32"), ``device``, ``verbose`` (the reference prints).
"""
import time

import numpy as np
import torch

from ..engine import InmfEngine, resolve_dtype


def _init_V(num_cells, num_samples, k, Xs):
    """k cells of every dataset (with replacement, numpy's global stream) as unit-norm columns (_utilities.py:29-38).
    ``Xs`` are the cells x genes matrices."""
    V = [np.asarray(Xs[i][np.random.choice(list(range(num_cells[i])), k), :].todense()).T for i in range(num_samples)]
    return [v / np.sqrt(np.sum(np.square(v), axis=0)) for v in V]


def _init_W(num_genes, k, rand_seed):
    np.random.seed(seed=rand_seed)
    W = np.abs(np.random.uniform(0, 2, (num_genes, k)))
    return W / np.sqrt(np.sum(np.square(W), axis=0))


def _init_H(num_cells, num_samples, k):
    return [np.random.uniform(0, 2, (k, num_cells[i])) for i in range(num_samples)]


def iNMF_HALS(liger_object, k=20, value_lambda=5.0, thresh=1e-4, max_iters=25, nrep=1, rand_seed=1, *,
              dtype="float64", device=None, verbose=True, return_info=False):
    num_samples = len(liger_object.adata_list)
    Xs = [adata.layers["scale_data"] for adata in liger_object.adata_list]
    num_cells = [x.shape[0] for x in Xs]
    num_genes = Xs[0].shape[1]
    engine = InmfEngine(Xs, k, value_lambda, dtype=resolve_dtype(dtype), device=device)
    if engine.world > 1:
        raise NotImplementedError("iNMF_HALS runs on one GPU (the multi-GPU paths are optimize_ALS and online_iNMF)")
    np.random.seed(seed=rand_seed)
    final = None
    info = {"objectives": [], "iters_run": []}
    for _ in range(nrep):
        V = _init_V(num_cells, num_samples, k, Xs)
        W = _init_W(num_genes, k, rand_seed)
        H = _init_H(num_cells, num_samples, k)
        engine.set_als_state(W.T, [v.T for v in V], [h.T for h in H])
        obj_train = engine.initial_objective()
        if verbose:
            print("Initial Training Obj: {}".format(obj_train))
        objs = [obj_train]
        iteration = 1
        total_time = 0
        delta = np.inf
        while delta > thresh and iteration <= max_iters:
            t0 = time.time()
            engine.hals_H()          # uses G, W + V of the last prep()
            engine.stats()           # X'H, H'H of the new H
            engine.hals_WV()
            engine.prep()
            obj_prev = obj_train
            obj_train = float(engine.objective().item())
            objs.append(obj_train)
            delta = np.absolute(obj_prev - obj_train) / ((obj_prev + obj_train) / 2)
            total_time += time.time() - t0
            if verbose:
                print("Iter: {}, Total time: {}, Training Obj: {}".format(iteration, total_time, obj_train))
            if obj_train <= obj_prev:
                final = (engine.H.clone(), engine.Wt.clone(), engine.Vt.clone())
            iteration += 1
        info["objectives"].append(objs)
        info["iters_run"].append(iteration - 1)
    if final is None:
        raise RuntimeError("iNMF_HALS: the objective never decreased, there is no factorisation to write back "
                           "(the reference fails here with an unbound final_H)")
    Hd, Wd, Vd = final
    W_host = Wd.to(torch.float64).cpu().numpy()
    for i in range(num_samples):
        b0, b1 = int(engine.X.own_off[i]), int(engine.X.own_off[i + 1])
        liger_object.adata_list[i].obsm["H"] = Hd[b0:b1, :k].to(torch.float64).cpu().numpy()
        liger_object.adata_list[i].varm["W"] = W_host
        liger_object.adata_list[i].varm["V"] = Vd[i].to(torch.float64).cpu().numpy()
    return info if return_info else None
