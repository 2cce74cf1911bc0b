"""optimize_ALS on B200: same signature, write-back and errors as pyliger
(/root/reference/src/pyliger/factorization/_iNMF_ANLS.py:8-196), CUDA kernels underneath.

Keyword-only extensions (defaults preserve reference behaviour):
  dtype       "float64" (default, parity mode), "float32" (throughput mode) or "mixed": float32 with the dense
              operand of the two X-streaming passes (W + V_i, H) rounded to bf16 in shared memory -- halves the
              operand traffic that bounds those kernels; X, the accumulation, the solves and the statistics stay
              fp32.  k <= 32.  A labelled reduced-precision mode (objective within the 1e-3 fp32 tolerance, see
              tests/test_gpu_parity.py::test_c1_mixed_precision_mode), never what "float32" means
  device      torch device (default: current CUDA device)
  warm_start  start each BPP solve from the previous iteration's passive sets -- the reference solver's own
              ``init`` argument (_utilities.py:222-225); NNLS with an SPD Gram matrix has a unique solution, so the
              iterates are the same (tests: trajectories equal to 4e-16 in fp64 with and without) while the
              solver needs ~2 instead of 4-5 factorisations per column.  Default True; False = the reference's
              literal cold start (init=None) in every iteration
  deterministic  (float32) bit-reproducible iterations: the sliced passes store per-tile partial sums that are added in
              a fixed order instead of with floating-point atomics, and the Gram / objective reductions use one CTA
              each.  Slower (about +10 % at C2); default False
  return_info if True, return a dict with the per-iteration objectives and solver counters
              (the reference returns None; the default stays None)
  tensor_core None (default) = CUDA-core passes; "tf32x3" / "tf32" = experimental dense-tile tcgen05 form
              of the two X-streaming passes (float32 only; "tf32" is reduced precision, ~1e-3 per product)
  presharded  multi-GPU only (one process per GPU under torchrun): if True every process passes ONLY
              its own cells of each dataset (H is written back for those cells); if False every process
              passes the full datasets and works on its 1/world row slice, and the full H is gathered.

Multi-GPU: when torch.distributed is initialised the cells are sharded over the ranks and the k x k
and genes x k statistics are all-reduced once per iteration.
"""

import numpy as np
import torch
from tqdm import tqdm

from .. import devcache
from ..engine import InmfEngine, launch_als_draw, require_cuda, resolve_dtype


def _als_init_host(ns, num_genes, k, seed, h_stream=None):
    """Seeded init in the reference's draw order and RNG (_iNMF_ANLS.py:103-108).
    ``h_stream``: pre-sharded multi-GPU runs draw their local H from a per-rank stream (W and V must be
    identical on every rank; H only enters the initial objective, App. B-17)."""
    np.random.seed(seed=seed)
    W = np.abs(np.random.uniform(0, 2, (k, num_genes)))
    V = [np.abs(np.random.uniform(0, 2, (k, num_genes))) for _ in ns]
    if h_stream is not None:
        np.random.seed(seed=(seed + 7919 * (h_stream + 1)) % (2 ** 32))
    H = [np.abs(np.random.uniform(0, 2, (n, k))) for n in ns]
    return W, V, H


def optimize_ALS(
    liger_object,
    k,
    value_lambda=5.0,
    thresh=1e-6,
    max_iters=30,
    nrep=1,
    H_init=None,
    W_init=None,
    V_init=None,
    rand_seed=1,
    print_obj=False,
    *,
    dtype="float64",
    device=None,
    warm_start=True,
    return_info=False,
    progress=True,
    presharded=False,
    tensor_core=None,
    deterministic=False,
):
    N = liger_object.num_samples
    ns = [adata.shape[0] for adata in liger_object.adata_list]
    num_genes = len(liger_object.var_genes)
    X = [adata.layers["scale_data"] for adata in liger_object.adata_list]

    if k >= np.min(ns):
        raise ValueError(
            "Select k lower than the number of cells in smallest dataset: {}".format(np.min(ns)))
    if k >= num_genes:
        raise ValueError(
            "Select k lower than the number of variable genes: {}".format(len(liger_object.var_genes)))

    # The first repetition's seeded init (numpy's legacy RNG stream, reproduced on the device) is generated
    # on a side stream while the main thread uploads X and builds the device formats.
    import torch.distributed as _dist
    _rank = _dist.get_rank() if (_dist.is_available() and _dist.is_initialized()) else 0
    _world = _dist.get_world_size() if (_dist.is_available() and _dist.is_initialized()) else 1
    h_stream = _rank if (presharded and _world > 1) else None
    if tensor_core not in (None, "tf32x3", "tf32"):
        raise ValueError("tensor_core must be None, 'tf32x3' or 'tf32'")
    user_init = not (W_init is None and V_init is None and H_init is None)
    pending = None
    if not user_init:  # the first seeded draw runs on one SM while X is uploaded and laid out
        pending = launch_als_draw(require_cuda(device), ns, num_genes, k, rand_seed - 1, h_stream)
    engine = InmfEngine(X, k, value_lambda, dtype=dtype, device=device, presharded=presharded,
                        use_tc=tensor_core is not None, tc_three_pass=tensor_core != "tf32",
                        deterministic=deterministic)
    best_obj = np.inf
    best = None
    best_dev = None
    info = {"objectives": [], "bpp": None, "iters_run": [], "h2d_bytes": engine.X.h2d_bytes}
    for j in range(nrep):
        if not user_init:  # the seeded draw, generated on the device (same bits as numpy's stream)
            engine.draw_als_state(rand_seed + j - 1, h_stream=h_stream, pending=pending if j == 0 else None)
        else:
            W, V, H = _als_init_host(ns, num_genes, k, rand_seed + j - 1, h_stream=h_stream)
            if W_init is not None:
                W = W_init
            if V_init is not None:
                V = V_init
            if H_init is not None:
                H = H_init
            engine.set_als_state(W, V, H)
        info["h2d_bytes"] += engine.h2d_state_bytes
        delta = 1
        obj0 = engine.initial_objective()
        objs = [obj0]
        iters_run = 0
        it_range = tqdm(range(max_iters)) if progress else range(max_iters)
        for it in it_range:
            if delta > thresh:
                obj_prev = obj0
                obj0 = engine.als_iteration_value(warm=warm_start and iters_run > 0, speculate=it + 1 < max_iters)
                objs.append(obj0)
                delta = np.absolute(obj_prev - obj0) / ((obj_prev + obj0) / 2)
                iters_run += 1
                if not delta > thresh:
                    engine.als_cancel()  # converged: the speculative pass 1 of the next iteration is dropped
            else:
                continue
        info["objectives"].append(objs)
        info["iters_run"].append(iters_run)
        info["d2h_bytes"] = info.get("d2h_bytes", 0) + 8 * len(objs)
        if obj0 < best_obj:
            best = engine.host_factors()
            best_dev = engine.device_H_views(clone=nrep > 1)
            esz = 4 if engine.dtype == torch.float32 else 8  # the factors cross the bus in the compute dtype
            info["d2h_bytes"] += esz * (sum(h.size for h in best[0]) + best[1].size + sum(v.size for v in best[2]))
            best_obj = obj0
        if print_obj:
            print("Objective: {}".format(best_obj))
    info["bpp"] = engine.bpp_report()
    info["fused_reduce"] = engine.mc is not None  # pass 2 fused with its all-reduce (multi-GPU, NVLink multicast)
    if engine.fused_reduce_error:
        info["fused_reduce_error"] = engine.fused_reduce_error
    info["best_objective"] = best_obj

    final_H, final_Wt, final_Vt = best
    for i in range(N):
        if best_dev is not None:  # the stages that consume H (quantile_norm) find it on the device
            devcache.register_dense(final_H[i], engine.device, best_dev[i])
        liger_object.adata_list[i].obsm["H"] = final_H[i]
        # the reference stores transposed views of its k x g matrices (:184-185); same values here
        liger_object.adata_list[i].varm["W"] = final_Wt
        liger_object.adata_list[i].varm["V"] = final_Vt[i]
    return info if return_info else None
