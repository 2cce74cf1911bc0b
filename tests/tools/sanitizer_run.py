#!/usr/bin/env python
"""One pass over every kernel family of libinmf_b200.so on small inputs, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tests/tools/sanitizer_run.py
    compute-sanitizer --tool racecheck python tests/tools/sanitizer_run.py

Covers: format builders, sliced passes (one lane per row k <= 32 and k <= 40, four lanes per row k = 50), row-ordered
passes (fp64), mixed-precision passes, half-warp + full-warp BPP (cold, warm, deferral list), V / W solves, prep /
objective, deterministic partial sums, CUDA-graphed iterations, online windows + HALS + update_ab, streamed slots, the
HALS sweep of H, preprocessing, quantile_norm.  Results are checked against the oracle so that a "clean" run is also a
correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oracle import inmf_oracle as orc
    from pyliger_b200 import (blocked, iNMF_HALS, liger_from_matrices, nnlsm_blockpivot, online_iNMF, optimize_ALS,
                              quantile_norm)
    from pyliger_b200.synth import make_datasets
    blocked.SMEM_BUDGET = 96 * 160  # small tiles: several gene and cell blocks
    for k, dtypes in ((12, ("float64", "float32", "mixed")), (36, ("float32",)), (50, ("float32",))):
        mats = make_datasets([400, 300], 200, k, density=0.12, seed0=8000 + k)
        want = orc.als(mats, k, 5.0, 1e-6, 4, rand_seed=1)
        for dtype in dtypes:
            for det in ((False, True) if dtype == "float32" else (False,)):
                lig = liger_from_matrices(mats)
                info = optimize_ALS(lig, k, 5.0, 1e-6, 4, rand_seed=1, dtype=dtype, deterministic=det, return_info=True,
                                    progress=False)
                rel = np.abs(np.array(info["objectives"][0]) - want["objectives"]) / np.array(want["objectives"])
                assert rel.max() < (1e-9 if dtype == "float64" else 2e-3), (k, dtype, det, rel.max())
                print("ALS k=%d %s deterministic=%s: objective rel err %.1e" % (k, dtype, det, rel.max()), flush=True)
    mats = make_datasets([500, 420], 150, 8, density=0.12, seed0=8100)
    for stream in (False, True):
        lig = liger_from_matrices(mats)
        online_iNMF(lig, k=8, max_epochs=2, miniBatch_size=200, verbose=False, stream=stream)
        wo = orc.online(mats, 8, 5.0, 2, 1, 200, 1000, 1)
        err = np.linalg.norm(lig.W - wo["W"]) / np.linalg.norm(wo["W"])
        assert err < 1e-7, err
        print("online stream=%s: W rel err %.1e" % (stream, err), flush=True)
    lig = liger_from_matrices(mats)
    iNMF_HALS(lig, k=8, max_iters=4, verbose=False)
    quantile_norm(lig, min_cells=10, knn_k=10, max_sample=60)
    print("iNMF_HALS + quantile_norm: clusters", np.bincount(lig.adata_list[0].obs["cluster"]).tolist(), flush=True)
    rng = np.random.default_rng(1)
    for kk in (20, 40):  # 70k columns: half-warp solver + deferral list
        M = rng.uniform(0, 2, (kk, 120))
        G = M @ M.T
        R = M @ (rng.gamma(0.3, 1.0, (120, 70000)) * (rng.random((120, 70000)) < 0.1)) - 0.3 * rng.random((kk, 70000))
        X, info = nnlsm_blockpivot(G, R, is_input_prod=True, dtype="float32")
        assert info[0] and (X >= 0).all()
        print("bpp k=%d: %d factorisations, success %s" % (kk, info[2], info[0]), flush=True)
    torch.cuda.synchronize()
    print("SANITIZER RUN OK", flush=True)


if __name__ == "__main__":
    main()
