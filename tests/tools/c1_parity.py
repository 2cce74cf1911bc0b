#!/usr/bin/env python
"""C1 of BASELINE.json: optimize_ALS on 2 datasets x 7 000 cells x 2 000 genes, k=20, lambda=5, 30 iterations,
rand_seed=1 -- full trajectory parity of the CUDA path (fp64 and fp32, cold and warm start) against the numpy
oracle (the pinned restatement of the reference), with both timed on the same box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oracle import inmf_oracle as orc
    from pyliger_b200 import liger_from_matrices, optimize_ALS
    from pyliger_b200.synth import make_datasets

    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    mats = make_datasets([7000, 7000], 2000, 20, density=0.10, seed0=1000)
    k, lam = 20, 5.0
    t0 = time.perf_counter()
    want = orc.als(mats, k, lam, 1e-6, iters, rand_seed=1, dense_objective=False)
    t_cpu = time.perf_counter() - t0
    cells = sum(m.shape[0] for m in mats)
    out = {"config": "C1 optimize_ALS %s x %d genes, k=%d, lambda=%g, %d iterations, rand_seed=1" % (
        [m.shape[0] for m in mats], mats[0].shape[1], k, lam, iters),
        "oracle_cpu": {"seconds": t_cpu, "cell_iters_per_s": cells * want["iters_run"] / t_cpu,
                       "final_objective": want["objectives"][-1], "threads": os.cpu_count()}}
    for dtype in ("float64", "float32"):
        for warm in (False, True):
            lig = liger_from_matrices(mats)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = optimize_ALS(lig, k, lam, 1e-6, iters, rand_seed=1, dtype=dtype, warm_start=warm, return_info=True,
                                progress=False)
            torch.cuda.synchronize()
            t_gpu = time.perf_counter() - t0
            objs = np.array(info["objectives"][0])
            rel = np.abs(objs - np.array(want["objectives"])) / np.array(want["objectives"])
            herr = [float(np.linalg.norm(lig.H[i] - want["H"][i]) / np.linalg.norm(want["H"][i])) for i in range(2)]
            agree = [float(np.mean(np.all((lig.H[i] > 0) == (want["H"][i] > 0), axis=1))) for i in range(2)]
            werr = float(np.linalg.norm(lig.W.T - want["W"]) / np.linalg.norm(want["W"]))
            verr = [float(np.linalg.norm(lig.V[i].T - want["V"][i]) / np.linalg.norm(want["V"][i])) for i in range(2)]
            out["%s%s" % (dtype, "_warm" if warm else "")] = {
                "max_rel_objective_error": float(rel.max()), "H_rel_fro": herr, "W_rel_fro": werr, "V_rel_fro": verr,
                "H_active_set_agreement_per_cell": agree, "seconds_end_to_end": t_gpu,
                "cell_iters_per_s_end_to_end": cells * info["iters_run"][0] / t_gpu, "bpp_H": info["bpp"]["H"]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
