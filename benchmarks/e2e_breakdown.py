#!/usr/bin/env python
"""Where the wall time of one public optimize_ALS call goes (host packing, H2D, format build, iterations, D2H)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pyliger_b200 import blocked, liger_from_matrices, optimize_ALS
    from pyliger_b200.engine import DeviceCSR, InmfEngine
    from pyliger_b200.factorization._iNMF_ANLS import _als_init_host
    from pyliger_b200.synth import make_datasets_device

    dev = torch.device("cuda:0")
    mats = make_datasets_device([50000] * 4, 3000, 30, 0.08, seed0=1000, device=dev)
    k, lam = 30, 5.0
    ns = [m.shape[0] for m in mats]
    g = mats[0].shape[1]

    def tic():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep in range(2):
        t0 = tic(); W, V, H = _als_init_host(ns, g, k, 0); t1 = tic()
        X = DeviceCSR(mats, torch.float32, dev); t2 = tic()
        p1 = blocked.build_pass1(X, 32, 4, 148); t3 = tic()
        p2 = blocked.build_pass2(X, 32, 4, 148); t4 = tic()
        eng = InmfEngine(mats, k, lam, dtype=torch.float32, device=dev); t5 = tic()
        eng.set_als_state(W, V, H); t6 = tic()
        eng.initial_objective(); t7 = tic()
        for _ in range(10):
            eng.als_iteration(warm=True).item()
        t8 = tic()
        Hh = eng.gather_H(); Wh = eng.host_W(); Vh = eng.host_V(); t9 = tic()
        print("rep %d: init_host %.1f | DeviceCSR %.1f | pass1 %.1f | pass2 %.1f | engine(total incl. CSR+passes) %.1f | "
              "set_state %.1f | obj0 %.1f | 10 iters %.1f | write-back %.1f ms" % (
                  rep, *(1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6), (t6, t7),
                                                    (t7, t8), (t8, t9)))))
        del eng, X, p1, p2
    lig = liger_from_matrices(mats)
    t0 = tic(); optimize_ALS(lig, k, lam, 1e-6, 10, dtype="float32", progress=False, warm_start=True); t1 = tic()
    print("optimize_ALS total %.1f ms" % (1e3 * (t1 - t0)))


if __name__ == "__main__":
    main()
