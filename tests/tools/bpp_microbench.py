#!/usr/bin/env python
"""C5 of BASELINE.json: batched nnlsm_blockpivot microbenchmark (columns/s) on one B200.

H-update-shaped problems: G = (W+V)(W+V)' + lambda V V', R = (W+V) X' from the synthetic generator,
``--cols`` columns (default 1M; BASELINE names 10M), k in {20, 40, 60}.  Reports cold-start and
warm-start (start set = the solution's own passive set) columns/s and the solver counters, next to the
numpy oracle on a bounded number of columns.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cols", type=int, default=1000000)
    ap.add_argument("--ks", default="20,40,60")
    ap.add_argument("--genes", type=int, default=5000)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-cols", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ginv", action="store_true", help="cold start without the G^-1 full-set shortcut")
    args = ap.parse_args()
    import torch
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    from pyliger_b200.synth import make_datasets_device

    dev = torch.device("cuda:0")
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    esize = 4 if args.dtype == "f32" else 8
    base_cells = min(args.cols, 100000)
    X = make_datasets_device([base_cells], args.genes, 40, 0.08, seed0=5000, device=dev)[0]
    g = X.shape[1]
    Xd = torch.sparse_csr_tensor(torch.from_numpy(X.indptr.astype(np.int64)), torch.from_numpy(X.indices.astype(np.int64)),
                                 torch.from_numpy(X.data), size=X.shape).to(dev)
    for k in [int(s) for s in args.ks.split(",")]:
        rng = np.random.default_rng(k)
        W = np.abs(rng.uniform(0, 2, (k, g)))
        V = np.abs(rng.uniform(0, 2, (k, g)))
        M = W + V
        G = M @ M.T + 5.0 * V @ V.T
        Rt_base = torch.sparse.mm(Xd, torch.from_numpy(M.T.copy()).to(dev))  # cells x k, float64
        reps_needed = (args.cols + base_cells - 1) // base_cells
        # replicate with a small multiplicative jitter so columns are not identical
        parts = []
        gen = torch.Generator(device=dev)
        gen.manual_seed(k)
        for r in range(reps_needed):
            jit = 1.0 + 0.05 * torch.randn(Rt_base.shape, device=dev, dtype=torch.float64, generator=gen)
            parts.append(Rt_base * jit)
        Rt = torch.cat(parts, 0)[:args.cols].to(dt).contiguous()
        Gd = torch.from_numpy(G).to(dev)
        Xo = torch.empty_like(Rt)
        mask = torch.zeros(args.cols, dtype=torch.int64, device=dev)
        seg = torch.tensor([0, args.cols], dtype=torch.int64, device=dev)
        ginv = _lib.bpp_workspace(args.cols, 1, k, dev)
        res = {"k": k, "cols": args.cols, "dtype": args.dtype}
        for mode, warm in (("cold", 0), ("warm", 1)):
            times = []
            for rep in range(args.reps + 1):
                stats = torch.zeros(8, dtype=torch.int64, device=dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.call("inmf_bpp", args.dtype, Gd, Rt, k, Xo, k, None, 0, mask, warm, seg, 1, args.cols, k, stats,
                          *_lib.ws_args(None if args.no_ginv else ginv), _stream())
                e1.record()
                torch.cuda.synchronize()
                if rep > 0:
                    times.append(e0.elapsed_time(e1))
            st = stats.cpu().numpy()
            ms = float(np.mean(times))
            res[mode] = {"ms": ms, "cols_per_s": args.cols / (ms * 1e-3), "rounds_per_col": st[4] / st[6],
                         "fact_per_col": st[1] / st[6], "not_converged": int(st[0]), "backup": int(st[2]),
                         "rounds_max": int(st[5]), "hbm_gbs": 2 * k * esize * args.cols / (ms * 1e-3) / 1e9}
        res["passive_fraction"] = float((Xo > 0).double().mean().item())
        if not args.no_cpu:
            from oracle import inmf_oracle as orc
            n = min(args.cpu_cols, args.cols)
            Rh = Rt[:n].double().cpu().numpy().T.copy()
            t0 = time.perf_counter()
            Xc, info = orc.bpp(G, Rh)
            dtc = time.perf_counter() - t0
            err = np.linalg.norm(Xo[:n].double().cpu().numpy().T - Xc) / np.linalg.norm(Xc)
            res["cpu_oracle"] = {"cols": n, "cols_per_s": n / dtc, "rel_err_vs_gpu": float(err)}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
