#!/usr/bin/env python
"""Device time and HBM fraction of the preprocessing kernels (normalize, scale_not_center) on a synthetic
count matrix: 200k cells x 20k genes, ~300 counts per cell, 3k variable genes kept.  CUDA-event timing."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pyliger_b200 import _lib
    from pyliger_b200.engine import _stream
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=200000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--per-cell", type=int, default=300)
    ap.add_argument("--keep", type=int, default=3000)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    n, g, per = args.cells, args.genes, args.per_cell
    # per-cell sorted unique gene draws (Zipf-like popularity so that the atomics see hot genes)
    pop = (1.0 / torch.arange(1, g + 1, device=dev, dtype=torch.float64)) ** 0.7
    cols = torch.multinomial(pop.expand(2048, g), per, replacement=False, generator=gen)  # 2048 patterns
    pat = torch.randint(0, 2048, (n,), device=dev, generator=gen)
    colidx = torch.sort(cols[pat], dim=1)[0].reshape(-1).to(torch.int32)
    rowptr = torch.arange(0, n + 1, device=dev, dtype=torch.int64) * per
    nnz = colidx.numel()
    peak = 6554.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"cells": n, "genes": g, "nnz": nnz, "kept_genes": args.keep, "peak_gbs": peak}
    for name, dt, es in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
        vals = torch.poisson(torch.full((nnz,), 1.5, device=dev, dtype=torch.float32), generator=gen).add_(1).to(dt)
        norm = torch.empty_like(vals)
        sums = torch.zeros(2, g, dtype=torch.float64, device=dev)
        miss = torch.zeros(1, dtype=torch.int64, device=dev)
        keep = torch.randperm(g, device=dev, generator=gen)[:args.keep].sort()[0]
        colmap = torch.full((g,), -1, dtype=torch.int32, device=dev)
        colmap[keep] = torch.arange(args.keep, dtype=torch.int32, device=dev)
        scaler = torch.rand(args.keep, dtype=torch.float64, device=dev) + 0.5
        counts = torch.zeros(n, dtype=torch.int64, device=dev)

        def timed(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.reps

        t_norm = timed(lambda: _lib.call("inmf_normalize_rows", name, rowptr, colidx, vals, n, norm, sums[0], sums[1],
                                         miss, _stream()))
        t_cnt = timed(lambda: _lib.call("inmf_select_count", "", rowptr, colidx, n, colmap, counts, _stream()))
        new_rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=new_rowptr[1:])
        nnz_new = int(new_rowptr[-1].item())
        nc = torch.empty(nnz_new, dtype=torch.int32, device=dev)
        nv = torch.empty(nnz_new, dtype=dt, device=dev)
        t_sel = timed(lambda: _lib.call("inmf_select_scale", name, rowptr, colidx, norm, n, colmap, scaler, new_rowptr,
                                        nc, nv, _stream()))
        b_norm = nnz * (es + 4 + es) + 16 * n
        b_cnt = nnz * 4 + 24 * n
        b_sel = nnz * (es + 4) + nnz_new * (es + 4) + 24 * n
        out[name] = {
            "normalize_rows": {"ms": t_norm, "algorithmic_bytes": b_norm, "gbs": b_norm / t_norm / 1e6,
                               "frac": b_norm / t_norm / 1e6 / peak},
            "select_count": {"ms": t_cnt, "algorithmic_bytes": b_cnt, "gbs": b_cnt / t_cnt / 1e6,
                             "frac": b_cnt / t_cnt / 1e6 / peak},
            "select_scale": {"ms": t_sel, "algorithmic_bytes": b_sel, "gbs": b_sel / t_sel / 1e6,
                             "frac": b_sel / t_sel / 1e6 / peak, "nnz_kept": nnz_new},
        }
        del vals, norm, nc, nv
    print(json.dumps(out))


if __name__ == "__main__":
    main()
