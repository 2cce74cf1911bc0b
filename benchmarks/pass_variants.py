#!/usr/bin/env python
"""Times the two X-streaming passes (and the H solve) of one ALS iteration for the kernel variants of
pyliger_b200.blocked (sliced SELL-8 vs row-ordered wide / split) on a BASELINE-shaped workload and checks
that they agree.  One JSON line per (config, variant) on stdout.

    python benchmarks/pass_variants.py [--config C2|C4] [--cells N] [--reps R]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pyliger_b200 import blocked  # noqa: E402
from pyliger_b200.engine import InmfEngine  # noqa: E402
from pyliger_b200.factorization._iNMF_ANLS import _als_init_host  # noqa: E402
from pyliger_b200.synth import make_datasets_device  # noqa: E402

CONFIGS = {"C2": dict(datasets=4, cells=50000, genes=3000, k=30), "C4": dict(datasets=8, cells=31250, genes=5000, k=40),
           "C1": dict(datasets=2, cells=7000, genes=2000, k=20)}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="sell,sell8,rows")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.cells:
        cfg["cells"] = args.cells
    if args.k:
        cfg["k"] = args.k
    k, g = cfg["k"], cfg["genes"]
    dev = torch.device("cuda:0")
    mats = make_datasets_device([cfg["cells"]] * cfg["datasets"], g, k, 0.08, seed0=1000, device=dev)
    g = mats[0].shape[1]
    ns = [m.shape[0] for m in mats]
    nnz = sum(m.nnz for m in mats)
    cells = sum(ns)
    W0, V0, H0 = _als_init_host(ns, g, k, 0)
    peak = 6538.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ref = None
    for variant in args.variants.split(","):
        blocked.SELL = variant.startswith("sell") or variant == "mixed"
        blocked.SELL_HEIGHT_KP32 = 8 if variant == "sell8" else 32
        blocked.SELL_HEIGHT_KP40 = 8 if variant == "sell8" else 33
        eng = InmfEngine(mats, k, 5.0, dtype="mixed" if variant == "mixed" else torch.float32, device=dev)
        eng.set_als_state(W0, V0, H0)
        eng.initial_objective()
        for i in range(4):
            eng.als_iteration(warm=i > 0)
        torch.cuda.synchronize()
        Hsol = eng.H.clone()
        t1 = timed(eng.rhs_H, args.reps)
        R = eng.H.clone()
        eng.H.copy_(Hsol)
        t2 = timed(eng.stats, args.reps)
        St, HtH = eng.St.clone(), eng.HtH.clone()

        def solve():
            eng.H.copy_(R)
            eng.bpp_H(mask=eng.mask_H, warm=True)
        tb = timed(solve, args.reps)
        tcopy = timed(lambda: eng.H.copy_(R), args.reps)
        bytes_pass = nnz * 8 + cells * 4 + cells * k * 4
        out = {"config": args.config, "variant": variant, "cells": cells, "genes": g, "k": k, "ldm": eng.ldm,
               "nnz_per_cell": nnz / cells,
               "pass1_ms": t1[0], "pass1_min_ms": t1[1], "pass2_ms": t2[0], "pass2_min_ms": t2[1],
               "bpp_H_ms": tb[0] - tcopy[0],
               "pass1_frac": bytes_pass / (t1[0] * 1e-3) / 1e9 / peak, "pass2_frac": bytes_pass / (t2[0] * 1e-3) / 1e9 / peak,
               "clk_per_nnz_pass1": t1[0] * 1e-3 * 1.965e9 * 148 / nnz, "clk_per_nnz_pass2": t2[0] * 1e-3 * 1.965e9 * 148 / nnz}
        if variant == "mixed":  # the pieces of the mixed passes on their own
            from pyliger_b200 import _lib
            from pyliger_b200.engine import _stream
            p1, p2 = eng.pass1, eng.pass2
            out["to_bf16_Mt_ms"] = timed(lambda: _lib.call("inmf_to_bf16_rows", "", eng.Mt, eng.ldm, eng.N * eng.g, k,
                                                          eng.Mt16, _stream()), args.reps)[0]
            out["to_bf16_H_ms"] = timed(lambda: _lib.call("inmf_to_bf16_rows", "", eng.H, eng.ldm, eng.H.shape[0], k,
                                                         eng.H16, _stream()), args.reps)[0]
            out["kernel_pass1_ms"] = timed(lambda: _lib.call("inmf_sell_pass_mixed", "", p1.work, p1.nwork, p1.slice_ptr,
                                                            p1.rowid, p1.pairs4, eng.Mt16, R, eng.ldm, k,
                                                            p1.max_tile_rows, _stream()), args.reps)[0]
            out["kernel_pass2_ms"] = timed(lambda: _lib.call("inmf_sell_pass_mixed", "", p2.work, p2.nwork, p2.slice_ptr,
                                                            p2.rowid, p2.pairs4, eng.H16, eng.St, k, k,
                                                            p2.max_tile_rows, _stream()), args.reps)[0]
            out["to_bf16_gram_ms"] = timed(lambda: _lib.call("inmf_to_bf16_gram", "", eng.H, eng.ldm, eng.seg_full_off,
                                                            eng.N, eng.max_rows_full, k, eng.H16, eng.HtH, _stream()),
                                           args.reps)[0]
            out["memset_St_ms"] = timed(lambda: (eng.St.zero_(), eng.HtH.zero_()), args.reps)[0]
        if isinstance(eng.pass1, blocked.SellPass):
            out["pad1"] = eng.pass1.padded / max(1, eng.pass1.nnz)
            out["pad2"] = eng.pass2.padded / max(1, eng.pass2.nnz)
            out["nwork"] = [eng.pass1.nwork, eng.pass2.nwork]
        if ref is None:
            ref = (R, St, HtH)
        else:
            rel = lambda a, b: float((a - b).norm() / b.norm())
            out["rel_vs_first"] = [rel(R, ref[0]), rel(St, ref[1]), rel(HtH, ref[2])]
        print(json.dumps(out), flush=True)
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
