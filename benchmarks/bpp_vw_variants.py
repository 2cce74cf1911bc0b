#!/usr/bin/env python
"""Times the three batched BPP solves of a C2 / C4-shaped ALS iteration (H, V, W) with the half-warp solver on and off
(warm bit 3) at steady state.  One JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyliger_b200 import _lib  # noqa: E402
from pyliger_b200.engine import InmfEngine, _stream  # noqa: E402
from pyliger_b200.factorization._iNMF_ANLS import _als_init_host  # noqa: E402
from pyliger_b200.synth import make_datasets_resident  # noqa: E402


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    cfgs = {"C2": (4, 50000, 3000, 30), "C4": (8, 31250, 5000, 40)}
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    N, n, g, k = cfgs[name]
    dev = torch.device("cuda:0")
    mats = make_datasets_resident([n] * N, g, k, 0.08, seed0=1000, device=dev)
    eng = InmfEngine(mats, k, 5.0, dtype=torch.float32, device=dev)
    eng.draw_als_state(0)
    eng.initial_objective()
    for i in range(12):
        eng.als_iteration(warm=i > 0)
    torch.cuda.synchronize()
    out = {"config": name}
    R = None
    for off in (0, 8):
        # H: right-hand sides of the current state, solved warm from the current masks (copy back each time)
        eng._begun = False
        eng.als_begin(); eng._begun = False
        R = eng.H.clone()
        m0 = eng.mask_H.clone()

        def solve_h():
            eng.H.copy_(R); eng.mask_H.copy_(m0)
            _lib.call("inmf_bpp", eng.sfx, eng.G, eng.H, eng.ldm, eng.H, eng.ldm, None, 0, eng.mask_H, 1 | 2 | off,
                      eng.seg_full_off, eng.N, eng.max_rows_full, k, eng.bpp_stats[0], *_lib.ws_args(eng.ws_H), _stream())
        tc = timed(lambda: (eng.H.copy_(R), eng.mask_H.copy_(m0)))
        out["H_%s" % ("full_warp" if off else "half_warp")] = timed(solve_h) - tc
        eng.H.copy_(R)
        eng.bpp_H(mask=eng.mask_H, warm=True)
        eng.stats()
        eng.vw_flags = 0
        base = eng._vw
        eng._vw = lambda warm, b=base, o=off: b(warm) | o
        out["V_%s" % ("full_warp" if off else "half_warp")] = timed(lambda: eng.solve_V(warm=True))
        out["W_%s" % ("full_warp" if off else "half_warp")] = timed(lambda: eng.solve_W(warm=True))
        eng._vw = base
    eng.bpp_stats.zero_()
    eng.solve_V(warm=True)
    out["V_report"] = eng.bpp_report()["V"]
    eng.bpp_stats.zero_()
    eng.solve_W(warm=True)
    out["W_report"] = eng.bpp_report()["W"]
    out["W_passive_mean"] = float((eng.Wt > 0).sum().item()) / eng.Wt.shape[0]
    out["V_passive_mean"] = float((eng.Vt > 0).sum().item()) / (eng.Vt.shape[0] * eng.Vt.shape[1])
    # latency of the W solve against the number of columns (is it one column's latency?)
    import types
    for frac in (1.0, 0.25):
        gsub = int(eng.g * frac)
        seg = torch.tensor([0, gsub], dtype=torch.int64, device=dev)
        def solve_sub():
            _lib.call("inmf_bpp", eng.sfx, eng.Gw, eng.Rw, k, eng.Wt, k, None, 0, eng.mask_W, eng._vw(True), seg, 1, gsub, k,
                      eng.bpp_stats[2], *_lib.ws_args(eng.ws_W), _stream())
        out["W_bpp_only_%d_cols" % gsub] = timed(solve_sub)
    out["rhs_w_only"] = timed(lambda: _lib.call("inmf_rhs_w", eng.sfx, eng.St, eng.Vt, eng.HtH, eng.Rw, eng.Gw, eng.N, eng.g, k, _stream()))
    out["rhs_v_only"] = timed(lambda: _lib.call("inmf_rhs_v", eng.sfx, eng.St, eng.Wt, eng.HtH, eng.Rv, eng.Gv, eng.lam, eng.N, eng.g, k, _stream()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
