#!/usr/bin/env python
"""C3 of BASELINE.json: online_iNMF throughput (cell-visits/s of the mini-batch loop, cells/s of the final H
pass) on synthetic data: --datasets x --cells cells x --genes genes, k=40, miniBatch_size=5000.

Launch with torchrun for N > 1 GPUs (X replicated, every rank takes 1/N of each mini-batch)."""
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
    ap.add_argument("--datasets", type=int, default=8)
    ap.add_argument("--cells", type=int, default=125000, help="cells per dataset")
    ap.add_argument("--genes", type=int, default=5000)
    ap.add_argument("--k", type=int, default=40)
    ap.add_argument("--mb", type=int, default=5000)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--cpu-cells", type=int, default=2500, help="cells per dataset for the CPU oracle leg (0 = skip)")
    ap.add_argument("--breakdown", action="store_true")
    args = ap.parse_args()
    import torch
    from pyliger_b200 import _lib
    from pyliger_b200.engine import InmfEngine, _stream
    from pyliger_b200.factorization import _online_iNMF as onl
    from pyliger_b200.synth import make_datasets_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    t0 = time.time()
    mats = make_datasets_device([args.cells] * args.datasets, args.genes, args.k, 0.08, seed0=7000, device=dev)
    g = mats[0].shape[1]
    ns = [m.shape[0] for m in mats]
    if rank == 0:
        print("data %d x %s x %d nnz/cell %.1f (%.1fs)" % (len(mats), ns[0], g, sum(m.nnz for m in mats) / sum(ns),
                                                          time.time() - t0), file=sys.stderr)
    k, lam = args.k, 5.0
    W = onl._init_W(g, k, 1)
    Vs = [onl._init_V_online(ns[i], k, mats[i], 1000, 1) for i in range(len(mats))]
    eng = InmfEngine(mats, k, lam, dtype=dt, device=dev, replicate=True, need_stats=False)
    mbs = np.asarray([np.round((ns[i] / np.sum(ns)) * args.mb).astype(int) for i in range(len(mats))])
    total_iters = int(np.floor(ns[0] * args.epochs / mbs[0]))
    eng.alloc_online(mbs)
    eng.Wt.copy_(onl._to_dev(eng, W))
    for i in range(len(mats)):
        eng.Vt[i].copy_(onl._to_dev(eng, Vs[i]))
    plan = eng.online_plan(total_iters)

    def step(it, epoch_state):
        epoch_prev = epoch_state[0]
        epoch = ((it + 1) * mbs[0]) // ns[0]
        epoch_state[0] = epoch
        scales = [0.0] * len(mats) if it == 0 else ([1.0 / m for m in mbs] if it == 1 else [(it - 1) / it] * len(mats))
        eng.online_step(plan[it], scales, bool(epoch > 0 and epoch - epoch_prev == 1), 1, step=it)

    st = [0]
    for it in range(min(3, total_iters)):  # warm-up (then restart from the same state)
        step(it, st)
    eng.Wt.copy_(onl._to_dev(eng, W))
    for i in range(len(mats)):
        eng.Vt[i].copy_(onl._to_dev(eng, Vs[i]))
    eng.A.zero_(); eng.B.zero_(); eng.A_old.zero_(); eng.B_old.zero_()
    st = [0]
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    l0 = _lib.launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for it in range(total_iters):
        step(it, st)
    e1.record()
    eng.final_H()
    e2.record()
    torch.cuda.synchronize()
    ms_train, ms_final = e0.elapsed_time(e1), e1.elapsed_time(e2)
    if world > 1:
        t = torch.tensor([ms_train, ms_final], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_train, ms_final = t.tolist()
    visits = total_iters * int(mbs.sum())
    res = {"config": "online_iNMF %d x %d cells x %d genes, k=%d, miniBatch_size=%d, epochs=%d, %s, %d GPU(s)" % (
        len(mats), ns[0], g, k, args.mb, args.epochs, args.dtype, world),
        "steps": total_iters, "train_ms": ms_train, "ms_per_step": ms_train / max(total_iters, 1),
        "train_cell_visits_per_s": visits / (ms_train * 1e-3), "final_H_ms": ms_final,
        "final_H_cells_per_s": sum(ns) / (ms_final * 1e-3), "launches": _lib.launch_count() - l0,
        "bpp": eng.bpp_report()["H"]}
    if args.breakdown and rank == 0:
        names = ["prep", "spmm", "bpp", "stats", "update_ab", "hals"]
        acc = dict.fromkeys(names, 0.0)
        reps = 10
        for rep in range(reps):
            it = 5 + rep
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            ev[0].record(); eng.prep()
            ev[1].record()
            if eng.win is not None:  # the tiled window pass the real step uses
                bp, work, counts = eng.win
                eng.H_mb.zero_()
                eng._blocked(bp, work[it], int(counts[it]), eng.Mt, eng.H_mb, eng.ldm, None)
            else:
                eng.rhs_H(seg_desc=plan[it], max_rows=eng.max_sub, out=eng.H_mb)
            ev[2].record(); eng.bpp_H(seg_desc=plan[it], max_rows=eng.max_sub, out=eng.H_mb)
            ev[3].record()
            if eng.win2 is not None:  # the block-aligned window pass the real step uses
                eng.stats_window(it)
            else:
                eng.stats(seg_desc=plan[it], max_rows=eng.max_sub, H=eng.H_mb, St=eng.XHt, HtH=eng.HHt)
            ev[4].record()
            _lib.call("inmf_update_ab", eng.sfx, eng.A, eng.A_old, eng.B, eng.B_old, eng.HHt, eng.XHt, eng.inv_mb,
                      0.5, 0, eng.N, eng.g, eng.k, _stream())
            ev[5].record()
            _lib.call("inmf_hals", eng.sfx, eng.A_all, eng.B_all, eng.Wt, eng.Vt_all, eng.lam, 0, eng.N, eng.g, eng.k, 1,
                      _stream())
            ev[6].record()
            torch.cuda.synchronize()
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1]) / reps
        res["step_breakdown_ms"] = {n: round(v, 4) for n, v in acc.items()}
    if args.cpu_cells and rank == 0 and world == 1:
        from oracle import inmf_oracle as orc
        sample = [m[:args.cpu_cells] for m in mats]
        mb_cpu = max(len(mats), args.mb * args.cpu_cells // ns[0])
        t0 = time.perf_counter()
        orc.online(sample, k, lam, 1, 1, mb_cpu, 1000, 1)
        dtc = time.perf_counter() - t0
        res["cpu_oracle"] = {"cells": args.cpu_cells * len(mats), "seconds": dtc,
                             "cells_per_s_end_to_end": args.cpu_cells * len(mats) / dtc}
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
