#!/usr/bin/env python
"""bench.py -- optimize_ALS cell-iterations/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--dtype f32|f64]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (N > 1; one rank per GPU, NCCL)

Workload (configs[1] of BASELINE.json, "C2"): 4 synthetic datasets x 50 000 cells x 3 000 genes, density
0.08, k = 30, lambda = 5, per GPU (weak scaling: every rank holds its own 4 x 50k cells; W, V and the
k x k / genes x k statistics are replicated and all-reduced once per iteration).  A "step" is one ALS
iteration (H, V, W block update + objective) over all resident cells.

One JSON line on stdout from rank 0.  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(datasets=4, cells_per_dataset=50000, genes=3000, k=30, density=0.08, value_lambda=5.0)
# configs[3]: 8 x 250k cells, cell-sharded over the ranks (strong scaling: cells per rank = 250k / world)
C4 = dict(datasets=8, cells_per_dataset=250000, genes=5000, k=40, density=0.08, value_lambda=5.0)
CPU_SAMPLE_CELLS = 5000  # per dataset, for the CPU legs


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(mats):
    return [m[:CPU_SAMPLE_CELLS] for m in mats]


def time_oracle(sample, cfg, steps, warmup):
    """CPU restatement of the reference's optimize_ALS (oracle/inmf_oracle.py) timed per iteration."""
    from oracle import inmf_oracle as orc
    marks = [time.perf_counter()]
    orc.als(sample, cfg["k"], cfg["value_lambda"], 0.0, warmup + steps, rand_seed=1, dense_objective=False,
            callback=lambda it, obj, H, W, V: marks.append(time.perf_counter()))
    dt = marks[-1] - marks[warmup]
    cells = sum(m.shape[0] for m in sample)
    return cells * steps / dt, dt / steps


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (numpy port = the oracle; the
    reference is pure Python and does not travel to the GPU box) on a bounded sample of the workload."""
    if rank != 0:
        return
    from pyliger_b200.synth import make_datasets
    cfg = C2 if getattr(args, "config", "C2") == "C2" else C4
    t0 = time.time()
    sample = make_datasets([CPU_SAMPLE_CELLS] * cfg["datasets"], cfg["genes"], cfg["k"], cfg["density"],
                           structured=False, seed0=1000)
    log("reference arm: sample generated in %.1fs" % (time.time() - t0))
    value, s_per_step = time_oracle(sample, cfg, args.steps, args.warmup)
    desc = "%d datasets x %d cells x %d genes (1/%d of the cells of the GPU workload), k=%d" % (
        cfg["datasets"], CPU_SAMPLE_CELLS, cfg["genes"], cfg["cells_per_dataset"] // CPU_SAMPLE_CELLS, cfg["k"])
    out = {
        "impl": "reference", "metric": "optimize_ALS cell-iterations/s", "value": value, "unit": "cell-iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s optimize_ALS %d datasets x %dk cells x %dk genes, k=%d, lambda=%g (CPU: bounded sample)" % (
            getattr(args, "config", "C2"), cfg["datasets"], cfg["cells_per_dataset"] // 1000, cfg["genes"] // 1000, cfg["k"],
            cfg["value_lambda"]), "sample": desc},
        "cpu_baseline": {"value": value, "unit": "cell-iters/s", "cores": cpu_threads(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "cell-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--warm-start", type=int, default=1)
    ap.add_argument("--config", default="C2", choices=["C2", "C4"],
                    help="C2 (default, the headline: weak scaling) or C4 (8 x 250k cells sharded over the ranks)")
    ap.add_argument("--cells", type=int, default=0, help="cells per dataset per GPU (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        log("note: raising warmup to 3 (timing rules)")
        args.warmup = 3

    import torch
    from pyliger_b200 import _lib, liger_from_matrices, optimize_ALS
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.factorization._iNMF_ANLS import _als_init_host
    from pyliger_b200.synth import make_datasets_device

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    base = C2 if args.config == "C2" else C4
    strong = args.config == "C4"
    cfg = dict(base, cells_per_dataset=args.cells or (base["cells_per_dataset"] // world if strong
                                                      else base["cells_per_dataset"]))
    k, lam, g = cfg["k"], cfg["value_lambda"], cfg["genes"]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    esize = 4 if args.dtype == "f32" else 8

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        return float(t.item())

    # ------------------------------------------------------------------ data (per rank: its own cells)
    t0 = time.time()
    mats = make_datasets_device([cfg["cells_per_dataset"]] * cfg["datasets"], g, k, cfg["density"],
                                seed0=1000 + 100 * rank, device=dev)
    g = mats[0].shape[1]
    if world > 1:  # the (rare) all-zero-gene filter must agree across ranks
        gs = torch.tensor([g], device=dev)
        torch.distributed.all_reduce(gs, op=torch.distributed.ReduceOp.MIN)
        g = int(gs.item())
        mats = [m[:, :g].tocsr() for m in mats]
    ns = [m.shape[0] for m in mats]
    cells_local = sum(ns)
    nnz_local = sum(m.nnz for m in mats)
    log("rank %d: data %s nnz/cell %.1f in %.1fs" % (rank, [m.shape for m in mats], nnz_local / cells_local,
                                                    time.time() - t0))

    # ------------------------------------------------------------------ device-resident loop -> value
    eng = InmfEngine(mats, k, lam, dtype=dtype, device=dev, presharded=world > 1)
    W0, V0, H0 = _als_init_host(ns, g, k, 0, h_stream=rank if world > 1 else None)
    eng.set_als_state(W0, V0, H0)
    obj = eng.initial_objective()
    objs = [obj]
    done = 0
    # als_iteration_value(speculate=True) is what optimize_ALS does: the next iteration's pass 1 is enqueued before
    # the host waits for the objective.  Never across the boundaries of the timed region: all work of the K timed
    # iterations (and nothing else) lies between the two events.
    for i in range(args.warmup):
        objs.append(eng.als_iteration_value(warm=bool(args.warm_start) and done > 0, speculate=i + 1 < args.warmup))
        done += 1
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        objs.append(eng.als_iteration_value(warm=bool(args.warm_start), speculate=i + 1 < args.steps))
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    cells_total = sum_over_ranks(cells_local)
    value = cells_total * args.steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps
    mono = bool(np.all(np.diff(objs) <= 1e-6 * abs(objs[0])))
    log("rank %d: objectives first/last %.6e %.6e monotone=%s bpp=%s" % (rank, objs[1], objs[-1], mono,
                                                                       eng.bpp_report()["H"]))

    # ------------------------------------------------------------------ per-kernel CUDA-event breakdown
    phases = ["spmm", "bpp_H", "stats", "solve_V", "solve_W", "prep", "objective"]
    acc = {p: 0.0 for p in phases}
    reps = 5
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 1)]
        evs[0].record()
        eng.rhs_H()
        evs[1].record()
        eng.bpp_H(mask=eng.mask_H, warm=bool(args.warm_start))
        evs[2].record()
        eng.stats()
        evs[3].record()
        eng.solve_V(warm=bool(args.warm_start))
        evs[4].record()
        eng.solve_W(warm=bool(args.warm_start))
        evs[5].record()
        eng.prep()
        evs[6].record()
        eng.objective()
        evs[7].record()
        torch.cuda.synchronize()
        for i, p in enumerate(phases):
            acc[p] += evs[i].elapsed_time(evs[i + 1]) / reps
    peak, peak_src = measured_peak_gbs()
    nnz_c = nnz_local / cells_local
    b_x = nnz_c * (esize + 4) + 4  # one streaming pass over a cell's CSR row
    alg = {  # algorithmic bytes per cell for each kernel that touches cells
        "spmm": b_x + k * esize, "bpp_H": 2 * k * esize, "stats": b_x + k * esize,
    }
    b_als = 2 * b_x + 2 * k * esize
    kern = {}
    for p in phases:
        e = {"ms": round(acc[p], 4)}
        if p in alg:
            e["algorithmic_bytes"] = alg[p] * cells_local
            e["achieved_gbs"] = alg[p] * cells_local / (acc[p] * 1e-3) / 1e9
            e["frac"] = e["achieved_gbs"] / peak
        kern[p] = e
    dom = max(alg, key=lambda p: acc[p])
    traffic = None
    try:  # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        if args.dtype == "f32" and world >= 1:
            traffic = tr["kernels"][dom]["dram_bytes_per_launch"]
            for p_, e_ in kern.items():
                if p_ in tr["kernels"]:
                    e_["ncu_dram_bytes_per_launch"] = tr["kernels"][p_]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "kernels": kern,
                "step": {"algorithmic_bytes_per_cell_iter": b_als,
                         "achieved_gbs": b_als * cells_local / (ms_per_step * 1e-3) / 1e9,
                         "frac": b_als * cells_local / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ------------------------------------------------------------------ end to end through the public API
    e2e = None
    if not args.no_e2e:
        del eng
        torch.cuda.empty_cache()
        lig = liger_from_matrices(mats)
        optimize_ALS(lig, k, lam, 1e-6, 2, dtype=args.dtype.replace("f", "float"), presharded=world > 1,
                     progress=False, warm_start=bool(args.warm_start))  # warm-up of the whole call path
        lig = liger_from_matrices(mats)
        barrier()
        t0 = time.perf_counter()
        info = optimize_ALS(lig, k, lam, 1e-6, args.steps, dtype=args.dtype.replace("f", "float"),
                            presharded=world > 1, progress=False, warm_start=bool(args.warm_start), return_info=True)
        torch.cuda.synchronize()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        iters = info["iters_run"][0]
        h2d, d2h = info["h2d_bytes"], info["d2h_bytes"]  # counted from the arrays actually copied
        e2e = {"value": cells_total * iters / t_e2e, "unit": "cell-iters/s", "h2d_bytes_per_step": h2d / iters,
               "d2h_bytes_per_step": d2h / iters, "seconds": t_e2e, "iters": iters,
               "what": "wall time of pyliger_b200.optimize_ALS(liger, k, max_iters=steps) on host scipy CSR "
                       "inputs: CSR pack + H2D + init + all iterations + H/W/V write-back to host float64"}

    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed loop, the breakdown and the e2e call

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = cpu_sample(mats)
        v, s_step = time_oracle(sample, cfg, 2, 1)
        cpu = {"value": v, "unit": "cell-iters/s", "cores": cpu_threads(), "kind": "port",
               "sample": "first %d cells of each of the %d datasets (1/%d of the workload), 2 timed ALS iterations "
                         "after 1 warm-up, numpy oracle (single Python thread + BLAS threads)" % (
                             CPU_SAMPLE_CELLS, len(mats), cfg["cells_per_dataset"] // CPU_SAMPLE_CELLS)}

    if rank == 0:
        out = {
            "metric": "optimize_ALS cell-iterations/s", "value": value, "unit": "cell-iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.config + " optimize_ALS: %d datasets x %d cells x %d genes per GPU, density %.2f, k=%d, "
                                   "lambda=%g" % (len(mats), cfg["cells_per_dataset"], g, cfg["density"], k, lam),
                       "cells_total": cells_total, "nnz_per_cell": nnz_c, "warm_start": bool(args.warm_start),
                       "l2": "inputs larger than L2 (CSR %.0f MB per GPU streamed twice per step)" % (
                           nnz_local * (esize + 4) / 1e6),
                       "parallelism": "cells sharded over %d GPU(s), all-reduce of k x k and genes x k statistics" % world},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "objective_monotone": mono,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
