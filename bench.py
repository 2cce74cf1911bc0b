#!/usr/bin/env python
"""bench.py -- optimize_ALS cell-iterations/s on B200 (BASELINE.json metric) + the other named configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--dtype f32|f64]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (N > 1; one rank per GPU, NCCL)

Headline (`value`, `e2e`, `roofline`): configs[1] of BASELINE.json, "C2": 4 synthetic datasets x 50 000 cells x 3 000
genes, density 0.08, k = 30, lambda = 5, per GPU (weak scaling: every rank holds its own 4 x 50k cells; W, V and the
k x k / genes x k statistics are replicated and all-reduced once per iteration).  A "step" is one ALS iteration (H, V,
W block update + objective) over all resident cells, with the API-default algorithm options of
pyliger_b200.optimize_ALS (warm-started BPP; fp32 is the throughput mode, fp64 the parity mode -- both reported).

`extra` carries the other configs of BASELINE.json, each device-timed with CUDA events, max over ranks:
  c2_cold_f32   the same C2 loop with cold-started solves (the reference's literal init = None)
  c2_mixed      the same C2 loop with dtype="mixed" (bf16 tile in the two passes; labelled reduced precision)
  c2_warm_f64   the same C2 loop in fp64 (parity mode)
  c4_strong     configs[3]: 8 x 250k cells x 5k genes, k = 40, the 2 M cells sharded over the ranks (STRONG scaling)
  c3_online     configs[2]: online_iNMF, 1 M cells (8 x 125k) in 5 000-cell mini-batches, k = 40: mini-batch loop
                (cell-visits/s) and the final H pass (cells/s, HBM fraction)
  c5_bpp        configs[4]: batched nnlsm_blockpivot, 10 M columns, k in {20, 40, 60}, cold and warm (N = 1 only)
  parity_vs_1gpu (N > 1) the row-sharded N-rank run against the 1-rank run of the same small problem, same process
  c2_fused_reduce / c4_strong.fused_reduce_leg (N > 1) the same loops with the OTHER reduction mode: pass 2 fused with its
                all-reduce (multimem.red into multicast-bound sum buffers) when the headline used NCCL, and vice versa

Launch policy: the engine default (`--graph -1`): eager launches on one GPU, CUDA-graphed warm iterations at N > 1;
`--graph 1` / `--graph 0` force either.

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
# configs[2]: 1 M cells streamed in 5k-cell mini-batches
C3 = dict(datasets=8, cells_per_dataset=125000, genes=5000, k=40, density=0.08, value_lambda=5.0, minibatch=5000)
C5 = dict(columns=10_000_000, ks=(20, 40, 60), genes=5000, base_cells=100000)
CPU_SAMPLE_CELLS = 5000  # per dataset, for the CPU legs
PROFILE_TRAFFIC = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")


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


# ---------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ---------------------------------------------------------------------------------------------------------
def cpu_sample(cfg):
    """The bounded sample BOTH CPU legs time (cpu_baseline of the b200 arm and --impl reference): the first
    CPU_SAMPLE_CELLS cells per dataset of the workload's generator (host numpy stream, seed0 = 1000)."""
    from pyliger_b200.synth import make_datasets
    return make_datasets([CPU_SAMPLE_CELLS] * cfg["datasets"], cfg["genes"], cfg["k"], cfg["density"], structured=True,
                         seed0=1000)


def sample_desc(cfg, name):
    frac = cfg["cells_per_dataset"] // CPU_SAMPLE_CELLS
    return ("%s sample: %d datasets x %d cells x %d genes (1/%d of the cells of the GPU workload, same generator), k=%d, "
            "fp64, cold start, numpy oracle (one Python thread + BLAS threads).  Bias: the gene-side V/W solves cost the "
            "same per iteration whatever the cell count, so per cell-iteration they weigh %dx more here than in the "
            "full workload (the phase split is reported in `phases_s`)" % (
                name, cfg["datasets"], CPU_SAMPLE_CELLS, cfg["genes"], frac, cfg["k"], frac))


def time_oracle(sample, cfg, steps, warmup):
    """CPU restatement of the reference's optimize_ALS (oracle/inmf_oracle.py) timed per iteration."""
    from oracle import inmf_oracle as orc
    marks = [time.perf_counter()]
    orc.als(sample, cfg["k"], cfg["value_lambda"], 0.0, warmup + steps, rand_seed=1, dense_objective=False,
            callback=lambda it, obj, H, W, V: marks.append(time.perf_counter()))
    dt = marks[-1] - marks[warmup]
    cells = sum(m.shape[0] for m in sample)
    return cells * steps / dt, dt / steps


def oracle_phase_split(sample, cfg):
    """Seconds of one oracle iteration spent in the cell-side (H) and the gene-side (V, W) solves."""
    from oracle import inmf_oracle as orc
    k, lam = cfg["k"], cfg["value_lambda"]
    rng = np.random.default_rng(0)
    g = sample[0].shape[1]
    W = rng.uniform(0, 2, (k, g))
    V = [rng.uniform(0, 2, (k, g)) for _ in sample]
    t0 = time.perf_counter()
    H = []
    for X, Vi in zip(sample, V):
        M = W + Vi
        G = M @ M.T + lam * Vi @ Vi.T
        H.append(orc.bpp(G, np.asarray(X.dot(M.T)).T)[0].T)
    t1 = time.perf_counter()
    for X, Hi in zip(sample, H):
        HtH = Hi.T @ Hi
        orc.bpp((1 + lam) * HtH, np.asarray(X.T.dot(Hi)).T - HtH @ W)
    orc.bpp(sum(Hi.T @ Hi for Hi in H), sum(np.asarray(X.T.dot(Hi)).T for X, Hi in zip(sample, H)))
    t2 = time.perf_counter()
    return {"H_update": t1 - t0, "VW_updates": t2 - t1}


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (numpy port = the oracle; the
    reference is pure Python and does not travel to the GPU box) on a bounded sample of the workload."""
    if rank != 0:
        return
    cfg = C2 if getattr(args, "config", "C2") == "C2" else C4
    t0 = time.time()
    sample = cpu_sample(cfg)
    log("reference arm: sample generated in %.1fs" % (time.time() - t0))
    value, s_per_step = time_oracle(sample, cfg, args.steps, args.warmup)
    desc = sample_desc(cfg, args.config)
    out = {
        "impl": "reference", "metric": "optimize_ALS cell-iterations/s", "value": value, "unit": "cell-iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s optimize_ALS %d datasets x %dk cells x %dk genes, k=%d, lambda=%g (CPU: bounded sample)" % (
            getattr(args, "config", "C2"), cfg["datasets"], cfg["cells_per_dataset"] // 1000, cfg["genes"] // 1000, cfg["k"],
            cfg["value_lambda"]), "sample": desc},
        "cpu_baseline": {"value": value, "unit": "cell-iters/s", "cores": cpu_threads(), "kind": "port", "sample": desc,
                         "phases_s": oracle_phase_split(sample, cfg)},
        "e2e": {"value": value, "unit": "cell-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------
# b200 arm
# ---------------------------------------------------------------------------------------------------------
class Ctx(object):
    pass


def timed_als(ctx, eng, steps, warmup, warm, objs=None):
    """W untimed + K timed ALS iterations between CUDA events (barrier + synchronize on both sides), max over ranks.
    als_iteration_value(speculate=True) is what optimize_ALS does: the next iteration's pass 1 is enqueued before the
    host waits for the objective -- never across the two ends of the timed region."""
    import torch
    done = 0
    objs = [] if objs is None else objs
    for i in range(warmup):
        objs.append(eng.als_iteration_value(warm=warm and done > 0, speculate=i + 1 < warmup))
        done += 1
    ctx.barrier()
    l0 = ctx.lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        objs.append(eng.als_iteration_value(warm=warm, speculate=i + 1 < steps))
    ev1.record()
    ctx.barrier()
    return ctx.max_over_ranks(ev0.elapsed_time(ev1)), ctx.lib.launch_count() - l0, objs


def kernel_breakdown(eng, warm, reps):
    """Per-kernel CUDA-event times of one iteration, launched exactly as als_iteration() does."""
    import torch
    phases = ["pass1", "bpp_H", "pass2", "allreduce", "solve_V", "solve_W", "prep", "objective"]
    acc = {p: 0.0 for p in phases}
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 1)]
        eng._begun = False
        evs[0].record()
        eng.als_begin()
        eng._begun = False
        evs[1].record()
        eng.bpp_H(mask=eng.mask_H, warm=warm, clear=eng.H_alt)
        evs[2].record()
        if eng.mc is not None:  # fused: the pass adds into every rank's copy (includes the two cross-rank barriers)
            eng.stats()
            evs[3].record()
        else:
            eng.stats(reduce=False)
            evs[3].record()
            eng.reduce_stats()  # (includes the wait for the slowest rank)
        evs[4].record()
        eng.solve_V(warm=warm)
        evs[5].record()
        eng.solve_W(warm=warm)
        evs[6].record()
        eng.prep()
        evs[7].record()
        eng.objective()
        evs[8].record()
        torch.cuda.synchronize()
        for i, p in enumerate(phases):
            acc[p] += evs[i].elapsed_time(evs[i + 1]) / reps
    return acc


def roofline_block(acc, ms_per_step, cells_local, nnz_local, k, esize, traffic_key=None):
    peak, peak_src = measured_peak_gbs()
    nnz_c = nnz_local / max(cells_local, 1)
    b_x = nnz_c * (esize + 4) + 4  # one streaming pass over a cell's CSR row
    alg = {"pass1": b_x + k * esize, "bpp_H": 2 * k * esize, "pass2": b_x + k * esize}
    b_als = 2 * b_x + 2 * k * esize
    kern = {}
    for p, ms in acc.items():
        e = {"ms": round(ms, 4)}
        if p in alg and ms > 0:
            e["algorithmic_bytes"] = alg[p] * cells_local
            e["achieved_gbs"] = alg[p] * cells_local / (ms * 1e-3) / 1e9
            e["frac"] = e["achieved_gbs"] / peak
        kern[p] = e
    dom = max(("pass1", "pass2"), key=lambda p: acc[p])  # the dominant HBM-streaming kernel
    traffic = None
    try:  # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture
        tr = json.load(open(PROFILE_TRAFFIC))
        if traffic_key and traffic_key in tr:
            traffic = tr[traffic_key]["kernels"][dom]["dram_bytes_per_launch"]
            for p_, e_ in kern.items():
                if p_ in tr[traffic_key]["kernels"]:
                    e_["ncu_dram_bytes_per_launch"] = tr[traffic_key]["kernels"][p_]["dram_bytes_per_launch"]
    except Exception:
        pass
    return {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src, "kernels": kern,
            "step": {"algorithmic_bytes_per_cell_iter": b_als,
                     "achieved_gbs": b_als * cells_local / (ms_per_step * 1e-3) / 1e9,
                     "frac": b_als * cells_local / (ms_per_step * 1e-3) / 1e9 / peak}}


def fused_leg(ctx, mats, k, lam, state, steps, warmup, warm, cells_total, fused_on, acc_main):
    """The same loop with the OTHER reduction mode (fused pass 2 + multimem.red when the headline used NCCL, and the
    other way round), so both are measured in one run."""
    import torch
    from pyliger_b200.engine import InmfEngine
    old = InmfEngine.fused_reduce
    InmfEngine.fused_reduce = bool(fused_on)
    try:
        eng = InmfEngine(mats, k, lam, dtype=torch.float32, device=ctx.dev, presharded=ctx.world > 1)
        if state is not None:
            eng.set_als_state(*state)
            eng.initial_objective()
        else:
            eng.draw_als_state(0, h_stream=ctx.rank if ctx.world > 1 else None)
            eng.initial_objective()
        ms, launches, objs = timed_als(ctx, eng, steps, warmup, warm, None)
        acc = kernel_breakdown(eng, warm, 3)
        out = {"what": "the same loop with fused_reduce=%s (pass 2 %s)" % (
                   bool(fused_on), "adds finished gene ranges into every rank's sum buffer with multimem.red, two cross-rank "
                   "barriers, no collective" if fused_on else "followed by an NCCL all-reduce"),
               "active": eng.mc is not None, "error": eng.fused_reduce_error,
               "value": cells_total * steps / (ms * 1e-3), "unit": "cell-iters/s", "ms_per_step": ms / steps,
               "steps": steps, "warmup": warmup, "gpu_launches": int(launches),
               "pass2_plus_reduction_ms": round(acc["pass2"] + acc["allreduce"], 4),
               "other_mode_pass2_plus_reduction_ms": round(acc_main["pass2"] + acc_main["allreduce"], 4),
               "last_objective": objs[-1] if objs else None}
        del eng
        torch.cuda.empty_cache()
        return out
    finally:
        InmfEngine.fused_reduce = old


def leg_c4(ctx, steps):
    """configs[3]: 8 x 250k cells x 5k genes, k = 40, the cells sharded over the ranks; data generated on the device."""
    import torch
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.synth import make_datasets_resident
    cfg = C4
    k, lam, g = cfg["k"], cfg["value_lambda"], cfg["genes"]
    per = cfg["cells_per_dataset"] // ctx.world
    t0 = time.time()
    mats = make_datasets_resident([per] * cfg["datasets"], g, k, cfg["density"], seed0=2000 + 100 * ctx.rank,
                                  device=ctx.dev, dtype=torch.float32)
    cells_local = sum(m.shape[0] for m in mats)
    nnz_local = sum(m.nnz for m in mats)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    eng = InmfEngine(mats, k, lam, dtype=torch.float32, device=ctx.dev, presharded=ctx.world > 1)
    eng.draw_als_state(0, h_stream=ctx.rank if ctx.world > 1 else None)
    objs = [eng.initial_objective()]
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    ms_total, launches, objs = timed_als(ctx, eng, steps, 3, True, objs)
    cells_total = ctx.sum_over_ranks(cells_local)
    ms_step = ms_total / steps
    acc = kernel_breakdown(eng, True, 3)
    out = {"workload": "C4 optimize_ALS: 8 datasets x 250000 cells x %d genes, k=%d, sharded over %d GPU(s) (%d cells per "
                       "dataset per GPU), fp32, warm start, data generated on the device" % (g, k, ctx.world, per),
           "value": cells_total * steps / (ms_total * 1e-3), "unit": "cell-iters/s", "scaling": "strong",
           "ms_per_step": ms_step, "steps": steps, "warmup": 3, "cells_total": cells_total,
           "nnz_per_cell": nnz_local / cells_local, "gpu_launches": int(launches),
           "objective_monotone": bool(np.all(np.diff(objs) <= 1e-6 * abs(objs[0]))),
           "roofline": roofline_block(acc, ms_step, cells_local, nnz_local, k, 4, "C4"),
           "fused_reduce": eng.mc is not None,
           "bpp_H": eng.bpp_report()["H"], "seconds": {"generate": round(t_gen, 2), "setup": round(t_setup, 2)}}
    fused_main = eng.mc is not None
    del eng
    torch.cuda.empty_cache()
    if ctx.world > 1 and ctx.fused_leg:
        out["fused_reduce_leg"] = fused_leg(ctx, mats, k, lam, None, steps, 3, True, cells_total, not fused_main, acc)
    del mats
    torch.cuda.empty_cache()
    return out


def _init_V_resident(m, k, seed):
    """_init_V_online (reference _utilities.py:69-88) for a device-resident dataset: k cells sampled with replacement
    (sorted) as unit-norm columns; same host RNG stream."""
    import torch
    np.random.seed(seed=seed)
    idx = np.sort(np.random.choice(list(range(m.shape[0])), k))
    V = torch.zeros(m.shape[1], k, dtype=torch.float64, device=m.vals.device)
    ip = m.indptr
    for j, r in enumerate(idx):
        lo, hi = int(ip[r]), int(ip[r + 1])
        V[m.colidx[lo:hi].long(), j] = m.vals[lo:hi].double()
    return (V / V.square().sum(dim=0).sqrt()).cpu().numpy()


def leg_c3(ctx):
    """configs[2]: online_iNMF over 1 M cells in 5 000-cell mini-batches (one epoch = 200 steps) + the final H pass."""
    import torch
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.factorization import _online_iNMF as onl
    from pyliger_b200.synth import make_datasets_resident
    cfg = C3
    k, lam, g, mb = cfg["k"], cfg["value_lambda"], cfg["genes"], cfg["minibatch"]
    t0 = time.time()
    mats = make_datasets_resident([cfg["cells_per_dataset"]] * cfg["datasets"], g, k, cfg["density"], seed0=7000,
                                  device=ctx.dev, dtype=torch.float32)  # the same data on every rank (X replicated)
    ns = [m.shape[0] for m in mats]
    nnz = sum(m.nnz for m in mats)
    W = onl._init_W(g, k, 1)
    Vs = [_init_V_resident(m, k, 1) for m in mats]
    eng = InmfEngine(mats, k, lam, dtype=torch.float32, device=ctx.dev, replicate=True, need_stats=False)
    mbs = np.asarray([np.round((ns[i] / np.sum(ns)) * mb).astype(int) for i in range(len(mats))])
    total_iters = int(np.floor(ns[0] * 1 / mbs[0]))
    eng.alloc_online(mbs)
    plan = eng.online_plan(total_iters)
    torch.cuda.synchronize()
    t_setup = time.time() - t0

    def reset():
        eng.Wt.copy_(onl._to_dev(eng, W))
        for i in range(len(mats)):
            eng.Vt[i].copy_(onl._to_dev(eng, Vs[i]))
        eng.A.zero_(); eng.B.zero_(); eng.A_old.zero_(); eng.B_old.zero_()

    def step(it, st):
        epoch_prev = st[0]
        epoch = ((it + 1) * mbs[0]) // ns[0]
        st[0] = epoch
        scales = [0.0] * len(mats) if it == 0 else ([1.0 / m for m in mbs] if it == 1 else [(it - 1) / it] * len(mats))
        eng.online_step(plan[it], scales, bool(epoch > 0 and epoch - epoch_prev == 1), 1, step=it)

    reset()
    st = [0]
    for it in range(min(3, total_iters)):  # warm-up, then restart from the same state
        step(it, st)
    reset()
    st = [0]
    ctx.barrier()
    l0 = ctx.lib.launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for it in range(total_iters):
        step(it, st)
    e1.record()
    eng.final_H()
    e2.record()
    ctx.barrier()
    ms_train = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_final = ctx.max_over_ranks(e1.elapsed_time(e2))
    launches = ctx.lib.launch_count() - l0
    peak, _ = measured_peak_gbs()
    cells = sum(ns)
    visits = total_iters * int(mbs.sum())
    b_cell = nnz / cells * 8 + 4 + 3 * k * 4  # pass 1 over the cell's row + the right-hand side + the in-place solve
    cells_rank = cells / ctx.world
    out = {"workload": "C3 online_iNMF: %d datasets x %d cells x %d genes (%d cells), k=%d, miniBatch_size=%d, 1 epoch, "
                       "fp32, %d GPU(s) (every rank takes 1/%d of each mini-batch and of the final H pass)" % (
                           len(mats), ns[0], g, cells, k, mb, ctx.world, ctx.world),
           "steps": total_iters, "ms_per_step": ms_train / max(total_iters, 1),
           "train_cell_visits_per_s": visits / (ms_train * 1e-3), "final_H_ms": ms_final,
           "final_H_cells_per_s": cells / (ms_final * 1e-3),
           "final_H_roofline": {"bound": "hbm", "achieved": b_cell * cells_rank / (ms_final * 1e-3) / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": b_cell * cells_rank / (ms_final * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes_per_cell": b_cell},
           "train_roofline": {"bound": "latency (chain of small dependent kernels)",
                              "achieved": b_cell * visits / ctx.world / (ms_train * 1e-3) / 1e9, "peak": peak,
                              "unit": "GB/s", "frac": b_cell * visits / ctx.world / (ms_train * 1e-3) / 1e9 / peak},
           "gpu_launches": int(launches), "bpp_H": eng.bpp_report()["H"], "seconds": {"setup": round(t_setup, 2)}}
    del eng, mats
    torch.cuda.empty_cache()
    return out


def leg_c5(ctx, cpu):
    """configs[4]: batched nnlsm_blockpivot through the C ABI, 10 M columns, H-update-shaped problems."""
    import torch
    from pyliger_b200.engine import _stream
    from pyliger_b200.synth import make_datasets_resident
    lib, dev = ctx.lib, ctx.dev
    cols = C5["columns"]
    X = make_datasets_resident([C5["base_cells"]], C5["genes"], 40, 0.08, seed0=5000, device=dev, dtype=torch.float64)[0]
    g = X.shape[1]
    Xd = torch.sparse_csr_tensor(X.rowptr, X.colidx.long(), X.vals, size=X.shape)
    res = []
    for k in C5["ks"]:
        rng = np.random.default_rng(k)
        Wm = np.abs(rng.uniform(0, 2, (k, g)))
        Vm = np.abs(rng.uniform(0, 2, (k, g)))
        M = Wm + Vm
        G = M @ M.T + 5.0 * Vm @ Vm.T
        Rt_base = torch.sparse.mm(Xd, torch.from_numpy(M.T.copy()).to(dev)).float()  # cells x k
        Rt = torch.empty(cols, k, dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(k)
        nb = Rt_base.shape[0]
        for o in range(0, cols, nb):  # replicate with a multiplicative jitter so that columns are not identical
            m = min(nb, cols - o)
            Rt[o:o + m] = Rt_base[:m] * (1.0 + 0.05 * torch.randn(m, k, device=dev, dtype=torch.float32, generator=gen))
        Gd = torch.from_numpy(G).to(dev)
        Xo = torch.empty_like(Rt)
        mask = torch.zeros(cols, dtype=torch.int64, device=dev)
        seg = torch.tensor([0, cols], dtype=torch.int64, device=dev)
        ws = lib.bpp_workspace(cols, 1, k, dev)
        row = {"k": k, "columns": cols, "dtype": "f32"}
        for mode, warm in (("cold", 0), ("warm", 1)):
            times = []
            for rep in range(3):
                stats = torch.zeros(8, dtype=torch.int64, device=dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.call("inmf_bpp", "f32", Gd, Rt, k, Xo, k, None, 0, mask, warm, seg, 1, cols, k, stats,
                         *lib.ws_args(ws), _stream())
                e1.record()
                torch.cuda.synchronize()
                if rep > 0:
                    times.append(e0.elapsed_time(e1))
            s = stats.cpu().numpy()
            ms = float(np.mean(times))
            row[mode] = {"ms": ms, "columns_per_s": cols / (ms * 1e-3), "factorizations_per_column": s[1] / s[6],
                         "not_converged": int(s[0]), "deferred_to_full_warp": int(s[7]),
                         "hbm_gbs": 2 * k * 4 * cols / (ms * 1e-3) / 1e9}
        row["passive_fraction"] = float((Xo[:1000000] > 0).float().mean().item())
        if cpu:
            from oracle import inmf_oracle as orc
            n = 20000
            Rh = Rt[:n].double().cpu().numpy().T.copy()
            t0 = time.perf_counter()
            Xc, _ = orc.bpp(G, Rh)
            dtc = time.perf_counter() - t0
            err = np.linalg.norm(Xo[:n].double().cpu().numpy().T - Xc) / np.linalg.norm(Xc)
            row["cpu_oracle"] = {"columns": n, "columns_per_s": n / dtc, "rel_err_gpu_vs_oracle": float(err),
                                 "cores": cpu_threads()}
        res.append(row)
        del Rt, Xo, mask, ws
        torch.cuda.empty_cache()
    return {"workload": "C5 batched nnlsm_blockpivot (inmf_bpp_f32 through the C ABI), %d columns per k, right-hand sides "
                        "X (W+V)' of the synthetic generator replicated with 5%% jitter, G = (W+V)(W+V)' + 5 VV'" % cols,
            "results": res}


def leg_parity(ctx):
    """N > 1: the row-sharded N-rank path (every rank is handed the FULL datasets and works on its slice; full H
    gathered) against the 1-rank path on the same small problem, in the same process."""
    import torch
    from pyliger_b200.engine import InmfEngine
    from pyliger_b200.factorization._iNMF_ANLS import _als_init_host
    from pyliger_b200.synth import make_datasets
    mats = make_datasets([1500, 1100, 700], 400, 12, density=0.1, seed0=4200)
    k, lam, iters = 12, 5.0, 8
    ns, g = [m.shape[0] for m in mats], mats[0].shape[1]
    W0, V0, H0 = _als_init_host(ns, g, k, 3)
    solo = [torch.distributed.new_group([r]) for r in range(ctx.world)][ctx.rank]  # collective: every rank creates all
    out = {}
    for name, group in (("sharded", None), ("single", solo)):
        for dtype in (torch.float64, torch.float32):
            eng = InmfEngine(mats, k, lam, dtype=dtype, device=ctx.dev, group=group)
            eng.set_als_state(W0, V0, H0)
            objs = [eng.initial_objective()] + [eng.als_iteration_value(warm=i > 0) for i in range(iters)]
            H = eng.gather_H()
            out[(name, dtype)] = (np.array(objs), np.vstack(H), eng.host_W())
            del eng
    res = {"problem": "3 datasets %s cells x %d genes, k=%d, %d iterations, non-presharded" % (ns, g, k, iters)}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        a, b = out[("sharded", dtype)], out[("single", dtype)]
        res[tag] = {"objective_rel_diff_max": float(np.max(np.abs(a[0] - b[0]) / np.abs(b[0]))),
                    "H_rel_fro": float(np.linalg.norm(a[1] - b[1]) / np.linalg.norm(b[1])),
                    "W_rel_fro": float(np.linalg.norm(a[2] - b[2]) / np.linalg.norm(b[2]))}
    res["ok"] = bool(res["f64"]["objective_rel_diff_max"] < 1e-9 and res["f32"]["objective_rel_diff_max"] < 1e-3)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--warm-start", type=int, default=1)
    ap.add_argument("--config", default="C2", choices=["C2", "C4"],
                    help="headline workload: C2 (default, weak scaling) or C4 (8 x 250k cells sharded over the ranks)")
    ap.add_argument("--cells", type=int, default=0, help="cells per dataset per GPU (default: the config's)")
    ap.add_argument("--graph", type=int, default=-1,
                    help="-1: the engine's policy (CUDA-graphed warm iterations when N > 1, eager launches on one GPU), "
                         "1: graphs from the first warm iteration, 0: eager")
    ap.add_argument("--fused-reduce", type=int, default=0,
                    help="N > 1: 0 = pass 2 + NCCL all-reduce (the engine default), 1 = pass 2 adds its finished gene "
                         "ranges into multicast-bound symmetric sum buffers (no separate all-reduce); the other mode is "
                         "measured in extra.c2_fused_reduce / c4_strong.fused_reduce_leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only (skip the c2 cold / fp64, C3, C4, C5 legs)")
    ap.add_argument("--extra", default="cold,mixed,f64,c4,c3,c5,parity,fused", help="which extra legs to run")
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
    InmfEngine.use_graph = None if args.graph < 0 else bool(args.graph)
    InmfEngine.fused_reduce = bool(args.fused_reduce)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    ctx = Ctx()
    ctx.rank, ctx.world, ctx.dev, ctx.lib = rank, world, dev, _lib
    ctx.fused_leg = "fused" in [e for e in args.extra.split(",") if e] and not args.no_extra

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

    ctx.barrier, ctx.max_over_ranks, ctx.sum_over_ranks = barrier, max_over_ranks, sum_over_ranks
    base = C2 if args.config == "C2" else C4
    strong = args.config == "C4"
    cfg = dict(base, cells_per_dataset=args.cells or (base["cells_per_dataset"] // world if strong
                                                      else base["cells_per_dataset"]))
    k, lam, g = cfg["k"], cfg["value_lambda"], cfg["genes"]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    esize = 4 if args.dtype == "f32" else 8
    warm = bool(args.warm_start)
    extras = set() if args.no_extra else set(args.extra.split(","))

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
    objs = [eng.initial_objective()]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, objs = timed_als(ctx, eng, args.steps, args.warmup, warm, objs)
    cells_total = sum_over_ranks(cells_local)
    value = cells_total * args.steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps
    mono = bool(np.all(np.diff(objs) <= 1e-6 * abs(objs[0])))
    log("rank %d: objectives first/last %.6e %.6e monotone=%s bpp=%s" % (rank, objs[1], objs[-1], mono,
                                                                       eng.bpp_report()["H"]))
    acc = kernel_breakdown(eng, warm, 5)
    fused = eng.mc is not None
    roofline = roofline_block(acc, ms_per_step, cells_local, nnz_local, k, esize,
                              args.config if args.dtype == "f32" else None)
    extra = {}

    # ------------------------------------------------------------------ same loop, cold-started solves
    if "cold" in extras:
        eng.set_als_state(W0, V0, H0)
        eng.initial_objective()
        ms_c, launches_c, _ = timed_als(ctx, eng, args.steps, args.warmup, False)
        extra["c2_cold_f32" if args.dtype == "f32" else "c2_cold_f64"] = {
            "what": "the headline loop with cold-started BPP solves (warm_start=False, the reference's init=None)",
            "value": cells_total * args.steps / (ms_c * 1e-3), "unit": "cell-iters/s", "ms_per_step": ms_c / args.steps,
            "steps": args.steps, "warmup": args.warmup, "gpu_launches": int(launches_c)}
    del eng
    torch.cuda.empty_cache()
    if world > 1 and "fused" in extras and args.dtype == "f32":
        extra["c2_fused_reduce"] = fused_leg(ctx, mats, k, lam, (W0, V0, H0), args.steps, args.warmup, warm, cells_total,
                                             not fused, acc)
    if "mixed" in extras and args.dtype == "f32" and k <= 32:
        engm = InmfEngine(mats, k, lam, dtype="mixed", device=dev, presharded=world > 1)
        engm.set_als_state(W0, V0, H0)
        om = [engm.initial_objective()]
        ms_m, launches_m, om = timed_als(ctx, engm, args.steps, args.warmup, warm, om)
        accm = kernel_breakdown(engm, warm, 5)
        extra["c2_mixed"] = {
            "what": "the headline loop with dtype='mixed': the dense operand of the two passes rounded to bf16 in shared "
                    "memory (64 instead of 128 operand bytes per non-zero), everything else fp32 -- a labelled "
                    "reduced-precision mode, not the fp32 headline",
            "value": cells_total * args.steps / (ms_m * 1e-3), "unit": "cell-iters/s", "ms_per_step": ms_m / args.steps,
            "steps": args.steps, "warmup": args.warmup, "gpu_launches": int(launches_m),
            "final_objective_rel_diff_vs_fp32": abs(om[-1] - objs[-1]) / abs(objs[-1]),
            "roofline": roofline_block(accm, ms_m / args.steps, cells_local, nnz_local, k, 4)}
        del engm
        torch.cuda.empty_cache()
    if "f64" in extras and args.dtype == "f32":
        eng64 = InmfEngine(mats, k, lam, dtype=torch.float64, device=dev, presharded=world > 1)
        eng64.set_als_state(W0, V0, H0)
        o64 = [eng64.initial_objective()]
        s64 = max(3, min(args.steps, 10))
        ms_d, launches_d, o64 = timed_als(ctx, eng64, s64, 3, warm, o64)
        acc64 = kernel_breakdown(eng64, warm, 2)
        extra["c2_warm_f64"] = {
            "what": "the headline loop in fp64 (dtype='float64': the API default / parity mode)",
            "value": cells_total * s64 / (ms_d * 1e-3), "unit": "cell-iters/s", "ms_per_step": ms_d / s64, "steps": s64,
            "warmup": 3, "gpu_launches": int(launches_d),
            "roofline": roofline_block(acc64, ms_d / s64, cells_local, nnz_local, k, 8)}
        del eng64
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ end to end through the public API
    e2e = None
    if not args.no_e2e:
        lig = liger_from_matrices(mats)
        optimize_ALS(lig, k, lam, 1e-6, 2, dtype=args.dtype.replace("f", "float"), presharded=world > 1,
                     progress=False, warm_start=warm)  # warm-up of the whole call path
        lig = liger_from_matrices(mats)
        barrier()
        t0 = time.perf_counter()
        info = optimize_ALS(lig, k, lam, 1e-6, args.steps, dtype=args.dtype.replace("f", "float"),
                            presharded=world > 1, progress=False, warm_start=warm, return_info=True)
        torch.cuda.synchronize()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        iters = info["iters_run"][0]
        h2d, d2h = info["h2d_bytes"], info["d2h_bytes"]  # counted from the arrays actually copied
        e2e = {"value": cells_total * iters / t_e2e, "unit": "cell-iters/s", "h2d_bytes_per_step": h2d / iters,
               "d2h_bytes_per_step": d2h / iters, "seconds": t_e2e, "iters": iters,
               "what": "wall time of pyliger_b200.optimize_ALS(liger, k, max_iters=steps) on host scipy CSR "
                       "inputs: CSR pack + H2D + init + all iterations + H/W/V write-back to host float64"}
        del lig
    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed loops and the e2e call

    # ------------------------------------------------------------------ the other named configs
    if args.config == "C2":
        if "c4" in extras:
            extra["c4_strong"] = leg_c4(ctx, 5)
        if "c3" in extras:
            extra["c3_online"] = leg_c3(ctx)
        if "c5" in extras and world == 1:
            extra["c5_bpp"] = leg_c5(ctx, cpu=not args.no_cpu)
        if "parity" in extras and world > 1:
            extra["parity_vs_1gpu"] = leg_parity(ctx)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        base_cfg = dict(cfg, cells_per_dataset=base["cells_per_dataset"])
        sample = cpu_sample(base_cfg)
        v, s_step = time_oracle(sample, cfg, 2, 1)
        cpu = {"value": v, "unit": "cell-iters/s", "cores": cpu_threads(), "kind": "port",
               "sample": sample_desc(base_cfg, args.config) + "; 2 timed ALS iterations after 1 warm-up",
               "phases_s": oracle_phase_split(sample, cfg)}

    if rank == 0:
        out = {
            "metric": "optimize_ALS cell-iterations/s", "value": value, "unit": "cell-iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.config + " optimize_ALS: %d datasets x %d cells x %d genes per GPU, density %.2f, k=%d, "
                                   "lambda=%g" % (len(mats), cfg["cells_per_dataset"], g, cfg["density"], k, lam),
                       "cells_total": cells_total, "nnz_per_cell": nnz_local / cells_local, "warm_start": warm,
                       "options": "optimize_ALS API defaults except dtype (warm_start=True is the default; dtype='float32' "
                                  "is the throughput mode, the fp64 default is in extra.c2_warm_f64)",
                       "l2": "inputs larger than L2 (CSR %.0f MB per GPU streamed twice per step)" % (
                           nnz_local * (esize + 4) / 1e6),
                       "parallelism": "cells sharded over %d GPU(s), %s" % (
                           world, "single GPU" if world == 1 else (
                               "pass 2 adds its k x k and genes x k statistics into every rank's copy through the "
                               "NVLink switch (multimem.red, no separate all-reduce)" if fused else
                               "NCCL all-reduce of the k x k and genes x k statistics"))},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "objective_monotone": mono, "extra": extra,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
