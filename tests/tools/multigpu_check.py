#!/usr/bin/env python
"""Multi-GPU parity check (launch with torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/tools/multigpu_check.py

Every rank passes the FULL datasets; the engine shards the cells, all-reduces the statistics and
gathers H.  Results must match the single-process numpy oracle (and hence the 1-GPU run) to fp64
tolerance; summation order differs across ranks, so the check is at tolerance, not bitwise."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import inmf_oracle as orc
    from pyliger_b200 import liger_from_matrices, online_iNMF, optimize_ALS
    from pyliger_b200.synth import make_datasets

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    mats = make_datasets([1301, 997, 640], 300, 10, density=0.1, seed0=6000)
    k, lam = 10, 5.0
    want = orc.als(mats, k, lam, 1e-6, 5, rand_seed=1)
    ok = True
    for dtype, otol, ftol in (("float64", 1e-9, 1e-6), ("float32", 1e-3, 1e-2)):
        lig = liger_from_matrices(mats)
        info = optimize_ALS(lig, k, lam, 1e-6, 5, rand_seed=1, dtype=dtype, return_info=True, progress=False)
        rel = np.abs(np.array(info["objectives"][0]) - np.array(want["objectives"])) / np.array(want["objectives"])
        herr = max(np.linalg.norm(lig.H[i] - want["H"][i]) / np.linalg.norm(want["H"][i]) for i in range(3))
        werr = np.linalg.norm(lig.W.T - want["W"]) / np.linalg.norm(want["W"])
        good = rel.max() < otol and herr < ftol and werr < ftol and lig.H[0].shape == want["H"][0].shape
        ok &= bool(good)
        print("rank %d/%d ALS %s: obj rel err %.2e, H err %.2e, W err %.2e -> %s (fused pass 2 + all-reduce: %s%s)" % (
            rank, world, dtype, rel.max(), herr, werr, "OK" if good else "FAIL", info.get("fused_reduce"),
            (", " + info["fused_reduce_error"][:200]) if info.get("fused_reduce_error") else ""), flush=True)
    # the fused path against pass 2 + NCCL all-reduce on a k = 40 shape (one-lane-per-row kernel with remainder tile)
    from pyliger_b200.engine import InmfEngine
    mats40 = make_datasets([2100, 1500], 500, 40, density=0.1, seed0=6100)
    res = {}
    for fused in (True, False):
        InmfEngine.fused_reduce = fused
        lig = liger_from_matrices(mats40)
        info = optimize_ALS(lig, 40, lam, 1e-6, 4, rand_seed=1, dtype="float32", return_info=True, progress=False)
        res[fused] = (np.array(info["objectives"][0]), lig.W.copy(), info.get("fused_reduce"))
    InmfEngine.fused_reduce = True
    rel = np.abs(res[True][0] - res[False][0]) / res[False][0]
    werr = np.linalg.norm(res[True][1] - res[False][1]) / np.linalg.norm(res[False][1])
    good = rel.max() < 1e-5 and werr < 1e-3 and res[False][2] is False
    ok &= bool(good)
    print("rank %d/%d k=40 fused (%s) vs NCCL: obj rel diff %.2e, W diff %.2e -> %s" % (
        rank, world, res[True][2], rel.max(), werr, "OK" if good else "FAIL"), flush=True)
    wo = orc.online(mats, k, lam, 2, 1, 300, 1000, 1)
    lig = liger_from_matrices(mats)
    online_iNMF(lig, k=k, value_lambda=lam, max_epochs=2, miniBatch_size=300, verbose=False)
    werr = np.linalg.norm(lig.W - wo["W"]) / np.linalg.norm(wo["W"])
    herr = max(np.linalg.norm(lig.H[i] - wo["H"][i]) / np.linalg.norm(wo["H"][i]) for i in range(3))
    aerr = np.linalg.norm(lig.adata_list[1].uns["A"] - wo["A"][1]) / np.linalg.norm(wo["A"][1])
    good = werr < 1e-7 and herr < 1e-6 and aerr < 1e-7
    ok &= bool(good)
    print("rank %d/%d online fp64: W err %.2e, H err %.2e, A err %.2e -> %s" % (
        rank, world, werr, herr, aerr, "OK" if good else "FAIL"), flush=True)
    # out-of-core path: every rank streams only its share of each mini-batch and of the final H pass (X sharded)
    lig = liger_from_matrices(mats)
    online_iNMF(lig, k=k, value_lambda=lam, max_epochs=2, miniBatch_size=300, verbose=False, stream=True)
    werr = np.linalg.norm(lig.W - wo["W"]) / np.linalg.norm(wo["W"])
    herr = max(np.linalg.norm(lig.H[i] - wo["H"][i]) / np.linalg.norm(wo["H"][i]) for i in range(3))
    good = werr < 1e-7 and herr < 1e-6 and lig.H[2].shape == wo["H"][2].shape
    ok &= bool(good)
    print("rank %d/%d streamed online fp64: W err %.2e, H err %.2e -> %s" % (
        rank, world, werr, herr, "OK" if good else "FAIL"), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item() != 0:
        raise SystemExit(1)
    if rank == 0:
        print("MULTIGPU CHECK OK (world=%d)" % world, flush=True)


if __name__ == "__main__":
    main()
