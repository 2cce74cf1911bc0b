#!/usr/bin/env python
"""tcgen05 plumbing check: one 128 x 64 x n tile, TF32 and 3xTF32, against numpy."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyliger_b200 import _lib
from pyliger_b200.engine import _stream

lib = _lib.load()
rng = np.random.default_rng(0)
verbose = len(sys.argv) > 1
for n in (32, 16, 48, 64):
    A = rng.uniform(-1, 1, (128, 64)).astype(np.float32)
    B = rng.uniform(-1, 1, (64, n)).astype(np.float32)
    if verbose:
        A[:] = 0; B[:] = 0
        A[np.arange(128), np.arange(128) % 64] = 1.0 + np.arange(128) / 128.0   # row m picks column m%64 of B
        B[:] = (np.arange(64)[:, None] * 100 + np.arange(n)[None, :]).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64)
    for three in (0, 1):
        Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        C = torch.full((128, n), float("nan"), device="cuda")
        rc = lib.inmf_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), n, three, _stream())
        torch.cuda.synchronize()
        got = C.cpu().numpy().astype(np.float64)
        err = np.abs(got - want).max() / np.abs(want).max()
        print("n=%d three_pass=%d rc=%d max rel err %.3e  nan=%d nonzero=%d" % (
            n, three, rc, err, int(np.isnan(got).sum()), int((got != 0).sum())), flush=True)
        if verbose and three == 0:
            np.set_printoptions(linewidth=200, precision=1, suppress=True)
            print("got rows 0,1,9,70:\n", got[[0, 1, 9, 70], :8]); print("want:\n", want[[0, 1, 9, 70], :8])
    if verbose:
        break
