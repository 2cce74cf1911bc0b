#!/usr/bin/env python
"""Wall time of the public optimize_ALS call on host scipy inputs (C2 shape), CUDA-graphed iterations on / off, after
warm-up calls: what the graphs cost (capture + instantiation per call) against what they save per iteration."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pyliger_b200 import liger_from_matrices, optimize_ALS  # noqa: E402
from pyliger_b200.engine import InmfEngine  # noqa: E402
from pyliger_b200.synth import make_datasets_device  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    mats = make_datasets_device([50000] * 4, 3000, 30, 0.08, seed0=1000, device=dev)
    k, lam = 30, 5.0
    for iters in (20, 100):
        for graph in (True, False, True, False):
            InmfEngine.use_graph = graph
            ts = []
            for rep in range(4):
                lig = liger_from_matrices(mats)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                optimize_ALS(lig, k, lam, 1e-12, iters, dtype="float32", progress=False)
                torch.cuda.synchronize()
                ts.append(1e3 * (time.perf_counter() - t0))
            print("iters %d graph %s: %s ms" % (iters, graph, ["%.1f" % t for t in ts]), flush=True)


if __name__ == "__main__":
    main()
