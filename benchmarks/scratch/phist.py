import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from pyliger_b200.engine import InmfEngine
from pyliger_b200.synth import make_datasets_device
dev = torch.device("cuda:0")
for (N, n, g, k) in ((4, 50000, 3000, 30), (8, 60000, 5000, 40)):
    mats = make_datasets_device([n] * N, g, k, 0.08, seed0=1000, device=dev)
    eng = InmfEngine(mats, k, 5.0, dtype=torch.float32, device=dev)
    eng.draw_als_state(0)
    eng.initial_objective()
    for it in range(8):
        eng.als_iteration(warm=it > 0)
    m = eng.mask_H.cpu().numpy().astype(np.uint64)
    p = np.array([bin(int(x)).count("1") for x in m[:200000]])
    print("k", k, "mean p %.1f" % p.mean(), "hist(<=8,<=15,<=23,<=31,>31):",
          [(p <= 8).mean().round(3), ((p > 8) & (p <= 15)).mean().round(3), ((p > 15) & (p <= 23)).mean().round(3),
           ((p > 23) & (p <= 31)).mean().round(3), (p > 31).mean().round(3)])
    print(eng.bpp_report()["H"])
