import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from pyliger_b200.engine import InmfEngine
from pyliger_b200.factorization._iNMF_ANLS import _als_init_host
from pyliger_b200.synth import make_datasets
big = len(sys.argv) > 1
if big:
    from pyliger_b200.synth import make_datasets_device
    mats = make_datasets_device([50000]*4, 3000, 30, 0.08, seed0=1000, device='cuda')
    k = 30
else:
    mats = make_datasets([700, 333], 200, 9, density=0.1, seed0=4100); k = 9
ns=[m.shape[0] for m in mats]; g=mats[0].shape[1]
W,V,H=_als_init_host(ns,g,k,0)
ref=InmfEngine(mats,k,5.0,dtype=torch.float32); ref.set_als_state(W,V,H); ref.prep()
for three in (True, False):
    tc=InmfEngine(mats,k,5.0,dtype=torch.float32,use_tc=True,tc_three_pass=three); tc.set_als_state(W,V,H); tc.prep()
    ref.set_als_state(W,V,H); ref.stats(); tc.stats(); torch.cuda.synchronize()
    e_st=(tc.St-ref.St).norm()/ref.St.norm(); e_hh=(tc.HtH-ref.HtH).norm()/ref.HtH.norm()
    ref.rhs_H(); tc.rhs_H(); torch.cuda.synchronize()
    e_r=(tc.H-ref.H).norm()/ref.H.norm()
    print('three_pass',three,'rel err R %.3e St %.3e HtH %.3e' % (e_r.item(), e_st.item(), e_hh.item()), 'pad zero', float(tc.H[:,k:].abs().max().item()) if tc.ldm>k else 0.0, flush=True)
    if big:
        for name,fn_ref,fn_tc in (('pass1',ref.rhs_H,tc.rhs_H),('pass2',ref.stats,tc.stats)):
            for nm,fn in (('blocked',fn_ref),('tc',fn_tc)):
                for _ in range(2): fn()
                torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5): fn()
                e1.record(); torch.cuda.synchronize()
                print('  %s %s %.3f ms' % (name,nm,e0.elapsed_time(e1)/5), flush=True)
    del tc
