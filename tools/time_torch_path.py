#!/usr/bin/env python
"""Times the oracle restatement with stock torch CUDA ops (what the reference would run on this GPU)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import synth
from oracle import diffmvs_ref as O, spec
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
args = synth.workload_args(wl)
sd = {k: v.cuda() for k, v in synth.synth_state_dict(spec.state_dict_shapes(args), 123).items()}
imgs, proj, dv = synth.workload_inputs(wl)
imgs = [i.cuda() for i in imgs]; proj = {k: v.cuda() for k, v in proj.items()}; dv = dv.cuda()
torch.backends.cudnn.benchmark = True
for tf32 in (True, False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        for _ in range(3):
            O.casdiffmvs_forward(sd, args, imgs, proj, dv)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            O.casdiffmvs_forward(sd, args, imgs, proj, dv)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{wl} stock torch CUDA path tf32={tf32}: {dt*1e3:.1f} ms/ref-view = {1/dt:.2f} ref-views/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
