#!/usr/bin/env python
"""Throughput of the graph-replayed forward with B reference views per call (cfg3): ref-views/s for B = 1, 2, 3, 4."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffmvs_b200 import ops, synth
from diffmvs_b200.models import CasDiffMVS

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dev = torch.device("cuda", 0)
args = synth.workload_args(wl)
model = CasDiffMVS(args, test=True)
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
model.load_state_dict(synth.synth_state_dict(shapes, 123), strict=False)
model.to(dev).eval()
model.use_cuda_graph(True)
for B in (1, 2, 3, 4):
    imgs, proj, dv = synth.workload_inputs(wl, seed=0, batch=B)
    imgs = [i.to(dev) for i in imgs]
    proj = {k: v.to(dev) for k, v in proj.items()}
    dv = dv.to(dev)
    with torch.no_grad():
        for _ in range(4):
            model(imgs, proj, dv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 12
        e0.record()
        for _ in range(n):
            model(imgs, proj, dv)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{wl} B={B}: {ms:.3f} ms per call, {ms / B:.3f} ms per ref-view, {1e3 * B / ms:.1f} ref-views/s", flush=True)
