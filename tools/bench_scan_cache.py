#!/usr/bin/env python
"""Scan-mode throughput with the cross-ref-view feature cache (NOT the headline metric: the
unit of work changes).  In a scan every image serves as a source view of its neighbours, so with the pyramids kept
each reference view needs FeatureNet on about one new image instead of V.  Prints one JSON line."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import synth
from diffmvs_b200.models import CasDiffMVS

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
args = synth.workload_args(wl)
model = CasDiffMVS(args, test=True)
model.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 123), strict=False)
model.cuda().eval()
imgs, proj, dv = synth.workload_inputs(wl)
imgs = [i.cuda() for i in imgs]; proj = {k: v.cuda() for k, v in proj.items()}; dv = dv.cuda()

def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

feats = model(imgs, proj, dv, return_features=True)["features"]
t_all = timed(lambda: model(imgs, proj, dv))                                  # eager, every view encoded
t_cached = timed(lambda: model(imgs, proj, dv, features=[None] + feats[1:]))  # new reference image, cached sources
print(json.dumps({"metric": "ref-views/s in scan mode (1 new image per ref-view, sources from the feature cache)",
                  "workload": wl, "ms_per_ref_view_no_cache_eager": t_all, "ms_per_ref_view_cached_eager": t_cached,
                  "value": 1e3 / t_cached, "unit": "ref-views/s", "note": "eager launches (no CUDA graph)"}))
