#!/usr/bin/env python
"""Regenerate diffmvs_b200/tuned/b200_default.json: the per-layer back-end table shipped for B200.

Runs one eager forward of every named workload (and the scan-mode call patterns of the headline one) with the shipped
table disabled, three times over, and keeps for every convolution signature the back end with the smallest median time.
Run on a B200:  DMVS_TUNED_TABLE=0 python tools/make_default_tuned.py"""
import json, os, statistics, sys
os.environ["DMVS_TUNED_TABLE"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffmvs_b200 import ops, synth
from diffmvs_b200.models import CasDiffMVS
from diffmvs_b200.scan import ScanRunner

dev = torch.device("cuda", 0)
names = {v: k for k, v in ops.PRECISIONS.items()}
votes = {}
for rep in range(3):
    ops._TUNED.clear()
    for wl in ("cfg3", "cfg4", "cfg2", "cfg1"):
        args = synth.workload_args(wl)
        model = CasDiffMVS(args, test=True)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(synth.synth_state_dict(shapes, 123), strict=False)
        model.to(dev).eval()
        imgs, proj, dv = synth.workload_inputs(wl, seed=0)
        imgs = [i.to(dev) for i in imgs]
        proj = {k: v.to(dev) for k, v in proj.items()}
        dv = dv.to(dev)
        with torch.no_grad():
            model(imgs, proj, dv)
            if wl == "cfg3":     # four reference views per call (bench.py `batched`)
                bi, bp, bd = synth.workload_inputs(wl, seed=0, batch=4)
                model([t.to(dev) for t in bi], {k: v.to(dev) for k, v in bp.items()}, bd.to(dev))
                del bi, bp, bd
            if wl == "cfg3":     # scan mode: FeatureNet on the views that miss the cache (1 .. V-1 images)
                runner = ScanRunner(model, capacity=2 * len(imgs))
                for start in range(4):
                    runner(list(range(start, start + len(imgs))), imgs, proj, dv)
        del model
        torch.cuda.empty_cache()
    for k, (choice, times) in ops._TUNED.items():
        sig = k[1:]
        v = votes.setdefault(sig, {})
        for code, t in times.items():
            v.setdefault(code, []).append(t)
rows = []
for sig, v in votes.items():
    med = {code: statistics.median(ts) for code, ts in v.items()}
    best = min(med, key=med.get)
    rows.append({"sig": [int(x) for x in sig], "choice": names[best], "ms": {names[c]: round(t, 5) for c, t in med.items()}})
out = os.path.join(ROOT, "diffmvs_b200", "tuned", "b200_default.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, "w"), indent=0)
print(f"{len(rows)} signatures -> {out}")
from collections import Counter
print(Counter(r["choice"] for r in rows))
