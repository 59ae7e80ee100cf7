#!/usr/bin/env python
"""Per-call CUDA-event breakdown of one forward (kernel family x layer shape), sorted by device time."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import ops, synth  # noqa: E402
from diffmvs_b200.models import CasDiffMVS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--top", type=int, default=60)
a = ap.parse_args()
args = synth.workload_args(a.workload)
model = CasDiffMVS(args, test=True)
model.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 123), strict=False)
model.cuda().eval()
imgs, proj, dv = synth.workload_inputs(a.workload)
imgs = [i.cuda() for i in imgs]; proj = {k: v.cuda() for k, v in proj.items()}; dv = dv.cuda()
for _ in range(3):
    model(imgs, proj, dv)
prof = ops.Profiler()
ops.set_profiler(prof)
model(imgs, proj, dv)
ops.set_profiler(None)
s = prof.summary()
tot = sum(v["ms"] for v in s.values())
print(f"total {tot:.3f} ms over {sum(v['calls'] for v in s.values())} calls")
print(f"{'op':22s} {'layer':48s} {'calls':>5s} {'ms':>8s} {'%':>6s} {'GB/s':>8s}")
for (name, tag), v in sorted(s.items(), key=lambda kv: -kv[1]["ms"])[:a.top]:
    gbs = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else 0
    print(f"{name:22s} {tag:48s} {v['calls']:5d} {v['ms']:8.3f} {100 * v['ms'] / tot:6.2f} {gbs:8.1f}")
