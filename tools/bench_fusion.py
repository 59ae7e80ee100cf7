#!/usr/bin/env python
"""Throughput of the GPU depth filter / fusion (one reference view against its source views) next to the numpy +
OpenCV path of the reference (oracle/filter_ref.py) on the host.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import fusion
from oracle import filter_ref as F
from tests.helpers import plane_scene

H, W, V = 1152, 1600, 7
sc = plane_scene(H, W, V, 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ref_d, confs, img = t(sc["depth"][0]), [t(c) for c in sc["conf"]], t(sc["img"])
src = [(t(sc["depth"][v]), sc["K"], sc["E"][v]) for v in range(1, V)]
run = lambda: fusion.fuse_view(ref_d, sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], confs, [0.3, 0.5, 0.5], src,
                               ref_img=img, geo_mask_thres=3)
for _ in range(3):
    out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t0 = time.perf_counter()
ref = F.fuse_view(sc["depth"][0], sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], sc["conf"], [0.3, 0.5, 0.5],
                  [(sc["depth"][v], sc["K"], sc["E"][v]) for v in range(1, V)], ref_img=sc["img"], geo_mask_thres=3)
cpu_s = time.perf_counter() - t0
px = H * W
# per source view: ref depth + gathered src depth in, mask + reprojected depth + 2 coordinate maps out; fuse: 3 maps in, avg + 2 masks + xyz out
algo = (V - 1) * px * (4 + 4 + 1 + 4 + 8 + 8) + px * (4 + 4 + 4 + 1 + 8 + 2 + 12)
same = bool(np.array_equal(out["final_mask"].cpu().numpy(), ref["final_mask"]))
print(json.dumps({"metric": "fused ref-views/s (geometric + photometric filter, 1600x1152, 6 source views)", "value": 1e3 / ms,
                  "ms_per_ref_view": ms, "algorithmic_GBps": algo / ms / 1e6, "points": int(out["points"].shape[0]),
                  "final_mask_equal_to_oracle": same, "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "ref-views/s", "kind": "port",
                  "sample": "1 reference view, numpy + cv2.remap on the host"}}))
