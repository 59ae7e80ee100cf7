#!/usr/bin/env python
"""Pipeline timeline of the TMA-fed convolution kernel (conv_ws2.cu) for one layer: run with DMVS_WS2_DBG=1.

    DMVS_WS2_DBG=1 python tools/ws2_timeline.py "feat.conv1.1"

Prints, for CTA 0, the SM-clock stamps of the first stages / tiles per role relative to the first event, and the
steady-state period of every role (clk per stage / per tile)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("DMVS_WS2_DBG", "1")
from diffmvs_b200 import _cabi, ops, packing
import math

sys.argv_saved = sys.argv
from importlib import import_module
LAYERS = None
src = open(os.path.join(ROOT, "tools", "bench_conv.py")).read()
ns = {}
exec(src[src.index("LAYERS = ["):src.index("REPS =")], ns)
LAYERS = ns["LAYERS"]
ROLES = ["P issued", "S landed", "S split", "M ready", "M issued", "E accready", "E stored", "M accfree", "E halo ld", "E halo bar", "E item0 ld", "E item0 st"]

for name, N, cin, cout, k, s, dims in LAYERS:
    if not any(o in name for o in sys.argv[1].split(",")):
        continue
    g = torch.Generator().manual_seed(0)
    three_d = k[0] > 1 or dims[0] > 1
    cin_st = 4 if cin == 3 else cin        # RGB images are staged with a zero fourth channel (pipeline.py)
    shape = (N, *dims, cin_st) if three_d else (N, dims[1], dims[2], cin_st)
    x = torch.rand(*shape, generator=g)
    if cin_st != cin:
        x[..., cin:] = 0
    x = x.cuda()
    wshape = (cout, cin, *k) if three_d else (cout, cin, k[1], k[2])
    w = (torch.rand(*wshape, generator=g) - 0.5) / math.sqrt(cin * k[0] * k[1] * k[2])
    pc = packing.pack_weight(w, torch.zeros(cout), pad_cin=cin_st if cin_st != cin else 0).to("cuda")
    ops.set_precision(os.environ.get("DMVS_TIMELINE_MODE", "ws2_tf32x3"))
    y = ops.conv(x, pc, stride=s, act=ops.ACT_RELU)
    ops.conv(x, pc, stride=s, act=ops.ACT_RELU, out=y)
    torch.cuda.synchronize()
    buf = (C.c_int64 * 768)()
    n = _cabi.lib().dmvs_conv_ws2_timeline(buf, 768)
    if n <= 0:
        print("timeline facility is off (DMVS_WS2_DBG=1?)")
        sys.exit(1)
    t = [[buf[r * 64 + i] for i in range(64)] for r in range(12)]
    t0 = min(v for row in t for v in row if v > 0)
    print(f"== {name}")
    for r, row in enumerate(t):
        vals = [v - t0 for v in row if v > 0]
        if not vals:
            continue
        per = (vals[-1] - vals[len(vals) // 2]) / max(1, len(vals) - 1 - len(vals) // 2) if len(vals) > 4 else float("nan")
        print(f"  {ROLES[r]:11s} n={len(vals):2d} period {per:8.0f} clk | " + " ".join(f"{v:6d}" for v in vals[:20]))
