#!/usr/bin/env python
"""Micro-benchmark of dmvs_conv_f32 on the layer shapes of cfg3 (CUDA-event timing, all back ends side by side).

    python tools/bench_conv.py [substring|all] [mode,mode,...]
"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import ops, packing

# (name, N, cin, cout, (kd,kh,kw), stride, (D,H,W) of the input)
LAYERS = [
    ("feat.conv0.0 3->8", 7, 3, 8, (1, 3, 3), 1, (1, 1152, 1600)),
    ("feat.conv0.1 8->8", 7, 8, 8, (1, 3, 3), 1, (1, 1152, 1600)),
    ("feat.conv1.0 8->16 5x5s2", 7, 8, 16, (1, 5, 5), 2, (1, 1152, 1600)),
    ("feat.conv1.1 16->16", 7, 16, 16, (1, 3, 3), 1, (1, 576, 800)),
    ("feat.conv2.0 16->32 5x5s2", 7, 16, 32, (1, 5, 5), 2, (1, 576, 800)),
    ("feat.conv2.1 32->32", 7, 32, 32, (1, 3, 3), 1, (1, 288, 400)),
    ("feat.conv3.0 32->64 5x5s2", 7, 32, 64, (1, 5, 5), 2, (1, 288, 400)),
    ("feat.conv3.1 64->64", 7, 64, 64, (1, 3, 3), 1, (1, 144, 200)),
    ("feat.inner2 16->64 1x1", 7, 16, 64, (1, 1, 1), 1, (1, 576, 800)),
    ("feat.out3 64->16", 7, 64, 16, (1, 3, 3), 1, (1, 576, 800)),
    ("feat.out2 64->32", 7, 64, 32, (1, 3, 3), 1, (1, 288, 400)),
    ("unet2.init 64->16 7x7", 1, 64, 16, (1, 7, 7), 1, (1, 288, 400)),
    ("unet3.init 32->8 7x7", 1, 32, 8, (1, 7, 7), 1, (1, 576, 800)),
    ("enc3 16->16", 1, 16, 16, (1, 3, 3), 1, (1, 576, 800)),
    ("enc3.out 32->15", 1, 32, 15, (1, 3, 3), 1, (1, 576, 800)),
    ("unet3.rb 8->8", 1, 8, 8, (1, 3, 3), 1, (1, 576, 800)),
    ("unet3.rb 16->8", 1, 16, 8, (1, 3, 3), 1, (1, 576, 800)),
    ("enc2 32->32", 1, 32, 32, (1, 3, 3), 1, (1, 288, 400)),
    ("enc2.out 64->31", 1, 64, 31, (1, 3, 3), 1, (1, 288, 400)),
    ("unet2.rb 16->16", 1, 16, 16, (1, 3, 3), 1, (1, 288, 400)),
    ("unet 32->32 @1/8", 1, 32, 32, (1, 3, 3), 1, (1, 144, 200)),
    ("gru.zr 64->64 1x5", 1, 64, 64, (1, 1, 5), 1, (1, 144, 200)),
    ("gru.q 64->32 5x1", 1, 64, 32, (1, 5, 1), 1, (1, 144, 200)),
    ("mask3 16->64", 1, 16, 64, (1, 3, 3), 1, (1, 576, 800)),
    ("ctx.layer1 16->16", 1, 16, 16, (1, 3, 3), 1, (1, 576, 800)),
    ("pvw 4->8 3d", 6, 4, 8, (3, 3, 3), 1, (48, 144, 200)),
    ("pvw 8->1 3d", 6, 8, 1, (3, 3, 3), 1, (48, 144, 200)),
    ("reg 8->8 3d", 1, 8, 8, (3, 3, 3), 1, (48, 144, 200)),
    ("reg 16->16 3d", 1, 16, 16, (3, 3, 3), 1, (24, 72, 100)),
]
REPS = int(os.environ.get("BENCH_CONV_REPS", "10"))   # 1 under ncu: two launches per (layer, mode)
only = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] != "all" else None
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "tc_tf32x3", "ws_tf32x3"]
print(f"{'layer':28s} " + " ".join(f"{m + ' ms':>13s}" for m in modes) + f" {'best':>10s} {'TFLOP/s':>8s} {'GB/s(io)':>9s}")
for name, N, cin, cout, k, s, dims in LAYERS:
    if only and not any(o in name for o in only):
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
    res = {}
    y = None
    for mode in modes:
        ops.set_precision(mode)
        try:
            y = ops.conv(x, pc, stride=s, act=ops.ACT_RELU)
            for _ in range(2 if REPS > 1 else 0):
                ops.conv(x, pc, stride=s, act=ops.ACT_RELU, out=y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = REPS
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ops.conv(x, pc, stride=s, act=ops.ACT_RELU, out=y)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / reps
        except Exception as e:  # noqa: BLE001
            res[mode] = float("nan")
            print(f"  {name} [{mode}]: {e}")
    best = min((m for m in modes if res[m] == res[m]), key=lambda m: res[m])
    ms = res[best]
    flops = 2.0 * y.numel() * cin * k[0] * k[1] * k[2]
    io = (x.numel() + y.numel()) * 4
    print(f"{name:28s} " + " ".join(f"{res[m]:13.3f}" for m in modes) +
          f" {best:>10s} {flops / ms / 1e9:8.2f} {io / ms / 1e6:9.1f}", flush=True)
