#!/usr/bin/env python
"""Micro-benchmark of dmvs_conv_f32 on the layer shapes that dominate cfg3 (CUDA-event timing)."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import ops, packing

# (name, N, cin, cout, k, stride, H, W)  -- H,W are INPUT sizes
LAYERS = [
    ("feat.conv0.0 3->8", 7, 3, 8, 3, 1, 1152, 1600),
    ("feat.conv0.1 8->8", 7, 8, 8, 3, 1, 1152, 1600),
    ("feat.conv1.0 8->16 5x5s2", 7, 8, 16, 5, 2, 1152, 1600),
    ("feat.conv1.1 16->16", 7, 16, 16, 3, 1, 576, 800),
    ("feat.conv2.0 16->32 5x5s2", 7, 16, 32, 5, 2, 576, 800),
    ("feat.conv2.1 32->32", 7, 32, 32, 3, 1, 288, 400),
    ("feat.conv3.0 32->64 5x5s2", 7, 32, 64, 5, 2, 288, 400),
    ("feat.conv3.1 64->64", 7, 64, 64, 3, 1, 144, 200),
    ("feat.inner2 16->64 1x1", 7, 16, 64, 1, 1, 576, 800),
    ("feat.out3 64->16", 7, 64, 16, 3, 1, 576, 800),
    ("feat.out2 64->32", 7, 64, 32, 3, 1, 288, 400),
    ("unet2.init 64->16 7x7", 1, 64, 16, 7, 1, 288, 400),
    ("unet3.init 32->8 7x7", 1, 32, 8, 7, 1, 576, 800),
    ("enc3 16->16", 1, 16, 16, 3, 1, 576, 800),
    ("unet 32->32 @1/8", 1, 32, 32, 3, 1, 144, 200),
    ("mask3 16->64", 1, 16, 64, 3, 1, 576, 800),
]
only = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "all" else None
if len(sys.argv) > 2:
    ops.set_precision(sys.argv[2])
print("precision:", ops.get_precision())
print(f"{'layer':28s} {'ms':>8s} {'TFLOP/s':>8s} {'GB/s(io)':>9s}")
for name, N, cin, cout, k, s, H, W in LAYERS:
    if only and only not in name:
        continue
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, H, W, cin, generator=g).cuda()
    w = (torch.rand(cout, cin, k, k, generator=g) - 0.5) / math.sqrt(cin * k * k)
    pc = packing.pack_weight(w, torch.zeros(cout)).to("cuda")
    y = ops.conv(x, pc, stride=s, act=ops.ACT_RELU)
    for _ in range(2):
        ops.conv(x, pc, stride=s, act=ops.ACT_RELU, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv(x, pc, stride=s, act=ops.ACT_RELU, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * y.numel() * cin * k * k
    io = (x.numel() + y.numel()) * 4
    print(f"{name:28s} {ms:8.3f} {flops / ms / 1e9:8.2f} {io / ms / 1e6:9.1f}")
