#!/usr/bin/env python
"""Print the tile plan of the width-stacked tcgen05 convolution for the cfg3 layer shapes (host only, no GPU)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffmvs_b200 import _cabi

def plan(N, cin, cout, k, stride, dims, gn=False, passes3=True, gen=1):
    d = _cabi.ConvDesc()
    D, H, W = dims
    kd, kh, kw = k
    pd, ph, pw = kd // 2, kh // 2, kw // 2
    d.x = 0x1000; d.w = 0x1000; d.w_ws = 0x1000; d.y = 0x1000; d.bias = 0x1000
    d.N, d.D, d.H, d.W, d.C1, d.C2 = N, D, H, W, cin, 0
    d.x_ps, d.y_ps = cin, cout
    d.KD, d.KH, d.KW, d.stride, d.pad_d, d.pad_h, d.pad_w = kd, kh, kw, stride, pd, ph, pw
    d.Do, d.Ho, d.Wo, d.Cout = (D + 2 * pd - kd) // stride + 1, (H + 2 * ph - kh) // stride + 1, (W + 2 * pw - kw) // stride + 1, cout
    d.precision = _cabi.PREC_WS_TF32X3 if passes3 else _cabi.PREC_WS_TF32
    if gn:
        d.in_stats, d.in_g1, d.in_g0 = 0x1000, 0x1000, 0x1000
    out = (C.c_int32 * 64)()
    fn = _cabi.lib().dmvs_conv_ws_plan if gen == 1 else _cabi.lib().dmvs_conv_ws2_plan
    n = fn(C.byref(d), out, 8)
    return [tuple(out[8 * i:8 * i + 8]) for i in range(max(n, 0))], n

if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    LAYERS = [
        ("feat.conv0.1 8->8", 7, 8, 8, (1, 3, 3), 1, (1, 1152, 1600)),
        ("feat.conv1.0 8->16 5x5s2", 7, 8, 16, (1, 5, 5), 2, (1, 1152, 1600)),
        ("feat.conv1.1 16->16", 7, 16, 16, (1, 3, 3), 1, (1, 576, 800)),
        ("feat.conv2.1 32->32", 7, 32, 32, (1, 3, 3), 1, (1, 288, 400)),
        ("feat.conv3.1 64->64", 7, 64, 64, (1, 3, 3), 1, (1, 144, 200)),
        ("feat.out3 64->16", 7, 64, 16, (1, 3, 3), 1, (1, 576, 800)),
        ("unet2.init 64->16 7x7", 1, 64, 16, (1, 7, 7), 1, (1, 288, 400)),
        ("unet3.rb 8->8", 1, 8, 8, (1, 3, 3), 1, (1, 576, 800)),
        ("unet 32->32 @1/8", 1, 32, 32, (1, 3, 3), 1, (1, 144, 200)),
        ("pvw 4->8 3d", 6, 4, 8, (3, 3, 3), 1, (48, 144, 200)),
        ("feat.inner2 16->64 1x1", 7, 16, 64, (1, 1, 1), 1, (1, 576, 800)),
        ("feat.conv2.0 16->32 5x5s2", 7, 16, 32, (1, 5, 5), 2, (1, 576, 800)),
        ("enc3 16->16", 1, 16, 16, (1, 3, 3), 1, (1, 576, 800)),
        ("gru 64->64 1x5", 1, 64, 64, (1, 1, 5), 1, (1, 144, 200)),
    ]
    print("layer: per launch (CC, N, TH, TW, n_blk, R, ctas, smem)")
    for name, N, cin, cout, k, s, dims in LAYERS:
        print(f"{name:28s}", plan(N, cin, cout, k, s, dims), " ws2:", plan(N, cin, cout, k, s, dims, gen=2)[0])
