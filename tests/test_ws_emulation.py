"""CPU emulation of the index arithmetic of the width-stacked tcgen05 convolution (csrc/conv_ws.cu).

The kernel cannot run without a GPU, but everything that is easy to get wrong in it is integer bookkeeping:
the planar-by-channel-quad operand layout, the host-packed weight slabs (`w_ws`), the
kernel-row descriptor offsets, the N = KW*CC column order, the 136-row staging ring and the shift-add
windows of the epilogue.  This test replays exactly that bookkeeping with numpy (one "MMA" = one matmul of a
128-row operand slice) and checks the result against `F.conv2d`.  It mirrors the kernel's variable names.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffmvs_b200 import packing

RING = 136


def emulate_ws_conv(x, pc, TH, TW, CC, CCE, co_base=0):
    """x [H,W,Cin] (numpy), pc PackedConv (2-D) -> y [H,W,Cout] via the kernel's data movement."""
    H, W, Cin = x.shape
    KD, KH, KW = pc.k
    assert KD == 1
    Cout = pc.cout
    pad_h, pad_w = KH // 2, KW // 2
    cin_pad = (Cin + 7) & ~7
    cout_pad = (Cout + 15) & ~15
    qtot = cin_pad // 4
    taps = KH * KW
    N = (KW * CC + 15) & ~15
    # this launch's packed slabs: the chunks before co_base come first in w_ws (packing.pack_ws)
    w_ws = pc.w_ws.numpy()
    nchunks = cin_pad // 8
    cc_max = packing.ws_cc_max(KW)
    w_off, rem, base = 0, (Cout + 7) & ~7, 0
    while base < co_base:
        cc = min(rem, cc_max)
        w_off += 2 * KD * nchunks * KH * 2 * ((KW * cc + 15) & ~15) * 4
        base += cc
        rem -= cc
    assert base == co_base and CC == min(rem, cc_max), "test case must follow the kernel's channel chunking"
    wslab_f = KH * 2 * N * 4
    w_plane = KD * nchunks * wslab_f
    in_rows, in_cols = TH + KH - 1, TW + KW - 1
    m_total = TH * in_cols
    n_blk = -(-m_total // 128)
    plane = (n_blk * 128 + (KH - 1) * in_cols + 8 + 7) & ~7
    SP = KW * CCE + 4
    y = np.full((H, W, Cout), np.nan, dtype=np.float64)
    rng = np.random.default_rng(0)
    for ty0 in range(0, H, TH):
        for tx0 in range(0, W, TW):
            E = np.zeros((n_blk * 128, N))
            for chunk in range(cin_pad // 8):
                c0 = chunk * 8
                # ---- issue_loads: raw tile, planar by channel quad; slack keeps garbage -------------------
                A = rng.standard_normal((2, plane, 4)) * 1e3     # garbage where the kernel does not write
                for row in range(in_rows):
                    iy = ty0 - pad_h + row
                    for col in range(in_cols):
                        ix = tx0 - pad_w + col
                        for q in range(2):
                            ch = c0 + q * 4
                            ok = 0 <= iy < H and 0 <= ix < W and ch < Cin
                            v = np.zeros(4)
                            if ok:
                                seg = x[iy, ix, ch:ch + 4]
                                v[:len(seg)] = seg
                            A[q, row * in_cols + col] = v
                # ---- weight slab [kh][quad][N][4] from the packed global layout ---------------------------
                src = w_off + (0 * nchunks + chunk) * wslab_f
                slab = (w_ws[src:src + wslab_f] + w_ws[src + w_plane:src + w_plane + wslab_f]).astype(np.float64)  # hi + lo
                slab = slab.reshape(KH, 2, N, 4)
                # ---- MMAs: per block and kernel row one instruction, N columns ---------------------------
                for blk in range(n_blk):
                    for kh in range(KH):
                        a_off = blk * 128 + kh * in_cols
                        a_op = np.concatenate([A[0, a_off:a_off + 128], A[1, a_off:a_off + 128]], axis=1)   # [128][8]
                        b_op = np.concatenate([slab[kh, 0], slab[kh, 1]], axis=1)                          # [N][8]
                        E[blk * 128:(blk + 1) * 128] += a_op @ b_op.T
            # ---- shift-add epilogue through the ring --------------------------------------------------------
            N4 = CCE // 4
            for e0 in range(0, CC, CCE):
                if co_base + e0 >= Cout:
                    break
                ring = np.full((RING, SP), np.nan)
                for blk in range(n_blk):
                    for m in range(128):                      # phase A
                        rrow = (blk * 128 + m) % RING
                        for g in range(KW * (CCE // 8)):
                            kw = g // (CCE // 8)
                            sub = g - kw * (CCE // 8)
                            col = kw * CC + e0 + sub * 8
                            ring[rrow, kw * CCE + sub * 8: kw * CCE + sub * 8 + 8] = E[blk * 128 + m, col:col + 8]
                    for m in range(128):                      # phase B
                        p = blk * 128 - (KW - 1) + m
                        if p < 0:
                            continue
                        py = int((np.float32(p) + np.float32(0.5)) * np.float32(1.0 / in_cols))
                        px = p - py * in_cols
                        oy, ox = ty0 + py, tx0 + px
                        if px >= TW or py >= TH or oy >= H or ox >= W:
                            continue
                        for q4 in range(N4):
                            cq = co_base + e0 + q4 * 4
                            if cq >= Cout:
                                continue
                            v = np.zeros(4)
                            for kw in range(KW):
                                v += ring[(p + kw) % RING, kw * CCE + q4 * 4: kw * CCE + q4 * 4 + 4]
                            for k in range(4):
                                if cq + k < Cout:
                                    assert np.isnan(y[oy, ox, cq + k]), "output written twice"
                                    y[oy, ox, cq + k] = v[k]
    return y


CASES = [
    # cin, cout, (kh,kw), H, W, TH, TW, CC, CCE
    (8, 8, (3, 3), 20, 37, 8, 16, 8, 8),
    (16, 16, (3, 3), 17, 40, 4, 30, 16, 16),
    (12, 20, (3, 3), 16, 24, 8, 24, 24, 8),
    (8, 16, (7, 7), 18, 30, 4, 26, 16, 8),
    (8, 40, (1, 5), 9, 33, 2, 33, 40, 8),
    (8, 20, (5, 1), 12, 20, 4, 20, 24, 8),
    (16, 32, (3, 3), 10, 22, 2, 22, 32, 16),
]


@pytest.mark.parametrize("cin,cout,k,H,W,TH,TW,CC,CCE", CASES)
def test_ws_bookkeeping_matches_conv2d(cin, cout, k, H, W, TH, TW, CC, CCE):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(1, cin, H, W, generator=g) - 0.5
    w = (torch.rand(cout, cin, *k, generator=g) - 0.5) / math.sqrt(cin * k[0] * k[1])
    pc = packing.pack_weight(w, None)
    ref = F.conv2d(x.double(), w.double(), padding=(k[0] // 2, k[1] // 2))[0].permute(1, 2, 0).numpy()
    got = emulate_ws_conv(x[0].permute(1, 2, 0).double().numpy(), pc, TH, TW, CC, CCE)
    assert not np.isnan(got).any(), "some output was never written"
    # hi + lo of the packed weights reproduces the fp32 weights to ~2^-22 relative
    assert np.abs(got - ref).max() < 1e-6


def test_ws_output_channel_chunks():
    """Cout = 80 with a 3-wide kernel runs as two launches (64 + 16 channels) over consecutive `w_ws` chunks."""
    g = torch.Generator().manual_seed(2)
    cin, cout, k, H, W = 8, 80, (3, 3), 9, 20
    x = torch.rand(1, cin, H, W, generator=g) - 0.5
    w = (torch.rand(cout, cin, *k, generator=g) - 0.5) / math.sqrt(cin * 9)
    pc = packing.pack_weight(w, None)
    ref = F.conv2d(x.double(), w.double(), padding=1)[0].permute(1, 2, 0).numpy()
    xin = x[0].permute(1, 2, 0).double().numpy()
    lo = emulate_ws_conv(xin, pc, 4, 20, 64, 16, co_base=0)
    hi = emulate_ws_conv(xin, pc, 4, 20, 16, 16, co_base=64)
    assert np.isnan(lo[..., 64:]).all() and np.isnan(hi[..., :64]).all()
    got = np.where(np.isnan(lo), hi, lo)
    assert np.abs(got - ref).max() < 1e-6
