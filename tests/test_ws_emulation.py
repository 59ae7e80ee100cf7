"""CPU emulation of the index arithmetic of the width-stacked tcgen05 convolution (csrc/conv_ws.cu).

The kernel cannot run without a GPU, but everything that is easy to get wrong in it is integer bookkeeping:
the planar-by-channel-quad operand layout, the host-packed weight slabs (`w_ws`), the
kernel-row descriptor offsets, the N = KW*CC column order and the shuffle / halo shift-add of the epilogue.  This test replays exactly that bookkeeping with numpy (one "MMA" = one matmul of a
128-row operand slice) and checks the result against `F.conv2d`.  It mirrors the kernel's variable names.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffmvs_b200 import packing

def emulate_ws_conv(x, pc, TH, TW, CC, co_base=0, S=1, pair=False, pad=None, out_hw=None):
    """x [H,W,Cin] (numpy), pc PackedConv (2-D) -> y [Ho,Wo,Cout] via the kernel's data movement (stride S)."""
    H, W, Cin = x.shape
    KD, KH_real, KW_real = pc.k
    assert KD == 1
    Cout = pc.cout
    pad_h, pad_w = (KH_real // 2, KW_real // 2) if pad is None else pad
    Ho, Wo = (H + 2 * pad_h - KH_real) // S + 1, (W + 2 * pad_w - KW_real) // S + 1
    if out_hw is not None:       # phase launch (`explicit_extent`): the output size is given, padding is one-sided
        Ho, Wo = out_hw
    smin_h, KH = packing.ws_extent(KH_real, pad_h, S)      # below, KH / KW are the extents in phase-plane shifts
    smin_w, KW = packing.ws_extent(KW_real, pad_w, S)
    cin_pad = (Cin + 7) & ~7
    cout_pad = (Cout + 15) & ~15
    qtot = cin_pad // 4
    taps = KH * KW
    N = (KW * CC + 15) & ~15
    # this launch's packed slabs: the chunks before co_base come first in w_ws (packing.pack_ws)
    w_ws = (pc.w_ws if S == 1 else packing.pack_ws_from_packed(pc.w, Cout, S, (pad_h, pad_w))).numpy()
    # paired kernel rows (conv_ws2.cu, <= 4 input channels): KHm row-MMAs per stage, K = (rows 2j, 2j+1) x 4 channels
    KHm = (KH + 1) // 2 if pair else KH
    if pair:
        assert S == 1 and Cin <= 4 and pc.w_ws_pair is not None
        w_ws = pc.w_ws_pair.numpy()
    nchunks = cin_pad // 8
    nphase = S * S
    cc_max = packing.ws_cc_max(KW)
    w_off, rem, base = 0, (Cout + 7) & ~7, 0
    while base < co_base:
        cc = min(rem, cc_max)
        w_off += 2 * KD * nphase * nchunks * KHm * 2 * ((KW * cc + 15) & ~15) * 4
        base += cc
        rem -= cc
    assert base == co_base and CC == min(rem, cc_max), "test case must follow the kernel's channel chunking"
    wslab_f = KHm * 2 * N * 4
    w_plane = KD * nphase * nchunks * wslab_f
    KHg = 2 * KHm if pair else KH                              # staged rows beyond the tile height, + 1
    in_rows, in_cols = TH + KHg - 1, TW + KW - 1
    m_total = TH * in_cols
    n_blk = -(-m_total // 128)
    plane = (n_blk * 128 + (KHg - 1) * in_cols + 8 + 7) & ~7
    y = np.full((Ho, Wo, Cout), np.nan, dtype=np.float64)
    rng = np.random.default_rng(0)
    for ty0 in range(0, Ho, TH):
        for tx0 in range(0, Wo, TW):
            E = np.zeros((n_blk * 128, N))
            for phase, chunk in [(ph, ck) for ph in range(nphase) for ck in range(cin_pad // 8)]:
                pa, pb = (phase >> 1, phase & 1) if S == 2 else (0, 0)
                c0 = chunk * 8
                # ---- issue_loads: raw tile, planar by channel quad; slack keeps garbage -------------------
                A = rng.standard_normal((2, plane, 4)) * 1e3     # garbage where the kernel does not write
                for row in range(in_rows):
                    iy = S * (ty0 + smin_h) + pa + S * row
                    for col in range(in_cols):
                        ix = S * (tx0 + smin_w) + pb + S * col
                        for q in range(1 if pair else 2):
                            ch = c0 + q * 4
                            ok = 0 <= iy < H and 0 <= ix < W and ch < Cin
                            v = np.zeros(4)
                            if ok:
                                seg = x[iy, ix, ch:ch + 4]
                                v[:len(seg)] = seg
                            A[q, row * in_cols + col] = v
                # ---- weight slab [kh][quad][N][4] from the packed global layout ---------------------------
                src = w_off + ((0 * nphase + phase) * nchunks + chunk) * wslab_f
                slab = (w_ws[src:src + wslab_f] + w_ws[src + w_plane:src + w_plane + wslab_f]).astype(np.float64)  # hi + lo
                slab = slab.reshape(KHm, 2, N, 4)
                # ---- MMAs: per block and kernel row one instruction, N columns ---------------------------
                for blk in range(n_blk):
                    for khp in range(KHm if pair else 0):
                        # the A descriptor's leading-dimension offset is one tile row: K quad 1 = quad 0 one row down
                        a_off = blk * 128 + 2 * khp * in_cols
                        a_op = np.concatenate([A[0, a_off:a_off + 128], A[0, a_off + in_cols:a_off + in_cols + 128]], axis=1)
                        b_op = np.concatenate([slab[khp, 0], slab[khp, 1]], axis=1)
                        E[blk * 128:(blk + 1) * 128] += a_op @ b_op.T
                    for kh in range(0 if pair else KH):
                        tap_row = S * (kh + smin_h) + pa + pad_h
                        if not 0 <= tap_row < KH_real:          # this phase has no kernel row behind shift kh
                            continue
                        a_off = blk * 128 + kh * in_cols
                        a_op = np.concatenate([A[0, a_off:a_off + 128], A[1, a_off:a_off + 128]], axis=1)   # [128][8]
                        b_op = np.concatenate([slab[kh, 0], slab[kh, 1]], axis=1)                          # [N][8]
                        E[blk * 128:(blk + 1) * 128] += a_op @ b_op.T
            # ---- shift-add epilogue: lane = accumulator row, shuffles inside a warp, halo across warps ---------
            ncg = CC // 8
            n_items = n_blk * ncg
            halo = {}
            for it in range(n_items):                          # pass 1
                blk, cg = divmod(it, ncg)
                for quadrant in range(4):
                    for kw in range(1, KW):
                        for lane in range(kw):
                            row = blk * 128 + quadrant * 32 + lane
                            halo[(it, quadrant, kw, lane)] = E[row, kw * CC + cg * 8: kw * CC + cg * 8 + 8].copy()
            for it in range(n_items):                          # pass 2
                blk, cg = divmod(it, ncg)
                c0 = co_base + cg * 8
                for quadrant in range(4):
                    have_next = quadrant < 3 or blk + 1 < n_blk
                    itn, qn = (it if quadrant < 3 else it + ncg), (quadrant + 1) & 3
                    for lane in range(32):
                        p = blk * 128 + quadrant * 32 + lane
                        acc = E[p, cg * 8: cg * 8 + 8].copy()
                        for kw in range(1, KW):
                            if lane + kw < 32:                  # __shfl_down_sync(v, kw): the row of lane + kw
                                v = E[p + kw, kw * CC + cg * 8: kw * CC + cg * 8 + 8]
                            elif have_next:
                                v = halo[(itn, qn, kw, lane + kw - 32)]
                            else:
                                v = np.zeros(8)
                            acc = acc + v
                        py = int((np.float32(p) + np.float32(0.5)) * np.float32(1.0 / in_cols))
                        px = p - py * in_cols
                        oy, ox = ty0 + py, tx0 + px
                        if not (px < TW and py < TH and oy < Ho and ox < Wo and c0 < Cout):
                            continue
                        for k in range(8):
                            if c0 + k < Cout:
                                assert np.isnan(y[oy, ox, c0 + k]), "output written twice"
                                y[oy, ox, c0 + k] = acc[k]
    return y


CASES = [
    # cin, cout, (kh,kw), H, W, TH, TW, CC
    (8, 8, (3, 3), 20, 37, 8, 16, 8),
    (16, 16, (3, 3), 17, 40, 4, 30, 16),
    (12, 20, (3, 3), 16, 24, 8, 24, 24),
    (8, 16, (7, 7), 18, 30, 4, 26, 16),
    (8, 40, (1, 5), 9, 33, 2, 33, 40),
    (8, 20, (5, 1), 12, 20, 4, 20, 24),
    (16, 32, (3, 3), 10, 22, 2, 22, 32),
    (8, 8, (3, 3), 40, 70, 8, 62, 8),
]


@pytest.mark.parametrize("cin,cout,k,H,W,TH,TW,CC", CASES)
def test_ws_bookkeeping_matches_conv2d(cin, cout, k, H, W, TH, TW, CC):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(1, cin, H, W, generator=g) - 0.5
    w = (torch.rand(cout, cin, *k, generator=g) - 0.5) / math.sqrt(cin * k[0] * k[1])
    pc = packing.pack_weight(w, None)
    ref = F.conv2d(x.double(), w.double(), padding=(k[0] // 2, k[1] // 2))[0].permute(1, 2, 0).numpy()
    got = emulate_ws_conv(x[0].permute(1, 2, 0).double().numpy(), pc, TH, TW, CC)
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
    lo = emulate_ws_conv(xin, pc, 4, 20, 64, co_base=0)
    hi = emulate_ws_conv(xin, pc, 4, 20, 16, co_base=64)
    assert np.isnan(lo[..., 64:]).all() and np.isnan(hi[..., :64]).all()
    got = np.where(np.isnan(lo), hi, lo)
    assert np.abs(got - ref).max() < 1e-6


STRIDED = [
    # cin, cout, k, H, W, TH, TW, CC
    (8, 16, (5, 5), 20, 36, 4, 18, 16),
    (16, 32, (5, 5), 16, 24, 2, 12, 32),
    (8, 16, (3, 3), 18, 30, 4, 15, 16),
    (8, 16, (2, 2), 16, 24, 4, 12, 16),
]


@pytest.mark.parametrize("cin,cout,k,H,W,TH,TW,CC", STRIDED)
def test_ws_stride2_phases_match_conv2d(cin, cout, k, H, W, TH, TW, CC):
    """Stride 2 = four stride-1 phases over decimated input planes accumulating into the same columns."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, cin, H, W, generator=g) - 0.5
    w = (torch.rand(cout, cin, *k, generator=g) - 0.5) / math.sqrt(cin * k[0] * k[1])
    pc = packing.pack_weight(w, None)
    ref = F.conv2d(x.double(), w.double(), stride=2, padding=(k[0] // 2, k[1] // 2))[0].permute(1, 2, 0).numpy()
    got = emulate_ws_conv(x[0].permute(1, 2, 0).double().numpy(), pc, TH, TW, CC, S=2)
    assert got.shape == ref.shape
    assert not np.isnan(got).any(), "some output was never written"
    assert np.abs(got - ref).max() < 1e-6


PAIR_CASES = [
    # cin, cout, (kh,kw), H, W, TH, TW, CC
    (3, 8, (3, 3), 21, 40, 8, 30, 8),
    (4, 8, (3, 3), 16, 35, 16, 30, 8),
    (4, 16, (5, 5), 18, 30, 4, 28, 16),
    (1, 8, (3, 3), 12, 31, 4, 30, 8),
]


@pytest.mark.parametrize("cin,cout,k,H,W,TH,TW,CC", PAIR_CASES)
def test_paired_rows_bookkeeping_matches_conv2d(cin, cout, k, H, W, TH, TW, CC):
    """<= 4 input channels: kernel rows paired along K (`w_ws_pair`, one staged channel quad, A leading offset = one row)."""
    g = torch.Generator().manual_seed(2)
    x = torch.rand(1, cin, H, W, generator=g) - 0.5
    w = (torch.rand(cout, cin, *k, generator=g) - 0.5) / math.sqrt(cin * k[0] * k[1])
    pc = packing.pack_weight(w, None)
    assert pc.w_ws_pair is not None
    pad = (k[0] // 2, k[1] // 2)
    ref = F.conv2d(x.double(), w.double(), padding=pad)[0].permute(1, 2, 0).numpy()
    got = emulate_ws_conv(x[0].permute(1, 2, 0).double().numpy(), pc, TH, TW, CC, pair=True)
    assert not np.isnan(got).any(), "some output was never written"
    assert np.abs(got - ref).max() < 1e-6


def test_paired_rows_two_channel_chunks():
    """Cout = 72 runs as two launches (64 + 8 channels) over consecutive `w_ws_pair` chunks."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 4, 9, 33, generator=g) - 0.5
    w = (torch.rand(72, 4, 3, 3, generator=g) - 0.5) / 6.0
    pc = packing.pack_weight(w, None)
    ref = F.conv2d(x.double(), w.double(), padding=1)[0].permute(1, 2, 0).numpy()
    xin = x[0].permute(1, 2, 0).double().numpy()
    a = emulate_ws_conv(xin, pc, 4, 30, 64, co_base=0, pair=True)
    b = emulate_ws_conv(xin, pc, 4, 30, 8, co_base=64, pair=True)
    got = np.where(np.isnan(a), b, a)
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() < 1e-6


def test_fp16_correction_planes_hold_hi_and_residual():
    """`w_ws16` (DMVS_PREC_WS2_TF32_F16C): same hi plane as `w_ws`; the lo plane holds, per (kernel row, column n), two
    16-byte units of 8 halves - unit 0 = fp16(hi) of the chunk's input channels 0..7, unit 1 = fp16(w - hi)."""
    g = torch.Generator().manual_seed(5)
    cin, cout, kh, kw = 16, 16, 3, 3
    w = (torch.rand(cout, cin, kh, kw, generator=g) - 0.5) / 12.0
    pc = packing.pack_weight(w, None)
    a, b = pc.w_ws.numpy(), pc.w_ws16
    n_half = a.size // 2
    assert np.array_equal(a[:n_half], b[:n_half].numpy())                 # identical hi planes
    N = (kw * cout + 15) & ~15
    hi = a[:n_half].reshape(1, 1, cin // 8, kh, 2, N, 4)                  # [kd][phase][chunk][kh][quad][n][4]
    c16 = b[n_half:].contiguous().view(torch.float16).numpy().astype(np.float64).reshape(1, 1, cin // 8, kh, 2, N, 8)
    hi8 = np.concatenate([hi[..., 0, :, :], hi[..., 1, :, :]], axis=-1)   # channels 0..7 of the chunk per (kh, n)
    assert np.abs(c16[..., 0, :, :] * 16 - hi8).max() <= 16 * 3.0e-8       # fp16(hi / 16): exact but for subnormal weights
    full = hi8 + c16[..., 1, :, :] / 16           # hi + fp16(16 (w - hi)) / 16 reproduces w to 2^-21 |w| (or 2^-29 absolute)
    ref = np.zeros_like(full)
    for chunk in range(cin // 8):
        for r in range(kh):
            for t in range(kw):
                ref[0, 0, chunk, r, t * cout:(t + 1) * cout, :] = w[:, chunk * 8:(chunk + 1) * 8, r, t].numpy()
    assert np.abs(full - ref).max() <= np.abs(ref).max() * 2.0 ** -21 + 2.0 ** -25


def test_fp16_correction_product_error_bound():
    """x*w ~ trunc_tf32(x)*rna_tf32(w) [TF32 MMA] + fp16(x - trunc)*fp16(w_hi) + fp16(trunc)*fp16(w - w_hi) [one fp16 MMA]:
    every factor keeps 11 significant bits, the dropped lo*lo term is 2^-21 relative - the 3xTF32 class."""
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(200000, generator=g) - 0.5) * 8.0
    w = (torch.rand(200000, generator=g) - 0.5) * 0.5
    x_hi = (x.view(torch.int32) & -8192).view(torch.float32)               # 0xffffe000: what the tensor core reads
    w_hi = packing.rna_tf32(w)
    main = x_hi.double() * w_hi.double()
    corr = ((x - x_hi) * 16).half().double() * (w_hi / 16).half().double() + (x_hi / 16).half().double() * ((w - w_hi) * 16).half().double()
    exact = x.double() * w.double()
    err = (main + corr - exact).abs() / (exact.abs() + 1e-30)
    # relative to the SCALE of the products (what a sum of them sees) the worst case is 2^-21; the mean relative error
    # per product is 2^-23.8 (the two products are balanced by 16 / (1/16) to stay clear of fp16's subnormal range)
    assert float((main + corr - exact).abs().max()) < 2.0 ** -21 * float(exact.abs().max())
    assert float(err.mean()) < 2.0 ** -23
    plain = (x_hi.double() * w_hi.double() - x.double() * w.double()).abs() / (x.double().abs() * w.double().abs() + 1e-30)
    assert float(plain.mean()) > 50 * float(err.mean())                   # what one TF32 pass alone would give


@pytest.mark.parametrize("py", [0, 1])
def test_phase_launch_bookkeeping(py):
    """ops.conv_up2's launches: KH = 2 x KW = 3 kernel, padding (1 - py, 1), output height = input height (the row
    below the image is the TMA's zero fill).  The emulation's tile loader already zero-fills everything outside."""
    g = torch.Generator().manual_seed(7)
    C, Cout, H, W = 8, 8, 9, 33
    x = torch.rand(1, C, H, W, generator=g) - 0.5
    w3 = torch.rand(Cout, C, 3, 3, generator=g) - 0.5
    pc = packing.pack_up2_phases(w3)[py]
    wp = pc.w[0].permute(3, 2, 0, 1)[:pc.cout, :pc.cin].contiguous()          # [2*Cout, C, 2, 3]
    ref = F.conv2d(F.pad(x.double(), (1, 1, 1 - py, py)), wp.double())[0].permute(1, 2, 0).numpy()
    got = emulate_ws_conv(x[0].permute(1, 2, 0).double().numpy(), pc, 4, 30, 16, pad=(1 - py, 1), out_hw=(H, W))
    assert got.shape == ref.shape and not np.isnan(got).any()
    assert np.abs(got - ref).max() < 1e-6
