"""CPU checks of the two algebraic rewrites behind FeatureNet's last level (pipeline.FeatureNetPlan):
conv3x3(nearest_x2(x)) as phase-collapsed 2x2 convolutions (packing.pack_up2_phases, ops.conv_up2) and
conv3x3(conv1x1(x) + b) as one 3x3 convolution plus a frame correction (packing.compose_1x1_into_3x3)."""
import torch
import torch.nn.functional as F

from diffmvs_b200 import packing


def _unpack(pc):
    """PackedConv.w [KD,KH,KW,cin_pad,cout_pad] -> [Cout,Cin,KH,KW]"""
    return pc.w[0].permute(3, 2, 0, 1)[:pc.cout, :pc.cin].contiguous()


def test_phase_collapsed_weights_equal_conv_of_upsampled_map():
    g = torch.Generator().manual_seed(0)
    C, Cout, H, W = 5, 3, 6, 7
    x = torch.rand(2, C, H, W, generator=g, dtype=torch.float64) - 0.5
    w3 = torch.rand(Cout, C, 3, 3, generator=g) - 0.5
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w3.double(), padding=1)
    got = torch.full_like(ref, float("nan"))
    for py, pc in enumerate(packing.pack_up2_phases(w3)):
        assert pc.k == (1, 2, 3) and pc.cout == 2 * Cout and pc.cin == C
        xp = F.pad(x, (1, 1, 1 - py, py))                      # rows (y-1, y) for py = 0, (y, y+1) for py = 1
        yp = F.conv2d(xp, _unpack(pc).double())                # [2, 2*Cout, H, W]
        for px in (0, 1):
            got[:, :, py::2, px::2] = yp[:, px * Cout:(px + 1) * Cout]
    assert not torch.isnan(got).any()
    assert (got - ref).abs().max() < 1e-6


def test_composed_1x1_3x3_with_frame_correction():
    g = torch.Generator().manual_seed(1)
    Cin, Cm, Cout, H, W = 4, 6, 3, 5, 8
    x = torch.rand(2, Cin, H, W, generator=g, dtype=torch.float64) - 0.5
    w1, b1 = torch.rand(Cm, Cin, 1, 1, generator=g) - 0.5, torch.rand(Cm, generator=g) - 0.5
    w3 = torch.rand(Cout, Cm, 3, 3, generator=g) - 0.5
    ref = F.conv2d(F.conv2d(x, w1.double(), b1.double()), w3.double(), padding=1)
    w, bias, table = packing.compose_1x1_into_3x3(w3, w1, b1)
    got = F.conv2d(x, w.double(), bias.double(), padding=1)
    assert float(table[1, 1].abs().max()) == 0.0
    for oy in range(H):
        for ox in range(W):
            ry = 0 if oy == 0 else (2 if oy == H - 1 else 1)
            rx = 0 if ox == 0 else (2 if ox == W - 1 else 1)
            got[:, :, oy, ox] += table[ry, rx].double()
    assert (got - ref).abs().max() < 1e-6
    # no bias: no correction
    _, b0, t0 = packing.compose_1x1_into_3x3(w3, w1, None)
    assert float(b0.abs().max()) == 0.0 and float(t0.abs().max()) == 0.0
