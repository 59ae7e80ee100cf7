"""Operator-level parity of the CUDA kernels (through the C ABI) against the CPU oracle / plain torch fp32.

Everything here needs a GPU (`-m gpu`).  References are computed on the CPU in fp32 with stock torch ops
(the same ops the reference model issues); tolerances are stated per test.  fp32 kernels with a
different summation order agree to ~1e-6 relative; integer outputs (floor index) must match exactly.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffmvs_b200 import ops, packing
from oracle import diffmvs_ref as O
from tests.helpers import load_golden, rel_l1

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, lo=-1.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return lo + (hi - lo) * torch.rand(*shape, generator=g)


def _nhwc(x):  # CPU NCHW -> GPU [N,H,W,C]
    return x.permute(0, 2, 3, 1).contiguous().to(DEV)


def _nchw(y):  # GPU [N,H,W,C] -> CPU NCHW
    return y.permute(0, 3, 1, 2).contiguous().cpu()


PREC_TOL = {"fp32": 2e-6, "tf32x3": 4e-6, "tf32": 3e-3, "tc_tf32x3": 2e-5, "tc_tf32": 3e-3, "auto": 2e-5,
            "ws_tf32x3": 2e-5, "ws_tf32": 3e-3, "ws2_tf32x3": 2e-5, "ws2_f16c": 2e-5}   # max-abs error relative to the output scale


def _modes():
    modes = ["fp32", "auto", "ws_tf32x3", "ws_tf32", "ws2_tf32x3", "ws2_f16c"]
    try:                       # the round-1 back ends only exist in a DMVS_BUILD_LEGACY=1 library
        if ops.legacy_backends():
            modes += list(ops.LEGACY_MODES)
    except Exception:
        pass
    return modes


@pytest.fixture(params=_modes())
def precision(request):
    old = ops.get_precision()
    ops.set_precision(request.param)
    yield request.param
    ops.set_precision(old)


def _close(got, ref, tol=2e-6, what=""):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    assert err <= tol * max(scale, 1.0), f"{what}: max abs err {err:.3e} (scale {scale:.3e})"


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # cin, cout, k, stride, H, W
    (3, 8, (3, 3), 1, 40, 72),
    (8, 8, (3, 3), 1, 33, 50),
    (8, 16, (5, 5), 2, 64, 96),
    (16, 32, (5, 5), 2, 32, 64),
    (32, 64, (5, 5), 2, 32, 32),
    (64, 64, (3, 3), 1, 16, 20),
    (64, 48, (1, 1), 1, 16, 20),
    (64, 16, (7, 7), 1, 32, 40),
    (32, 8, (7, 7), 1, 64, 80),
    (24, 31, (3, 3), 1, 32, 40),
    (6, 32, (3, 3), 1, 32, 40),
    (64, 36, (1, 1), 1, 32, 40),
    (64, 144, (1, 1), 1, 16, 24),
    (52, 40, (1, 5), 1, 16, 20),
    (52, 20, (5, 1), 1, 16, 20),
    (16, 2, (1, 1), 1, 32, 40),
    (4, 1, (3, 3), 1, 17, 19),
    # multi-tile / multi-block shapes for the persistent tcgen05 kernels
    (16, 16, (3, 3), 1, 70, 150),
    (8, 8, (3, 3), 1, 100, 260),
    (64, 16, (3, 3), 1, 40, 200),
    (32, 32, (3, 3), 1, 36, 100),
    (64, 64, (1, 5), 1, 18, 50),
    (64, 32, (5, 1), 1, 18, 50),
]


@pytest.mark.parametrize("cin,cout,k,stride,H,W", CONV_CASES)
def test_conv2d_matches_torch(cin, cout, k, stride, H, W, precision):
    N = 2
    x = _rand(N, cin, H, W, seed=1)
    w = _rand(cout, cin, *k, seed=2) * (1.0 / math.sqrt(cin * k[0] * k[1]))
    b = _rand(cout, seed=3)
    pad = (k[0] // 2, k[1] // 2)
    ref = F.relu(F.conv2d(x, w, b, stride=stride, padding=pad))
    pc = packing.pack_weight(w, b).to(DEV)
    y = ops.conv(_nhwc(x), pc, stride=stride, act=ops.ACT_RELU)
    assert tuple(y.shape) == (N, ref.shape[2], ref.shape[3], cout)
    _close(_nchw(y), ref, tol=PREC_TOL[precision], what=f"conv {cin}->{cout} k{k} s{stride} [{precision}]")


def test_conv2d_epilogues_and_views(precision):
    if precision in ("tf32", "tc_tf32", "ws_tf32"):
        pytest.skip("epilogue logic is precision independent; covered by the fp32-class modes")
    N, H, W = 1, 24, 40
    x1, x2 = _rand(N, 16, H, W, seed=1), _rand(N, 8, H, W, seed=2)
    w, b = _rand(20, 24, 3, 3, seed=3) * 0.1, _rand(20, seed=4)
    res = _rand(N, 20, H, W, seed=5)
    pc = packing.pack_weight(w, b).to(DEV)
    conv = F.conv2d(torch.cat((x1, x2), 1), w, b, padding=1)
    _c = lambda got, ref, what: _close(got, ref, tol=PREC_TOL[precision], what=what)
    # virtual concat + residual before ReLU (ResidualBlock)
    y = ops.conv(_nhwc(x1), pc, x2=_nhwc(x2), act=ops.ACT_RELU, res=_nhwc(res), res_mode=ops.RES_PRE_ACT)
    _c(_nchw(y), F.relu(conv + res), "concat+res_pre")
    # residual after ReLU (CostRegNet skip)
    y = ops.conv(_nhwc(x1), pc, x2=_nhwc(x2), act=ops.ACT_RELU, res=_nhwc(res), res_mode=ops.RES_POST_ACT)
    _c(_nchw(y), F.relu(conv) + res, "res_post")
    # activation only from channel 5 on, tanh / sigmoid / silu
    for act, fn in ((ops.ACT_TANH, torch.tanh), (ops.ACT_SIGMOID, torch.sigmoid), (ops.ACT_SILU, F.silu)):
        y = ops.conv(_nhwc(x1), pc, x2=_nhwc(x2), act=act, act_c0=5)
        ref = conv.clone()
        ref[:, 5:] = fn(conv[:, 5:])
        _c(_nchw(y), ref, f"act {act} from c5")
    # channel-sliced input and output inside wider buffers
    wide_in = torch.zeros(N, H, W, 40, device=DEV)
    wide_in[..., 8:24] = _nhwc(x1)
    wide_in[..., 32:40] = _nhwc(x2)
    wide_out = torch.full((N, H, W, 64), 7.0, device=DEV)
    ops.conv(wide_in[..., 8:24], pc, x2=wide_in[..., 32:40], out=wide_out[..., 4:24])
    _c(_nchw(wide_out[..., 4:24]), conv, "sliced io")
    assert torch.all(wide_out[..., :4] == 7.0) and torch.all(wide_out[..., 24:] == 7.0)
    # nearest-upsampled residual (FPN lateral, module.py:409-416)
    small = _rand(N, 20, H // 2, W // 2, seed=6)
    y = ops.conv(_nhwc(x1), pc, x2=_nhwc(x2), res=_nhwc(small), res_mode=ops.RES_PRE_ACT, res_up2=True)
    _c(_nchw(y), conv + F.interpolate(small, scale_factor=2, mode="nearest"), "res_up2")
    # nearest-upsampled input (update.py:38-42)
    xs = _rand(N, 24, H // 2, W // 2, seed=7)
    y = ops.conv(_nhwc(xs), pc, in_up2=True)
    _c(_nchw(y), F.conv2d(F.interpolate(xs, scale_factor=2, mode="nearest"), w, b, padding=1), "in_up2")


def test_conv_unshuffle_equals_reference_downsample():
    x = _rand(2, 8, 32, 48, seed=1)
    w, b = _rand(16, 32, 1, 1, seed=2) * 0.2, _rand(16, seed=3)
    ref = F.conv2d(O.pixel_unshuffle2(x), w, b)
    pc = packing.pack_unshuffle_conv({"d.weight": w, "d.bias": b}, "d").to(DEV)
    y = ops.conv(_nhwc(x), pc, stride=2, pad=(0, 0, 0))
    _close(_nchw(y), ref, what="pixel-unshuffle conv")


def test_conv_bn_folding():
    sd = {"c.conv.weight": _rand(16, 8, 3, 3, seed=1) * 0.2, "c.bn.weight": _rand(16, seed=2, lo=0.5, hi=1.5),
          "c.bn.bias": _rand(16, seed=3), "c.bn.running_mean": _rand(16, seed=4),
          "c.bn.running_var": _rand(16, seed=5, lo=0.5, hi=2.0)}
    x = _rand(1, 8, 20, 36, seed=6)
    ref = O.conv_bn_act(sd, "c", x)
    y = ops.conv(_nhwc(x), packing.pack_conv_bn(sd, "c").to(DEV), act=ops.ACT_RELU)
    _close(_nchw(y), ref, tol=5e-6, what="conv+bn fold")


@pytest.mark.parametrize("cin,cout,stride", [(4, 8, 1), (8, 8, 1), (8, 16, 2), (16, 32, 2), (32, 32, 1), (8, 1, 1)])
def test_conv3d_matches_torch(cin, cout, stride, precision):
    N, D, H, W = 2, 8, 12, 20
    x = _rand(N, cin, D, H, W, seed=1)
    w, b = _rand(cout, cin, 3, 3, 3, seed=2) * (1 / math.sqrt(27 * cin)), _rand(cout, seed=3)
    ref = F.relu(F.conv3d(x, w, b, stride=stride, padding=1))
    y = ops.conv(x.permute(0, 2, 3, 4, 1).contiguous().to(DEV), packing.pack_weight(w, b).to(DEV), stride=stride,
                 act=ops.ACT_RELU)
    _close(y.permute(0, 4, 1, 2, 3).cpu(), ref, tol=PREC_TOL[precision], what=f"conv3d {cin}->{cout} s{stride} [{precision}]")


@pytest.mark.parametrize("cin,cout,k,dims", [
    (3, 8, (1, 3, 3), (1, 100, 260)),      # RGB staged as 4 channels (FeatureNet conv0.0), multi-tile
    (4, 16, (1, 5, 5), (1, 40, 72)),
    (4, 8, (1, 7, 7), (1, 33, 50)),
    (4, 8, (3, 3, 3), (6, 40, 70)),        # PixelViewWeight / CostRegNet first layers
    (1, 8, (3, 3, 3), (5, 20, 33)),
    (4, 72, (1, 3, 3), (1, 24, 40)),       # two output-channel chunks
])
def test_conv_paired_kernel_rows(cin, cout, k, dims):
    """<= 4 input channels on the TMA-fed tcgen05 back end: kernel rows paired along K (`w_ws_pair`)."""
    old = ops.get_precision()
    ops.set_precision("ws2_tf32x3")
    try:
        N = 2
        D, H, W = dims
        three_d = k[0] > 1
        x = _rand(N, cin, D, H, W, seed=11)
        w = _rand(cout, cin, *k, seed=12) * (1.0 / math.sqrt(cin * k[0] * k[1] * k[2]))
        b = _rand(cout, seed=13)
        ref = F.relu(F.conv3d(x, w, b, padding=(k[0] // 2, k[1] // 2, k[2] // 2)))
        cin_st = 4 if cin == 3 else cin                         # ws2 needs 16-byte pixel strides
        xs = torch.zeros(N, D, H, W, cin_st)
        xs[..., :cin] = x.permute(0, 2, 3, 4, 1)
        pc = packing.pack_weight(w if three_d else w[:, :, 0], b, pad_cin=cin_st if cin_st != cin else 0).to(DEV)
        assert pc.w_ws_pair is not None
        xin = xs.to(DEV) if three_d else xs[:, 0].contiguous().to(DEV)
        y = ops.conv(xin, pc, act=ops.ACT_RELU)
        if cin_st % 4 == 0:
            assert ops._LAST_CONV_BACKEND == ops.PREC_WS2_TF32X3
        got = y.cpu() if three_d else y.cpu().unsqueeze(1)
        _close(got.permute(0, 4, 1, 2, 3), ref, tol=PREC_TOL["ws2_tf32x3"], what=f"paired conv {cin}->{cout} k{k}")
    finally:
        ops.set_precision(old)


@pytest.mark.parametrize("N,D,H,W,ps", [(2, 5, 13, 37, 8), (1, 9, 20, 70, 12), (3, 1, 8, 32, 8), (1, 48, 18, 50, 8), (1, 20, 9, 33, 8)])
def test_conv3d_to1_matches_torch(N, D, H, W, ps):
    """Depth-marching Conv3d(8 -> 1): logits, and PixelViewWeight's sigmoid + max over depth (module.py:459-463)."""
    x = _rand(N, 8, D, H, W, seed=21)
    w, b = _rand(1, 8, 3, 3, 3, seed=22) * (1 / math.sqrt(27 * 8)), _rand(1, seed=23)
    ref = F.conv3d(x, w, b, padding=1)[:, 0]
    buf = torch.zeros(N, D, H, W, ps)
    buf[..., :8] = x.permute(0, 2, 3, 4, 1)
    xin = buf.to(DEV)[..., :8]                                   # a channel slice of a wider buffer when ps > 8
    pc = packing.pack_weight(w, b).to(DEV)
    y = ops.conv3d_to1(xin, pc)
    _close(y.cpu(), ref, tol=2e-6, what="conv3d_to1 logits")
    vw = ops.conv3d_to1(xin, pc, sigmoid_max=True)
    _close(vw.cpu(), torch.sigmoid(ref).max(dim=1)[0], tol=2e-6, what="conv3d_to1 sigmoid-max")
    pc0 = packing.pack_weight(w, None).to(DEV)                   # CostRegNet_small.prob has no bias
    _close(ops.conv3d_to1(xin, pc0).cpu(), F.conv3d(x, w, None, padding=1)[:, 0], tol=2e-6, what="conv3d_to1 no bias")


@pytest.mark.parametrize("C,Cout,H,W", [(64, 16, 36, 50), (16, 8, 17, 33), (32, 32, 8, 70)])
def test_conv_up2_matches_conv_of_upsampled_map(C, Cout, H, W):
    """ops.conv_up2: conv3x3(nearest_x2(x)) as two phase launches (one-sided padding, every other output row)."""
    N = 2
    x = _rand(N, C, H, W, seed=31)
    w3 = _rand(Cout, C, 3, 3, seed=32) * (1.0 / math.sqrt(9 * C))
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w3, padding=1)
    pcu = tuple(pc.to(DEV) for pc in packing.pack_up2_phases(w3))
    y = ops.conv_up2(_nhwc(x), pcu)
    _close(_nchw(y), ref, tol=2e-5, what="conv_up2")
    base = _rand(N, Cout, 2 * H, 2 * W, seed=33)
    y2 = _nhwc(base).contiguous()
    ops.conv_up2(_nhwc(x), pcu, out=y2, accumulate=True)
    _close(_nchw(y2), ref + base, tol=2e-5, what="conv_up2 accumulate")


def test_fpn_last_level_composed_matches_layerwise():
    """out3(nearest_x2(intra) + inner2(conv1)) (module.py:415-417): composed / phase-collapsed form vs the layers."""
    N, H, W = 2, 20, 34                                       # quarter-resolution size; conv1 is [N,16,2H,2W]
    intra, c1 = _rand(N, 64, H, W, seed=41), _rand(N, 16, 2 * H, 2 * W, seed=42)
    wi, bi = _rand(64, 16, 1, 1, seed=43) * 0.25, _rand(64, seed=44)
    w3 = _rand(16, 64, 3, 3, seed=45) * (1.0 / 24.0)
    ref = F.conv2d(F.interpolate(intra, scale_factor=2, mode="nearest") + F.conv2d(c1, wi, bi), w3, padding=1)
    w, b, table = packing.compose_1x1_into_3x3(w3, wi, bi)
    y = ops.conv(_nhwc(c1), packing.pack_weight(w, b).to(DEV))
    ops.conv_up2(_nhwc(intra), tuple(pc.to(DEV) for pc in packing.pack_up2_phases(w3)), out=y, accumulate=True)
    ops.border_bias_add(y, table.to(DEV))
    _close(_nchw(y), ref, tol=2e-5, what="composed FPN level")


@pytest.mark.parametrize("cin,cout", [(32, 16), (16, 8)])
def test_deconv3d_matches_torch(cin, cout):
    N, D, H, W = 1, 3, 5, 7
    sd = {"d.conv.weight": _rand(cin, cout, 3, 3, 3, seed=1) * 0.1, "d.bn.weight": _rand(cout, seed=2, lo=0.5, hi=1.5),
          "d.bn.bias": _rand(cout, seed=3), "d.bn.running_mean": _rand(cout, seed=4) * 0.1,
          "d.bn.running_var": _rand(cout, seed=5, lo=0.5, hi=2.0)}
    x = _rand(N, cin, D, H, W, seed=6)
    skip = _rand(N, cout, 2 * D, 2 * H, 2 * W, seed=7)
    ref = skip + O.deconv3d_bn_relu(sd, "d", x)
    w, b = packing.pack_deconv3d_bn(sd, "d")
    y = ops.deconv3d(x.permute(0, 2, 3, 4, 1).contiguous().to(DEV), w.to(DEV), b.to(DEV),
                     skip.permute(0, 2, 3, 4, 1).contiguous().to(DEV))
    _close(y.permute(0, 4, 1, 2, 3).cpu(), ref, tol=5e-6, what="deconv3d")


def test_groupnorm_pipeline_matches_resnet_block(precision):
    if precision in ("tf32", "tc_tf32", "ws_tf32"):
        pytest.skip("covered by the fp32-class modes")
    """conv(+stats) -> conv(GN+SiLU prologue, +stats) -> groupnorm_silu_add == oracle resnet_block."""
    from diffmvs_b200 import pipeline
    dim_in, dim_out, H, W, N = 24, 16, 20, 28, 2
    g = lambda *s, seed: _rand(*s, seed=seed) * 0.3
    sd = {"rb.block1.proj.weight": g(dim_out, dim_in, 3, 3, seed=1), "rb.block1.proj.bias": g(dim_out, seed=2),
          "rb.block1.norm.weight": _rand(dim_out, seed=3, lo=0.5, hi=1.5), "rb.block1.norm.bias": g(dim_out, seed=4),
          "rb.block2.proj.weight": g(dim_out, dim_out, 3, 3, seed=5), "rb.block2.proj.bias": g(dim_out, seed=6),
          "rb.block2.norm.weight": _rand(dim_out, seed=7, lo=0.5, hi=1.5), "rb.block2.norm.bias": g(dim_out, seed=8),
          "rb.res_conv.weight": g(dim_out, dim_in, 1, 1, seed=9), "rb.res_conv.bias": g(dim_out, seed=10),
          "rb.mlp.1.weight": g(2 * dim_out, 64, seed=11), "rb.mlp.1.bias": g(2 * dim_out, seed=12)}
    temb = _rand(1, 64, seed=13)
    x = _rand(N, dim_in, H, W, seed=14)
    ref = O.resnet_block(sd, "rb", x, temb.expand(N, -1))
    plan = pipeline.ResnetBlockPlan(sd, "rb", DEV, temb)
    arena = pipeline.StatsArena(DEV, N, 2)
    xg = _nhwc(x)
    y = plan(xg[..., :16], arena, x2=xg[..., 16:])
    assert rel_l1(_nchw(y), ref) < 5e-6
    _close(_nchw(y), ref, tol=3e-5, what="resnet block")


def test_gru_matches_oracle():
    from diffmvs_b200.models.module import SepConvGRU
    hid, cin, H, W = 20, 32, 18, 25
    gru = SepConvGRU(hid, cin).eval()
    sd = {k: _rand(*v.shape, seed=i) * 0.2 for i, (k, v) in enumerate(gru.state_dict().items())}
    gru.load_state_dict(sd)
    h, x = torch.tanh(_rand(1, hid, H, W, seed=50)), _rand(1, cin, H, W, seed=51)
    ref = O.sep_conv_gru({"g." + k: v for k, v in sd.items()}, "g", h, x)
    out = gru.to(DEV)(h.to(DEV), x.to(DEV))
    _close(out.cpu(), ref, tol=5e-6, what="SepConvGRU")


# ------------------------------------------------------------------------------------------------
# warp / correlation
# ------------------------------------------------------------------------------------------------
def _cameras(B, V, W, H, seed=0, rot=True):
    """[B,V,2,4,4] with mild rotations so that some samples leave the image and z varies."""
    g = torch.Generator().manual_seed(seed)
    P = torch.zeros(B, V, 2, 4, 4)
    for b in range(B):
        for v in range(V):
            E = torch.eye(4)
            if v > 0:
                if rot:
                    a = (torch.rand(3, generator=g) - 0.5) * 0.08
                    Rx = torch.tensor([[1, 0, 0], [0, math.cos(a[0]), -math.sin(a[0])], [0, math.sin(a[0]), math.cos(a[0])]])
                    Ry = torch.tensor([[math.cos(a[1]), 0, math.sin(a[1])], [0, 1, 0], [-math.sin(a[1]), 0, math.cos(a[1])]])
                    Rz = torch.tensor([[math.cos(a[2]), -math.sin(a[2]), 0], [math.sin(a[2]), math.cos(a[2]), 0], [0, 0, 1]])
                    E[:3, :3] = Rz @ Ry @ Rx
                E[:3, 3] = (torch.rand(3, generator=g) - 0.5) * torch.tensor([120.0, 60.0, 20.0])
            K = torch.tensor([[1.8 * W, 0, W / 2], [0, 1.8 * W, H / 2], [0, 0, 1.0]])
            P[b, v, 0] = E
            P[b, v, 1, :3, :3] = K
    return P


def test_warp_matches_golden_reference():
    g = load_golden("cas_tiny")
    t = lambda k: torch.from_numpy(g["tap_" + k])
    from diffmvs_b200.models.module import differentiable_warping
    out = differentiable_warping(t("warp_src").to(DEV), t("warp_src_proj").to(DEV), t("warp_ref_proj").to(DEV),
                                 t("warp_depth").to(DEV))
    assert tuple(out.shape) == tuple(t("warp_out").shape)
    assert rel_l1(out, t("warp_out")) < 1e-5


@pytest.mark.parametrize("C", [16, 32, 48, 5])
def test_warp_volume_matches_oracle_with_out_of_bounds(C):
    B, H, W, D = 2, 24, 40, 6
    P = _cameras(B, 2, W, H, seed=3)
    src = _rand(B, C, H, W, seed=1)
    depth = 400 + 600 * torch.rand(B, D, H, W, generator=torch.Generator().manual_seed(2))
    depth[0, 0, :4] = -50.0         # negative z is not masked by the reference
    ref = O.differentiable_warping(src, O.compose_projection(P[:, 1]), O.compose_projection(P[:, 0]), depth)
    hom = ops.compose_homographies(P.to(DEV))[:, 0].contiguous()
    out = ops.warp_volume(_nhwc(src), hom, depth.to(DEV))
    got = out.permute(0, 4, 1, 2, 3).cpu()
    # homography computed in fp64 here vs fp32 LU in torch: coordinates agree to ~1e-4 px
    assert rel_l1(got, ref) < 2e-4
    frac_oob = (ref == 0).float().mean().item()
    assert 0.0 < frac_oob < 0.9


def test_plane_sweep_aggregate_and_regression_match_oracle():
    B, V, C, G, D, H, W = 1, 4, 48, 4, 8, 16, 20
    P = _cameras(B, V, W, H, seed=5)
    feats = [_rand(B, C, H, W, seed=10 + v) for v in range(V)]
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384).view(1, -1)
    rng = O.DepthRange(dv)
    planes = rng.to_depth((torch.arange(D).float() / (D - 1.0)).view(1, D, 1, 1).repeat(1, 1, H, W))
    ref_proj = O.compose_projection(P[:, 0])
    cors = [O.group_correlation(O.differentiable_warping(feats[v], O.compose_projection(P[:, v]), ref_proj, planes),
                                feats[0], G) for v in range(1, V)]
    hom = ops.compose_homographies(P.to(DEV))
    fs = torch.stack([_nhwc(f) for f in feats], 0)
    cor = ops.plane_sweep_corr(fs, hom, planes[:, :, 0, 0].contiguous().to(DEV), G)
    for v in range(V - 1):
        got = cor[v].permute(3, 0, 1, 2).cpu().unsqueeze(0)       # [1,G,D,H,W]
        assert rel_l1(got, cors[v]) < 2e-4, v
    # aggregation with given weights, using the kernel's own correlation volumes as input
    w = torch.rand(B * (V - 1), H, W, generator=torch.Generator().manual_seed(3))
    vol = ops.aggregate_views(cor, w.to(DEV), B)
    cc = cor.cpu()
    acc, ws = 0, 1e-8
    for v in range(V - 1):
        ws = ws + w[v].view(1, H, W, 1)
        acc = acc + w[v].view(1, H, W, 1) * cc[v]
    _close(vol[0].cpu(), acc / ws, tol=1e-6, what="aggregate")
    # view-weight max
    logit = _rand(3, D, H, W, seed=4) * 3
    _close(ops.view_weight_max(logit.to(DEV)).cpu(), torch.sigmoid(logit).max(1)[0], tol=1e-6, what="view weight max")


@pytest.mark.parametrize("D", [8, 48, 96])
def test_depth_regression_indices_bit_exact(D):
    B, H, W = 2, 36, 50
    logits = _rand(B, D, H, W, seed=D) * 4.0
    logits[0, :, 0, 0] = 0.0                     # uniform -> expected index (D-1)/2
    logits[0, :, 0, 1] = -30.0
    logits[0, D - 1, 0, 1] = 30.0                # one-hot at the last plane
    logits[0, 0, 0, 2] = 40.0                    # one-hot at the first plane
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384).view(1, -1).repeat(B, 1)
    rng = O.DepthRange(dv)
    idx, j, conf = O.depth_regression(logits)
    n_ref = idx / (D - 1.0)
    n, depth, cf, fl = ops.depth_regression(logits.to(DEV), rng.depth_min.view(B).to(DEV), rng.depth_max.view(B).to(DEV),
                                            want_floor=True)
    mism = (fl.cpu().long() != j[:, 0]).float().mean().item()
    assert mism == 0.0, f"floor(expected index) differs on {mism:.2%} of pixels"
    _close(n.cpu(), n_ref[:, 0], tol=2e-6, what="normalised index")
    assert rel_l1(depth, rng.to_depth(n_ref)[:, 0]) < 1e-6
    _close(cf.cpu(), conf[:, 0], tol=2e-6, what="confidence")


@pytest.mark.parametrize("C,D,with_conf,wshift", [(32, 4, False, 1), (32, 6, True, 1), (16, 4, True, 2), (16, 4, False, 0)])
def test_get_cost_matches_oracle(C, D, with_conf, wshift):
    B, V, G, H, W = 2, 3, 4, 32, 48
    P = _cameras(B, V, W, H, seed=7, rot=True)
    feats = [_rand(B, C, H, W, seed=20 + v) for v in range(V)]
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384).view(1, -1).repeat(B, 1)
    rng = O.DepthRange(dv)
    inv = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(1))
    inv[0, 0, :2] = 0.0
    inv[0, 0, 2:4] = 1.0
    conf = torch.rand(B, H, W, generator=torch.Generator().manual_seed(2)) if with_conf else None
    vw_small = torch.rand(B, V - 1, H >> wshift, W >> wshift, generator=torch.Generator().manual_seed(3))
    vw = F.interpolate(vw_small, scale_factor=2 ** wshift, mode="nearest") if wshift else vw_small
    interval = 2.0 / 384
    cost_ref, samp_ref = O.get_cost(inv, feats, P, interval, rng, D, vw, conf, G, 0.125, 8.0)
    fs = torch.stack([_nhwc(f) for f in feats], 0)
    hom = ops.compose_homographies(P.to(DEV))
    cost, samp = ops.get_cost(fs, hom, inv[:, 0].contiguous().to(DEV), None if conf is None else conf.to(DEV),
                              vw_small.to(DEV), rng.depth_min.view(B).to(DEV), rng.depth_max.view(B).to(DEV), G, D, wshift,
                              interval, 0.125, 8.0)
    _close(_nchw(samp), samp_ref, tol=1e-6, what="hypotheses")
    assert rel_l1(_nchw(cost), cost_ref) < 3e-4


@pytest.mark.parametrize("ratio", [2, 4])
def test_upsample_depth_matches_oracle(ratio):
    B, H, W = 2, 18, 26
    n = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(1))
    mask = _rand(B, 9 * ratio * ratio, H, W, seed=2) * 2
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384).view(1, -1).repeat(B, 1)
    rng = O.DepthRange(dv)
    up = O.upsample_depth(n, mask, ratio)
    dep_ref = rng.to_depth(up.unsqueeze(1)).squeeze(1)
    raw, dep, nrm = ops.upsample_depth(n[:, 0].contiguous().to(DEV), _nhwc(mask), rng.depth_min.view(B).to(DEV),
                                       rng.depth_max.view(B).to(DEV), ratio, want="raw+depth+norm")
    _close(raw.cpu(), up, tol=2e-6, what="convex upsample")
    assert rel_l1(dep, dep_ref) < 1e-6
    assert rel_l1(nrm, rng.to_norm(dep_ref.unsqueeze(1)).squeeze(1)) < 1e-5


def test_refine_update_and_ddim_step():
    B, H, W = 2, 10, 12
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384).view(1, -1).repeat(B, 1)
    rng = O.DepthRange(dv)
    inv0 = torch.rand(B, H, W)
    noise = torch.randn(B, H, W)
    delta = torch.empty(B, H, W, device=DEV)
    inv = torch.empty(B, H, W, device=DEV)
    ubuf = torch.zeros(B, H, W, 8, device=DEV)
    ops.refine_update(0, inv0.to(DEV), noise.to(DEV), 1, 0.5, delta, inv, ubuf[..., 7:8], 8)
    ref_inv = (inv0 + 0.5 * noise).clamp(0, 1)
    assert torch.equal(inv.cpu(), ref_inv) and torch.equal(ubuf[..., 7].cpu(), ref_inv)
    assert torch.equal(delta.cpu(), ref_inv - inv0)
    head = torch.randn(B, H, W, 2) * 0.1
    depth = torch.empty(B, H, W, device=DEV)
    ops.refine_update(1, inv0.to(DEV), head.to(DEV), 2, 1.0, delta, inv, ubuf[..., 7:8], 8, rng.depth_min.view(B).to(DEV),
                      rng.depth_max.view(B).to(DEV), depth)
    ref2 = (inv0 + ((ref_inv - inv0) + head[..., 0])).clamp(0, 1)
    assert torch.equal(inv.cpu(), ref2)
    assert rel_l1(depth, rng.to_depth(ref2.unsqueeze(1)).squeeze(1)) < 1e-6
    img = torch.randn(B, H, W)
    img_d = img.to(DEV)
    ops.ddim_step(img_d, delta, noise.to(DEV), 2.0, 1.7, 0.8, 0.3, 0.2, 0.5)
    dl = delta.cpu()
    ref3 = dl * 0.8 + 0.3 * ((2.0 * img - dl) / 1.7) + 0.2 * (0.5 * noise)
    _close(img_d.cpu(), ref3, tol=1e-6, what="ddim step")


def test_layout_transposes_roundtrip():
    x = _rand(2, 5, 7, 9, seed=1)
    y = ops.to_nhwc(x.to(DEV))
    assert torch.equal(y.cpu(), x.permute(0, 2, 3, 1))
    assert torch.equal(ops.to_nchw_dense(y).cpu(), x)
    assert torch.equal(ops.upsample_nearest(x[:, 0].contiguous().to(DEV), 4).cpu(),
                       F.interpolate(x[:, :1], scale_factor=4, mode="nearest")[:, 0])


def test_cpu_tensors_are_rejected():
    pc = packing.pack_weight(_rand(4, 4, 3, 3), None)
    with pytest.raises(ValueError):
        ops.conv(torch.zeros(1, 8, 8, 4), pc)


def test_uint8_images_are_staged_bit_identically():
    """8-bit images (extension): ops.image_to_nhwc4 divides by 255 on the device with the loader's exact fp32 division
    (`np.array(img, dtype=np.float32) / 255.`, datasets/data_io.py:166-170) - planar and interleaved layouts - and the
    whole model gives bit-identical depth maps for uint8 and float32 inputs."""
    rng = np.random.default_rng(0)
    hwc = rng.integers(0, 256, size=(2, 40, 56, 3), dtype=np.uint8)
    ref = torch.from_numpy((hwc.astype(np.float32) / 255.0).astype(np.float32))            # the loader's arithmetic
    u8 = torch.from_numpy(hwc).to(DEV)
    for x in (u8.permute(0, 3, 1, 2), u8.permute(0, 3, 1, 2).contiguous()):                  # interleaved view / planar
        y = ops.image_to_nhwc4(x)
        assert torch.equal(y[..., :3].cpu(), ref) and torch.all(y[..., 3] == 0)
    yf = ops.image_to_nhwc4(ref.permute(0, 3, 1, 2).contiguous().to(DEV))
    assert torch.equal(yf, y)

    from diffmvs_b200 import synth
    from diffmvs_b200.models import CasDiffMVS
    from oracle import spec
    args = synth.workload_args("cas_tiny")
    model = CasDiffMVS(args, test=True)
    model.load_state_dict(synth.synth_state_dict(spec.state_dict_shapes(args), 123), strict=False)
    model.to(DEV).eval()
    imgs, proj, dv = synth.workload_inputs("cas_tiny")
    imgs_u8 = [(i * 255).round().clamp(0, 255).to(torch.uint8) for i in imgs]
    imgs_f = [i.float() / 255.0 for i in imgs_u8]
    proj, dv = {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV)
    torch.manual_seed(3)
    a = model([i.to(DEV) for i in imgs_f], proj, dv)["depth"]
    torch.manual_seed(3)
    b = model([i.to(DEV) for i in imgs_u8], proj, dv)["depth"]
    assert all(torch.equal(x, y) for x, y in zip(a, b))
