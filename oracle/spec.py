"""State-dict layout of the reference `CasDiffMVS` (oracle, test infrastructure).

Enumerates every key and shape `CasDiffMVS(args).state_dict()` holds, derived from the
constructors in `/root/reference/models/{diffusion.py:12-136, module.py, update.py:161-390}`,
without instantiating any module.  `oracle/make_golden.py` checks it against the real reference
with `load_state_dict(strict=True)`; `tests/test_models_host.py` checks the product modules
against it.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

Shapes = "OrderedDict[str, Tuple[int, ...]]"

SCHEDULE_BUFFERS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas",
    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
)


def _bn(out: dict, p: str, c: int):
    out[p + ".weight"] = (c,)
    out[p + ".bias"] = (c,)
    out[p + ".running_mean"] = (c,)
    out[p + ".running_var"] = (c,)
    out[p + ".num_batches_tracked"] = ()


def _conv_bn(out: dict, p: str, cin: int, cout: int, k: Tuple[int, ...]):
    out[p + ".conv.weight"] = (cout, cin) + tuple(k)
    _bn(out, p + ".bn", cout)


def _conv(out: dict, p: str, cin: int, cout: int, k: Tuple[int, ...], bias: bool = True):
    out[p + ".weight"] = (cout, cin) + tuple(k)
    if bias:
        out[p + ".bias"] = (cout,)


def _linear(out: dict, p: str, cin: int, cout: int):
    out[p + ".weight"] = (cout, cin)
    out[p + ".bias"] = (cout,)


def _resnet_block(out: dict, p: str, cin: int, cout: int, time_dim):
    if time_dim is not None:
        _linear(out, p + ".mlp.1", time_dim, 2 * cout)
    for b, ci in (("block1", cin), ("block2", cout)):
        _conv(out, f"{p}.{b}.proj", ci, cout, (3, 3))
        out[f"{p}.{b}.norm.weight"] = (cout,)
        out[f"{p}.{b}.norm.bias"] = (cout,)
    if cin != cout:
        _conv(out, p + ".res_conv", cin, cout, (1, 1))


def _unet(out: dict, p: str, dim: int, mults, hidden: int, cin: int):
    _conv(out, p + ".init_conv", cin, dim, (7, 7))
    time_dim = 4 * dim
    _linear(out, p + ".time_mlp.1", dim, time_dim)
    _linear(out, p + ".time_mlp.3", time_dim, time_dim)
    dims = [dim] + [dim * m for m in mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    n = len(in_out)
    for i, (di, do) in enumerate(in_out):
        _resnet_block(out, f"{p}.downs.{i}.0", di, di, time_dim)
        if i < n - 1:
            _conv(out, f"{p}.downs.{i}.1.1", 4 * di, do, (1, 1))
        else:
            _conv(out, f"{p}.downs.{i}.1", di, do, (3, 3))
    mid = dims[-1]
    for tag, k in (("1", (1, 5)), ("2", (5, 1))):
        for g in "zrq":
            _conv(out, f"{p}.gru.conv{g}{tag}", hidden + mid, hidden, k)
    # registration order in the reference is z1,r1,q1,z2,r2,q2 - order is irrelevant for a dict compare
    _resnet_block(out, p + ".mid", hidden, mid, None)
    for i, (di, do) in enumerate(reversed(in_out)):
        _resnet_block(out, f"{p}.ups.{i}.0", do + di, do, time_dim)
        if i < n - 1:
            _conv(out, f"{p}.ups.{i}.1.1", do, di, (3, 3))
        else:
            _conv(out, f"{p}.ups.{i}.1", do, di, (3, 3))
    _resnet_block(out, p + ".final_res_block", 2 * dim, dim, time_dim)
    _conv(out, p + ".final_conv", dim, 1, (1, 1))
    _conv(out, p + ".conf", dim, 1, (1, 1))


def _update_block(out: dict, p: str, dim: int, mults, hidden: int, num_sample: int, cost_dim: int,
                  ctx: int, ratio: int, timesteps: int):
    _conv(out, p + ".encoder.convc1", cost_dim, ctx, (3, 3))
    _conv(out, p + ".encoder.convc2", ctx, ctx, (3, 3))
    _conv(out, p + ".encoder.convd1", num_sample, ctx, (3, 3))
    _conv(out, p + ".encoder.convd2", ctx, ctx, (3, 3))
    _conv(out, p + ".encoder.output", 2 * ctx, ctx - 1, (3, 3))
    _conv(out, p + ".mask.0", ctx, 64, (3, 3))
    _conv(out, p + ".mask.2", 64, ratio * ratio * 9, (1, 1))
    _unet(out, p + ".unet", dim, mults, hidden, 2 * ctx)
    for b in SCHEDULE_BUFFERS:
        out[f"{p}.{b}"] = (timesteps,)


def state_dict_shapes(args) -> "OrderedDict[str, Tuple[int, ...]]":
    cas = args.stage_iters[2] != 0
    out: Dict[str, Tuple[int, ...]] = OrderedDict()
    feat_dims = [48, 32, 16 if cas else 0]
    hd, cd = args.hidden_dim, args.context_dim
    ctx_out = [hd[i] + cd[i] for i in range(3)]
    # FeatureNet(base_channels=8)
    _conv_bn(out, "feature.conv0.0", 3, 8, (3, 3))
    _conv_bn(out, "feature.conv0.1", 8, 8, (3, 3))
    c = 8
    for lvl in (1, 2, 3):
        _conv_bn(out, f"feature.conv{lvl}.0", c, 2 * c, (5, 5))
        _conv_bn(out, f"feature.conv{lvl}.1", 2 * c, 2 * c, (3, 3))
        _conv_bn(out, f"feature.conv{lvl}.2", 2 * c, 2 * c, (3, 3))
        c *= 2
    _conv(out, "feature.out1", 64, feat_dims[0], (1, 1), bias=False)
    _conv(out, "feature.inner1", 32, 64, (1, 1))
    _conv(out, "feature.out2", 64, feat_dims[1], (3, 3), bias=False)
    if cas:
        _conv(out, "feature.inner2", 16, 64, (1, 1))
        _conv(out, "feature.out3", 64, feat_dims[2], (3, 3), bias=False)
    # ContextNet
    _conv_bn(out, "context.conv1", 3, 8, (3, 3))
    cin = 8
    for li, dim in ((1, 16), (2, 32), (3, 48)):
        _conv_bn(out, f"context.layer{li}.0.conv1", cin, dim, (3, 3))
        _conv_bn(out, f"context.layer{li}.0.conv2", dim, dim, (3, 3))
        _conv_bn(out, f"context.layer{li}.0.downsample", cin, dim, (3, 3))
        _conv_bn(out, f"context.layer{li}.1.conv1", dim, dim, (3, 3))
        _conv_bn(out, f"context.layer{li}.1.conv2", dim, dim, (3, 3))
        cin = dim
    _conv(out, "context.output1", 48, ctx_out[0], (3, 3))
    _conv(out, "context.output2", 32, ctx_out[1], (3, 3))
    if ctx_out[2] > 0:
        _conv(out, "context.output3", 16, ctx_out[2], (3, 3))
    # hidden_init
    _conv_bn(out, "hidden_init.0.0", hd[1], 32, (3, 3))
    _conv(out, "hidden_init.0.1", 32, hd[1], (3, 3), bias=False)
    if cas:
        _conv_bn(out, "hidden_init.1.0", hd[2], 32, (3, 3))
        _conv_bn(out, "hidden_init.1.1", 32, 32, (3, 3))
        _conv(out, "hidden_init.1.2", 32, hd[2], (3, 3), bias=False)
    # refinement blocks (+ ModuleList aliases, `diffusion.py:71,128`)
    ratio = 2 if cas else 4
    mults = [(1,), (1, 2), (1, 2, 4)]
    stages = (1, 2) if cas else (1,)
    for s in stages:
        for prefix in (f"update_block_depth{s + 1}", f"update_block.{s - 1}"):
            _update_block(out, prefix, args.unet_dim[s], mults[s], hd[s], args.CostNum[s],
                          args.cost_dim_stage[s] * args.CostNum[s], cd[s], ratio, args.timesteps[s])
    # InitialCost(cdim_stage[0], cost_dim_stage[0])
    G = args.cost_dim_stage[0]
    _conv_bn(out, "depthnet.pixel_view_weight.conv.0", G, 8, (3, 3, 3))
    _conv(out, "depthnet.pixel_view_weight.conv.1", 8, 1, (3, 3, 3))
    chans = [(G, 8), (8, 8), (8, 16), (16, 16), (16, 32), (32, 32)]
    for i, (ci, co) in enumerate(chans):
        _conv_bn(out, f"depthnet.cost_regularization.conv{i}", ci, co, (3, 3, 3))
    # ConvTranspose3d weights are [C_in, C_out, k, k, k]
    out["depthnet.cost_regularization.conv6.conv.weight"] = (32, 16, 3, 3, 3)
    _bn(out, "depthnet.cost_regularization.conv6.bn", 16)
    out["depthnet.cost_regularization.conv7.conv.weight"] = (16, 8, 3, 3, 3)
    _bn(out, "depthnet.cost_regularization.conv7.bn", 8)
    _conv(out, "depthnet.cost_regularization.prob", 8, 1, (3, 3, 3), bias=False)
    _conv(out, "depthnet.mask.0", cd[0], 64, (3, 3))
    _conv(out, "depthnet.mask.2", 64, 36, (1, 1))
    return out
