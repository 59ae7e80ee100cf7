"""Operator surface of the reference's `models/module.py`, backed by the sm_100a kernels.

Same class / function names, constructor arguments, sub-module attribute names (hence identical
`state_dict` keys, so the reference's checkpoints load) and forward signatures as
`/root/reference/models/module.py`.  Tensors cross this surface in the reference's logical NCHW
shapes; results are channels-last in memory.  Inference only: there is no CPU path and no autograd -
`forward` raises on CPU tensors or in training mode.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops, pipeline

Tensor = torch.Tensor


class _PlannedModule(nn.Module):
    """Caches the kernel-ready constants ("plan") of a module per device; rebuilt after the
    parameters change (`load_state_dict`, `.to()`, `.cuda()`)."""

    def __init__(self):
        super().__init__()
        self._plan = None
        self._plan_key = None
        self.register_load_state_dict_post_hook(lambda module, _keys: module.invalidate_plan())

    def invalidate_plan(self):
        for m in self.modules():
            if isinstance(m, _PlannedModule):
                m._plan, m._plan_key = None, None

    def _apply(self, fn, *a, **k):
        self.invalidate_plan()
        return super()._apply(fn, *a, **k)

    def _build_plan(self, sd: Dict[str, Tensor], device):  # pragma: no cover - abstract
        raise NotImplementedError

    def plan(self, device):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: diffmvs_b200 implements the inference path only; call .eval()")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"{type(self).__name__}: CUDA tensors required (got {device}); there is no CPU fallback")
        key = (device, tuple(p._version for p in self.parameters()))
        if self._plan is None or self._plan_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._plan = self._build_plan(sd, device)
            self._plan_key = key
        return self._plan


# ---------------------------------------------------------------------------------------------
# parameter containers with the reference's attribute layout (module.py:24-150, 279-319)
# ---------------------------------------------------------------------------------------------
class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.kernel_size, self.stride, self.relu = kernel_size, stride, relu
        self.bn = nn.BatchNorm2d(out_channels, momentum=bn_momentum) if bn else None


class Conv3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert stride in [1, 2]
        self.out_channels, self.kernel_size, self.stride, self.relu = out_channels, kernel_size, stride, relu
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum) if bn else None


class Deconv3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert stride in [1, 2]
        self.out_channels, self.stride, self.relu = out_channels, stride, relu
        self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=stride, bias=(not bn), **kwargs)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum) if bn else None


class ConvBnReLU(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)


class ConvBn(ConvBnReLU):
    pass


class ResidualBlock(nn.Module):
    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = ConvBnReLU(in_planes, planes, 3, stride=stride, pad=1)
        self.conv2 = ConvBn(planes, planes, 3, stride=1, pad=1)
        self.downsample = None if stride == 1 else ConvBn(in_planes, planes, 3, stride=stride, pad=1)


# ---------------------------------------------------------------------------------------------
# SepConvGRU (module.py:152-179)
# ---------------------------------------------------------------------------------------------
class SepConvGRU(_PlannedModule):
    """Separable convolutional GRU from RAFT."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        self.hidden_dim = hidden_dim
        c = hidden_dim + input_dim
        self.convz1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convr1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convq1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convz2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))
        self.convr2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))
        self.convq2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))

    def _build_plan(self, sd, device):
        from .. import packing
        return [tuple(pc.to(device) for pc in packing.pack_gru(sd, "", t)) for t in ("1", "2")]

    def forward(self, h, x):
        plan = self.plan(h.device)
        hh, xx = ops.to_nhwc(h), ops.to_nhwc(x)
        hid = self.hidden_dim
        for (zr_pc, q_pc), pad in zip(plan, ((0, 0, 2), (0, 2, 0))):
            zr = ops.conv(hh, zr_pc, x2=xx, pad=pad, epi=ops.EPI_GRU_ZR, aux1=hh, gru_hidden=hid)
            hh = ops.conv(zr[..., hid:], q_pc, x2=xx, pad=pad, epi=ops.EPI_GRU_Q, aux1=zr[..., :hid], aux2=hh)
        return ops.to_nchw_view(hh)


# ---------------------------------------------------------------------------------------------
# functions (module.py:181-277)
# ---------------------------------------------------------------------------------------------
def differentiable_warping(src_fea, src_proj, ref_proj, depth_values):
    """get warped source image features: `[B,C,Hs,Ws]`, two composed `[B,4,4]` projections and
    `depth_values [B,D,H,W]` -> `[B,C,D,H,W]` (module.py:181-218)."""
    B = src_fea.shape[0]
    pair = torch.zeros((B, 2, 2, 4, 4), device=src_fea.device, dtype=torch.float32)
    pair[:, :, 1] = torch.eye(4, device=src_fea.device)
    pair[:, 0, 0] = ref_proj.float()
    pair[:, 1, 0] = src_proj.float()
    hom = ops.compose_homographies(pair)[:, 0].contiguous()
    vol = ops.warp_volume(ops.to_nhwc(src_fea.float()), hom, depth_values.float().contiguous())
    return vol.permute(0, 4, 1, 2, 3)


def disp_to_depth(disp, min_depth, max_depth):
    """transform normalized inverse depth to metric depth (module.py:220-227; plain tensor algebra)."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    scaled_disp = scaled_disp.clamp(min=1e-6)
    return scaled_disp, 1 / scaled_disp


def depth_to_disp(depth, min_depth, max_depth):
    """transform metric depth to normalized inverse depth (module.py:229-235)."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    return (1 / depth - min_disp) / (max_disp - min_disp)


def upsample_depth(depth, mask, ratio=8):
    """upsample depth map using convex combination: `[N,1,H,W]`, `[N,9*r*r,H,W]` -> `[N,rH,rW]` (module.py:237-248)."""
    n = depth.float()[:, 0].contiguous()
    (raw,) = ops.upsample_depth(n, ops.to_nhwc(mask.float()), None, None, ratio, want="raw")
    return raw


def get_cur_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, confidence=None, min=0.2, max=2):
    """sample new depth hypotheses in the inverse range (module.py:250-277).  Stand-alone version of the
    sampler that `dmvs_get_cost` evaluates in its prologue; plain tensor algebra."""
    if confidence is None:
        lo = cur_depth - ndepth // 2 * depth_inteval_pixel
        hi = cur_depth + ndepth // 2 * depth_inteval_pixel
    else:
        radius = ndepth // 2 * depth_inteval_pixel
        radius = min * radius + (1 - confidence) * (max * radius - min * radius)
        lo, hi = cur_depth - radius, cur_depth + radius
    step = (hi - lo) / (ndepth - 1)
    k = torch.arange(0, ndepth, device=cur_depth.device, dtype=cur_depth.dtype).reshape(1, -1, 1, 1)
    return torch.clamp(k * step.unsqueeze(1) + lo.unsqueeze(1), min=0, max=1)


# ---------------------------------------------------------------------------------------------
# ContextNet / FeatureNet (module.py:321-420)
# ---------------------------------------------------------------------------------------------
class ContextNet(_PlannedModule):
    """context feature extraction of reference image"""

    def __init__(self, out_dim=[16, 16, 16], hidden_dim=(0, 0, 0)):
        super().__init__()
        self.in_planes = 8
        self.out_dim = list(out_dim)
        self.hidden_split = list(hidden_dim)   # not in the reference: how CasDiffMVS splits the heads
        self.conv1 = ConvBnReLU(3, 8)
        self.layer1 = self._make_layer(16, stride=2)
        self.layer2 = self._make_layer(32, stride=2)
        self.layer3 = self._make_layer(48, stride=2)
        self.output1 = nn.Conv2d(48, out_dim[0], 3, stride=1, padding=1)
        self.output2 = nn.Conv2d(32, out_dim[1], 3, stride=1, padding=1)
        if out_dim[2] > 0:
            self.output3 = nn.Conv2d(16, out_dim[2], 3, stride=1, padding=1)

    def _make_layer(self, dim, stride=1):
        layers = (ResidualBlock(self.in_planes, dim, stride=stride), ResidualBlock(dim, dim))
        self.in_planes = dim
        return nn.Sequential(*layers)

    def _build_plan(self, sd, device):
        return pipeline.ContextNetPlan(sd, device, self.out_dim, self.hidden_split)

    def forward(self, x):
        raw = self.plan(x.device).raw(ops.to_nhwc(x.float()))
        return {k: ops.to_nchw_view(v) for k, v in raw.items()}


class FeatureNet(_PlannedModule):
    """image feature extraction"""

    def __init__(self, base_channels=8, out_channel=[32, 16, 8]):
        super().__init__()
        if base_channels != 8:
            raise NotImplementedError("FeatureNet kernels are instantiated for base_channels=8 (diffusion.py:47-50)")
        self.base_channels, self.out_channel = base_channels, list(out_channel)
        b = base_channels
        self.conv0 = nn.Sequential(Conv2d(3, b, 3, 1, padding=1), Conv2d(b, b, 3, 1, padding=1))
        self.conv1 = nn.Sequential(Conv2d(b, b * 2, 5, stride=2, padding=2), Conv2d(b * 2, b * 2, 3, 1, padding=1),
                                   Conv2d(b * 2, b * 2, 3, 1, padding=1))
        self.conv2 = nn.Sequential(Conv2d(b * 2, b * 4, 5, stride=2, padding=2), Conv2d(b * 4, b * 4, 3, 1, padding=1),
                                   Conv2d(b * 4, b * 4, 3, 1, padding=1))
        self.conv3 = nn.Sequential(Conv2d(b * 4, b * 8, 5, stride=2, padding=2), Conv2d(b * 8, b * 8, 3, 1, padding=1),
                                   Conv2d(b * 8, b * 8, 3, 1, padding=1))
        self.out1 = nn.Conv2d(b * 8, out_channel[0], 1, bias=False)
        final_chs = b * 8
        self.inner1 = nn.Conv2d(b * 4, final_chs, 1, bias=True)
        self.out2 = nn.Conv2d(final_chs, out_channel[1], 3, padding=1, bias=False)
        if out_channel[2] > 0:
            self.inner2 = nn.Conv2d(b * 2, final_chs, 1, bias=True)
            self.out3 = nn.Conv2d(final_chs, out_channel[2], 3, padding=1, bias=False)

    def _build_plan(self, sd, device):
        return pipeline.FeatureNetPlan(sd, device, self.out_channel[2] > 0)

    def forward(self, x):
        out = self.plan(x.device)(ops.to_nhwc(x.float()))
        return {k: ops.to_nchw_view(v) for k, v in out.items()}


# ---------------------------------------------------------------------------------------------
# stage-1 operators (module.py:422-573)
# ---------------------------------------------------------------------------------------------
class CostRegNet_small(_PlannedModule):
    """3D cost volume regularization"""

    def __init__(self, in_channels, base_channels):
        super().__init__()
        if base_channels != 8:
            raise NotImplementedError("CostRegNet_small kernels are instantiated for base_channels=8 (module.py:477-479)")
        b = base_channels
        self.conv0 = Conv3d(in_channels, b, padding=1)
        self.conv1 = Conv3d(b, b, padding=1)
        self.conv2 = Conv3d(b, b * 2, stride=2, padding=1)
        self.conv3 = Conv3d(b * 2, b * 2, padding=1)
        self.conv4 = Conv3d(b * 2, b * 4, stride=2, padding=1)
        self.conv5 = Conv3d(b * 4, b * 4, padding=1)
        self.conv6 = Deconv3d(b * 4, b * 2, stride=2, padding=1, output_padding=1)
        self.conv7 = Deconv3d(b * 2, b * 1, stride=2, padding=1, output_padding=1)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1, bias=False)

    def _build_plan(self, sd, device):
        return pipeline.CostRegPlan(sd, device)

    def forward(self, x):
        """`[B,G,D,H,W]` -> `[B,1,D,H,W]` logits (module.py:441-448)."""
        return self.plan(x.device)(_volume_to_cl(x)).unsqueeze(1)


class PixelViewWeight(_PlannedModule):
    """Estimate pixel-wise view weight"""

    def __init__(self, G):
        super().__init__()
        self.conv = nn.Sequential(Conv3d(G, 8, padding=1), nn.Conv3d(8, 1, 3, stride=1, padding=1))

    def _build_plan(self, sd, device):
        return pipeline.ViewWeightPlan(sd, device)

    def forward(self, x):
        """`[B,G,D,H,W]` -> `[B,1,H,W]` (module.py:459-463)."""
        return self.plan(x.device)(_volume_to_cl(x)).unsqueeze(1)


def _volume_to_cl(x: Tensor) -> Tensor:
    """[B,C,D,H,W] (any memory format) -> dense [B,D,H,W,C]."""
    return x.float().permute(0, 2, 3, 4, 1).contiguous()


class InitialCost(_PlannedModule):
    """Cost volume construction in depth initialization"""

    def __init__(self, feature_dim, group_dim=8, ratio=2):
        super().__init__()
        self.group_dim = group_dim
        self.pixel_view_weight = PixelViewWeight(self.group_dim)
        self.cost_regularization = CostRegNet_small(in_channels=group_dim, base_channels=8)
        self.mask = nn.Sequential(nn.Conv2d(feature_dim, 64, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(64, ratio * ratio * 9, 1, padding=0))

    def _build_plan(self, sd, device):
        return pipeline.InitialCostPlan(sd, device, self.group_dim)

    def forward(self, features, context, proj_matrices, depth_values, scale_inv_depth=None):
        """Same contract as module.py:487-573.  `depth_values [B,D,H,W]` must hold fronto-parallel planes
        (one depth per plane, as `CasDiffMVS.forward` builds them, diffusion.py:187-192);
        `scale_inv_depth` must be the `functools.partial(disp_to_depth, min_depth=, max_depth=)` the
        reference passes (diffusion.py:146)."""
        plan = self.plan(context.device)
        kw = getattr(scale_inv_depth, "keywords", None)
        if not kw or "min_depth" not in kw or "max_depth" not in kw:
            raise ValueError("scale_inv_depth must be functools.partial(disp_to_depth, min_depth=..., max_depth=...)")
        B = context.shape[0]
        depth_min = kw["min_depth"].float().reshape(B).contiguous()
        depth_max = kw["max_depth"].float().reshape(B).contiguous()
        feats = torch.stack([ops.to_nhwc(f.float()).contiguous() for f in features], 0)
        hom = ops.compose_homographies(proj_matrices.float())
        plane_depth = depth_values.float()[:, :, 0, 0].contiguous()
        mask, n, depth, vw, conf = plan(feats, ops.to_nhwc(context.float()), hom, plane_depth, depth_min, depth_max)
        return ops.to_nchw_view(mask), n.unsqueeze(1), depth, vw, conf.unsqueeze(1)


class GetCost(nn.Module):
    """compute local cost volume"""

    def __init__(self, group_dim=4, min_radius=0.2, max_radius=2):
        super().__init__()
        self.group_dim, self.min_radius, self.max_radius = group_dim, min_radius, max_radius

    def forward(self, inverse_depth, features, proj_matrices, depth_interval, depth_max, depth_min, CostNum=4,
                view_weights=None, confidence=None):
        """Same contract as module.py:583-667: returns (cost `[B,G*D,H,W]`, samples `[B,D,H,W]`)."""
        B = inverse_depth.shape[0]
        feats = torch.stack([ops.to_nhwc(f.float()).contiguous() for f in features], 0)
        hom = ops.compose_homographies(proj_matrices.float())
        dmin = depth_min.float().reshape(B).contiguous()
        dmax = depth_max.float().reshape(B).contiguous()
        inv = inverse_depth.float().reshape(B, *inverse_depth.shape[-2:]).contiguous()
        conf = None if confidence is None else confidence.float().contiguous()
        cost, samples = ops.get_cost(feats, hom, inv, conf, view_weights.float().contiguous(), dmin, dmax,
                                     self.group_dim, CostNum, 0, float(depth_interval), float(self.min_radius),
                                     float(self.max_radius))
        return ops.to_nchw_view(cost), ops.to_nchw_view(samples)
