"""Operator surface of the reference's `models/update.py` (diffusion refinement), kernel-backed.

Class names, constructor arguments, attribute names (state-dict keys) and forward signatures follow
`/root/reference/models/update.py`; the modules only store parameters - the arithmetic runs in the
plans of `diffmvs_b200/pipeline.py`.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .. import ops, packing, pipeline
from .module import SepConvGRU, _PlannedModule


def cosine_beta_schedule(timesteps, s=0.008):
    """cosine schedule (update.py:26-36)."""
    return packing.cosine_schedule(timesteps, s)["betas"].double()


class _Rearrange(nn.Module):
    """Parameter-free stand-in for einops' `Rearrange` at index 0 of `Downsample` (update.py:44-48) so
    the 1x1 convolution keeps the state-dict index `.1`."""


def Upsample(dim, dim_out=None):
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(dim, dim_out or dim, 3, padding=1))


def Downsample(dim, dim_out=None):
    return nn.Sequential(_Rearrange(), nn.Conv2d(dim * 4, dim_out or dim, 1))


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class WeightStandardizedConv2d(nn.Conv2d):
    """Weight standardisation is applied when the plan is packed (`packing.standardize_weight`)."""


class Block(nn.Module):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.proj = WeightStandardizedConv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()


class ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, *, time_emb_dim=None, groups=8):
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2)) if time_emb_dim is not None else None
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class Unet(_PlannedModule):
    def __init__(self, dim, hidden_dim=32, input_dim=3, out_dim=1, dim_mults=(1, 2), resnet_block_groups=4,
                 learned_sinusoidal_cond=False, random_fourier_features=False, learned_sinusoidal_dim=16):
        super().__init__()
        if learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError("only the sinusoidal time embedding the reference ships is implemented")
        if resnet_block_groups != 4:
            raise NotImplementedError("GroupNorm kernels are written for the reference's 4 groups (update.py:169)")
        self.out_dim, self.dim, self.hidden_dim, self.dim_mults = out_dim, dim, hidden_dim, tuple(dim_mults)
        self.init_conv = nn.Conv2d(input_dim, dim, 7, padding=3)
        dims = [dim, *map(lambda m: dim * m, dim_mults)]
        in_out = list(zip(dims[:-1], dims[1:]))
        block_klass = partial(ResnetBlock, groups=resnet_block_groups)
        time_dim = dim * 4
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        n = len(in_out)
        for ind, (dim_in, dim_out_) in enumerate(in_out):
            is_last = ind >= (n - 1)
            self.downs.append(nn.ModuleList([
                block_klass(dim_in, dim_in, time_emb_dim=time_dim),
                Downsample(dim_in, dim_out_) if not is_last else nn.Conv2d(dim_in, dim_out_, 3, padding=1)]))
        mid_dim = dims[-1]
        self.gru = SepConvGRU(hidden_dim, mid_dim)
        self.mid = block_klass(hidden_dim, mid_dim)
        for ind, (dim_in, dim_out_) in enumerate(reversed(in_out)):
            is_last = ind == (n - 1)
            self.ups.append(nn.ModuleList([
                block_klass(dim_out_ + dim_in, dim_out_, time_emb_dim=time_dim),
                Upsample(dim_out_, dim_in) if not is_last else nn.Conv2d(dim_out_, dim_in, 3, padding=1)]))
        self.final_res_block = block_klass(dim * 2, dim, time_emb_dim=time_dim)
        self.final_conv = nn.Conv2d(dim, 1, 1)
        self.conf = nn.Conv2d(dim, 1, 1)

    def _build_plan(self, sd, device):
        return {"sd": sd, "device": device, "by_t": {}}

    def forward(self, x, hidden, time):
        """`x [B,Cin,H,W]`, `hidden [B,hid,h,w]`, `time [B]` (one value) -> (hidden, delta, confidence)
        as update.py:245-274.  Reading `time` synchronises; the model-level path passes python ints."""
        cache = self.plan(x.device)
        t = int(time.reshape(-1)[0])
        if t not in cache["by_t"]:
            cache["by_t"][t] = pipeline.UnetPlan(cache["sd"], cache["device"], self.dim, self.dim_mults, self.hidden_dim, t)
        plan = cache["by_t"][t]
        B = x.shape[0]
        arena = pipeline.StatsArena(x.device, B, plan.slots_per_call)
        hid, head = plan(ops.to_nhwc(x.float()), ops.to_nhwc(hidden.float()), arena)
        out = ops.to_nchw_dense(head)
        return ops.to_nchw_view(hid), out[:, 0:1], out[:, 1:2]


class ConditionEncoder(_PlannedModule):
    def __init__(self, num_sample, cost_dim, hidden_dim, out_chs):
        super().__init__()
        self.out_chs = out_chs
        self.convc1 = nn.Conv2d(cost_dim, hidden_dim, 3, padding=1)
        self.convc2 = nn.Conv2d(hidden_dim, hidden_dim, 3, padding=1)
        self.convd1 = nn.Conv2d(num_sample, hidden_dim, 3, padding=1)
        self.convd2 = nn.Conv2d(hidden_dim, hidden_dim, 3, padding=1)
        self.output = nn.Conv2d(2 * hidden_dim, out_chs - 1, 3, padding=1)

    def _build_plan(self, sd, device):
        return pipeline.EncoderPlan(sd, device)

    def forward(self, depth, depth_values, cost_volume):
        """update.py:289-297: returns `cat([ReLU(output(...)), depth], dim=1)`."""
        plan = self.plan(depth.device)
        B, _, H, W = depth.shape
        out = torch.empty((B, H, W, self.out_chs), device=depth.device, dtype=torch.float32)
        plan(ops.to_nhwc(cost_volume.float()), ops.to_nhwc(depth_values.float()), out[..., :self.out_chs - 1])
        out[..., self.out_chs - 1] = depth.float()[:, 0]
        return ops.to_nchw_view(out)


class DiffusionUpdateBlockDepth(_PlannedModule):
    def __init__(self, args, dim=16, dim_mults=(1, 2), hidden_dim=32, num_sample=4, cost_dim=16, context_dim=32,
                 stage_idx=0, iters=3, ratio=2):
        super().__init__()
        self.iters = iters
        self.encoder = ConditionEncoder(num_sample=num_sample, cost_dim=cost_dim, hidden_dim=context_dim,
                                        out_chs=context_dim)
        self.mask = nn.Sequential(nn.Conv2d(context_dim, 64, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(64, ratio * ratio * 9, 1, padding=0))
        self.unet = Unet(dim=dim, hidden_dim=hidden_dim, input_dim=self.encoder.out_chs + context_dim, out_dim=1,
                         dim_mults=dim_mults)
        self.stage_idx = stage_idx
        self.dim, self.dim_mults, self.hidden_dim, self.context_dim = dim, tuple(dim_mults), hidden_dim, context_dim
        timesteps = args.timesteps[stage_idx]
        sampling_timesteps = args.sampling_timesteps[stage_idx]
        self.timesteps = timesteps
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = args.ddim_eta[stage_idx]
        self.scale = args.scale[stage_idx]
        for name, buf in packing.cosine_schedule(timesteps).items():   # update.py:354-390
            self.register_buffer(name, buf)

    def _build_plan(self, sd, device):
        return pipeline.UpdateBlockPlan(sd, device, dim=self.dim, mults=self.dim_mults, hidden_dim=self.hidden_dim,
                                        context_dim=self.context_dim, iters=self.iters, scale=self.scale,
                                        timesteps=self.timesteps, sampling_timesteps=self.sampling_timesteps,
                                        eta=self.ddim_sampling_eta)

    def forward(self, depth_cost_func, inv_depth, hidden, context, gt_inv_depth=None, inv_init_depth=None):
        """Eval branch of update.py:466-521.  `depth_cost_func(inv [B,1,H,W], confidence=[B,H,W] or None)`
        -> (cost `[B,G*D,H,W]`, samples `[B,D,H,W]`); returns (mask, hidden, [inv_depth], [confidence])."""
        plan = self.plan(inv_depth.device)
        B, _, H, W = inv_depth.shape
        ctx = self.context_dim
        dev = inv_depth.device
        ubuf = torch.empty((B, H, W, 2 * ctx), device=dev, dtype=torch.float32)
        ubuf[..., :ctx].copy_(context.float().permute(0, 2, 3, 1))

        def cost_fn(inv, conf):
            c = None if conf is None else conf.reshape(B, H, W).contiguous()
            cost, samples = depth_cost_func(inv.view(B, 1, H, W), confidence=c)
            return ops.to_nhwc(cost.float()), ops.to_nhwc(samples.float())

        arena = pipeline.StatsArena(dev, B, plan.stats_slots())
        mask, hid, inv_last, conf_last, _ = plan(cost_fn, inv_depth.float()[:, 0].contiguous(),
                                                 ops.to_nhwc(hidden.float()), ubuf, None, None, arena)
        return (ops.to_nchw_view(mask), ops.to_nchw_view(hid), [inv_last.view(B, 1, H, W)],
                [conf_last.reshape(B, H, W).contiguous()])
