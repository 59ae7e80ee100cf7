"""Synthetic DTU-shaped workloads: configs, cameras, images and seeded weights.

The reference ships no checkpoints and no data, so every parity
test and the benchmark run on seeded synthetic inputs of the shapes the reference's
loader produces (`/root/reference/datasets/mvs.py:129-210`):

* ``imgs``          list of V tensors ``[B,3,H,W]`` fp32 in [0,1]
* ``proj_matrices`` dict ``stage1..stage4`` -> ``[B,V,2,4,4]`` (``[:,:,0]`` extrinsic,
                    ``[:,:,1,:3,:3]`` intrinsics scaled x0.125/0.25/0.5/1), view 0 = reference
* ``depth_values``  ``[B,numdepth]`` inverse depths ascending (``mvs.py:162-166``)

Everything here is host-side plumbing (CPU torch, uniform RNG only so results are
bit-identical across machines).
"""
from __future__ import annotations

import argparse
import math
import zlib
from typing import Dict, Iterable, List, Tuple

import torch

# ----------------------------------------------------------------------------------------
# Hyper-parameters: /root/reference/scripts/test/test_dtu_{diffmvs,casdiffmvs}.sh:13-21
# ----------------------------------------------------------------------------------------
_DIFFMVS = dict(
    stage_iters=[1, 4, 0], cost_dim_stage=[4, 4, 0], CostNum=[0, 6, 0],
    hidden_dim=[0, 32, 0], context_dim=[32, 32, 0], unet_dim=[0, 16, 8],
    scale=[0.0, 0.5, 0.0], sampling_timesteps=[0, 1, 1], ddim_eta=[0, 1, 0],
    min_radius=0.25, max_radius=4.0,
)
_CASDIFFMVS = dict(
    stage_iters=[1, 3, 3], cost_dim_stage=[4, 4, 4], CostNum=[0, 4, 4],
    hidden_dim=[0, 32, 20], context_dim=[32, 32, 16], unet_dim=[0, 16, 8],
    scale=[0.0, 0.5, 0.1], sampling_timesteps=[0, 1, 1], ddim_eta=[0, 1, 1],
    min_radius=0.125, max_radius=8.0,
)

# name -> (variant, H, W, views, numdepth_initial); BASELINE.json "configs" / SURVEY.md section 8
WORKLOADS: Dict[str, Tuple[str, int, int, int, int]] = {
    "cfg1": ("diffmvs", 128, 160, 3, 8),
    "cfg2": ("diffmvs", 512, 640, 5, 48),
    "cfg3": ("casdiffmvs", 1152, 1600, 7, 48),
    "cfg4": ("casdiffmvs", 1024, 1920, 11, 96),
    # small CasDiffMVS cases used by the parity tests
    "cas_tiny": ("casdiffmvs", 128, 160, 3, 8),
    "cas_small": ("casdiffmvs", 256, 320, 4, 16),
}


def make_args(variant: str, numdepth_initial: int = 48, numdepth: int = 384, **over) -> argparse.Namespace:
    """The argparse namespace `CasDiffMVS(args)` consumes (`/root/reference/test.py:20-77`)."""
    base = dict(_DIFFMVS if variant == "diffmvs" else _CASDIFFMVS)
    base = {k: (list(v) if isinstance(v, list) else v) for k, v in base.items()}
    base.update(numdepth_initial=numdepth_initial, numdepth=numdepth, timesteps=[1000, 1000, 1000])
    base.update(over)
    return argparse.Namespace(**base)


# per-workload overrides of the DTU defaults: Tanks & Temples runs with a smaller noise scale
# (`/root/reference/scripts/test/test_tank_casdiffmvs.sh:11-17`: numdepth_initial 96, scale 0 .125 .025, ddim_eta 0 1 1)
_WORKLOAD_OVERRIDES: Dict[str, dict] = {
    "cfg4": dict(scale=[0.0, 0.125, 0.025]),
}


def workload_args(name: str, **over) -> argparse.Namespace:
    variant, _, _, _, d_init = WORKLOADS[name]
    kw = dict(_WORKLOAD_OVERRIDES.get(name, {}))
    kw.update(over)
    return make_args(variant, numdepth_initial=d_init, **kw)


# ----------------------------------------------------------------------------------------
# Inputs
# ----------------------------------------------------------------------------------------
DEPTH_MIN, DEPTH_MAX = 425.0, 935.0  # `/root/reference/datasets/data_io.py:156-158`


def _texture(x: torch.Tensor, y: torch.Tensor, gen: torch.Generator) -> torch.Tensor:
    """Analytic RGB texture (sum of seeded sinusoids) evaluated at float coordinates."""
    out = []
    for _ in range(3):
        acc = torch.zeros_like(x)
        for k in range(6):
            fx, fy, ph = (torch.rand(3, generator=gen, dtype=torch.float64) - 0.5).tolist()
            freq = 0.02 * (2.0 ** (k * 0.9))
            acc = acc + torch.sin(2 * math.pi * (fx * freq * 4 * x + fy * freq * 4 * y + ph)) / (1 + 0.5 * k)
        out.append(acc)
    t = torch.stack(out, 0)
    return 0.5 + 0.2 * t


def make_inputs(H: int, W: int, views: int, numdepth: int = 384, seed: int = 0, batch: int = 1):
    """Seeded synthetic sample: a textured plane at depth 650 seen by V laterally shifted cameras.

    Cameras follow SURVEY.md section 8(d): shared pinhole K (f = 1.8 W), identity rotation,
    x-translation +-30*ceil(v/2) mm, DTU depth range [425, 935].
    """
    assert H % 32 == 0 and W % 32 == 0, "reference requires H,W multiples of 32 (mvs.py:104-115)"
    gen = torch.Generator().manual_seed(1000 + seed)
    f = 1.8 * W
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]], dtype=torch.float64)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    plane_depth = 650.0
    imgs: List[torch.Tensor] = []
    extr = []
    tex_gen_state = gen.get_state()
    for v in range(views):
        sign = 1.0 if v % 2 == 1 else -1.0
        tx = 0.0 if v == 0 else sign * 30.0 * math.ceil(v / 2)
        ty = 0.0 if v == 0 else 4.0 * ((v % 3) - 1)
        E = torch.eye(4, dtype=torch.float64)
        E[0, 3], E[1, 3] = tx, ty
        extr.append(E)
        # pixel (x,y) of view v sees plane point X = ((x-cx)*d/f - tx, ...); texture is indexed by
        # the reference-view pixel that sees the same point.
        gen.set_state(tex_gen_state)
        xr = xs - f * tx / plane_depth
        yr = ys - f * ty / plane_depth
        tex = _texture(xr, yr, gen)
        ngen = torch.Generator().manual_seed(2000 + 17 * seed + v)
        noise = (torch.rand((batch, 3, H, W), generator=ngen, dtype=torch.float32) - 0.5) * 0.08
        img = (tex.to(torch.float32).unsqueeze(0) + noise).clamp_(0.0, 1.0)
        imgs.append(img.contiguous())
    proj: Dict[str, torch.Tensor] = {}
    for s, sc in zip((1, 2, 3, 4), (0.125, 0.25, 0.5, 1.0)):
        P = torch.zeros(batch, views, 2, 4, 4, dtype=torch.float32)
        for v in range(views):
            Ks = K.clone()
            Ks[:2] *= sc
            P[:, v, 0] = extr[v].to(torch.float32)
            P[:, v, 1, :3, :3] = Ks.to(torch.float32)
        proj[f"stage{s}"] = P
    depth_values = torch.linspace(1.0 / DEPTH_MAX, 1.0 / DEPTH_MIN, numdepth, dtype=torch.float32)
    depth_values = depth_values.unsqueeze(0).repeat(batch, 1).contiguous()
    return imgs, proj, depth_values


def workload_inputs(name: str, seed: int = 0, batch: int = 1):
    _, H, W, V, _ = WORKLOADS[name]
    return make_inputs(H, W, V, seed=seed, batch=batch)


# ----------------------------------------------------------------------------------------
# Seeded weights. Keyed by tensor name so the fill is independent of enumeration order;
# `update_block.{0,1}.*` aliases (`/root/reference/models/diffusion.py:71,128`) get the same
# values as `update_block_depth{2,3}.*`.
# ----------------------------------------------------------------------------------------
SCHEDULE_BUFFERS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas",
    "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
)
# Output-head gains: keep the stage-1 logits O(3) (a soft, non-degenerate soft-argmax) and the
# refinement updates O(0.03) in normalised inverse depth (no wholesale clamping at [0,1]).
PROB_GAIN = 0.2
DELTA_GAIN = 0.004


def canonical_name(name: str) -> str:
    if name.startswith("update_block.0."):
        return "update_block_depth2." + name[len("update_block.0."):]
    if name.startswith("update_block.1."):
        return "update_block_depth3." + name[len("update_block.1."):]
    return name


def _uniform(shape, lo, hi, gen):
    return lo + (hi - lo) * torch.rand(tuple(shape), generator=gen, dtype=torch.float32)


def synth_tensor(name: str, shape: Iterable[int], seed: int = 123):
    """Value for one state-dict entry, or None for entries the model computes itself."""
    cname = canonical_name(name)
    shape = tuple(shape)
    leaf = cname.rsplit(".", 1)[-1]
    if leaf in SCHEDULE_BUFFERS:
        return None
    gen = torch.Generator().manual_seed((zlib.crc32(cname.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return _uniform(shape, -0.1, 0.1, gen)
    if leaf == "running_var":
        return _uniform(shape, 0.6, 1.4, gen)
    is_norm = ".bn." in cname or ".norm." in cname
    if leaf == "weight" and is_norm:
        return _uniform(shape, 0.8, 1.2, gen)
    if leaf == "bias" and is_norm:
        return _uniform(shape, -0.1, 0.1, gen)
    if leaf == "weight":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = math.sqrt(6.0 / fan_in)  # He-uniform: keeps activations O(1) through ReLU stacks
        w = _uniform(shape, -bound, bound, gen)
        if cname == "depthnet.cost_regularization.prob.weight":
            w = w * PROB_GAIN
        if cname.endswith(".unet.final_conv.weight"):
            w = w * DELTA_GAIN
        return w
    if leaf == "bias":
        b = _uniform(shape, -0.05, 0.05, gen)
        if cname.endswith(".unet.final_conv.bias"):
            b = b * 0.1
        return b
    raise KeyError(f"no synthetic recipe for state-dict entry {name!r}")


def synth_state_dict(named_shapes: Dict[str, Tuple[int, ...]], seed: int = 123) -> Dict[str, torch.Tensor]:
    """Fill every entry of `named_shapes` (name -> shape); schedule buffers are skipped."""
    out = {}
    for name, shape in named_shapes.items():
        t = synth_tensor(name, shape, seed)
        if t is not None:
            out[name] = t
    return out
