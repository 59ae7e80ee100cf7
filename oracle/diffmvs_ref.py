"""Functional restatement of the DiffMVS / CasDiffMVS forward pass (oracle, test infrastructure).

All functions take a flat ``sd`` (state-dict: name -> tensor, same key names as the reference's
`CasDiffMVS.state_dict()`) plus a key prefix, and run stock PyTorch ops in whatever dtype/device
the tensors live in (fp32 CPU for parity and the CPU baseline; fp64 for "how far is fp32 from
the truth" experiments).  Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def _bn(sd: SD, p: str, x: Tensor) -> Tensor:
    """Eval-mode batch norm (`module.py:46,54-55`, `nn.BatchNorm{2,3}d` defaults eps=1e-5)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=1e-5)


def conv_bn_act(sd: SD, p: str, x: Tensor, stride: int = 1, padding: int = 1, relu: bool = True) -> Tensor:
    """`module.Conv2d` / `ConvBnReLU` / `ConvBn` wrappers: conv(no bias) -> BN(eval) -> ReLU
    (`module.py:24-58`, `:279-301`)."""
    y = F.conv2d(x, sd[p + ".conv.weight"], None, stride=stride, padding=padding)
    y = _bn(sd, p + ".bn", y)
    return F.relu(y) if relu else y


def conv3d_bn_relu(sd: SD, p: str, x: Tensor, stride: int = 1) -> Tensor:
    """`module.Conv3d` (`module.py:66-102`): 3x3x3 conv, pad 1 -> BN -> ReLU."""
    y = F.conv3d(x, sd[p + ".conv.weight"], None, stride=stride, padding=1)
    return F.relu(_bn(sd, p + ".bn", y))


def deconv3d_bn_relu(sd: SD, p: str, x: Tensor) -> Tensor:
    """`module.Deconv3d` as used by CostRegNet_small (`module.py:110-144,436-437`):
    ConvTranspose3d(k3,s2,p1,output_padding 1) -> BN -> ReLU."""
    y = F.conv_transpose3d(x, sd[p + ".conv.weight"], None, stride=2, padding=1, output_padding=1)
    return F.relu(_bn(sd, p + ".bn", y))


def conv(sd: SD, p: str, x: Tensor, stride=1, padding=0) -> Tensor:
    """Plain `nn.Conv2d` with optional bias."""
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


# ----------------------------------------------------------------------------------------------
# a1 FeatureNet  (`module.py:357-420`)
# ----------------------------------------------------------------------------------------------
def feature_net(sd: SD, p: str, img: Tensor, cas: bool) -> Dict[str, Tensor]:
    x = conv_bn_act(sd, f"{p}.conv0.0", img)
    c0 = conv_bn_act(sd, f"{p}.conv0.1", x)
    levels = [c0]
    for lvl in (1, 2, 3):
        x = conv_bn_act(sd, f"{p}.conv{lvl}.0", levels[-1], stride=2, padding=2)  # 5x5 s2
        x = conv_bn_act(sd, f"{p}.conv{lvl}.1", x)
        x = conv_bn_act(sd, f"{p}.conv{lvl}.2", x)
        levels.append(x)
    _, c1, c2, c3 = levels
    out = {"stage1": conv(sd, f"{p}.out1", c3)}
    intra = F.interpolate(c3, scale_factor=2, mode="nearest") + conv(sd, f"{p}.inner1", c2)
    out["stage2"] = conv(sd, f"{p}.out2", intra, padding=1)
    if cas:
        intra = F.interpolate(intra, scale_factor=2, mode="nearest") + conv(sd, f"{p}.inner2", c1)
        out["stage3"] = conv(sd, f"{p}.out3", intra, padding=1)
    return out


# ----------------------------------------------------------------------------------------------
# a2 ContextNet (`module.py:303-355`)
# ----------------------------------------------------------------------------------------------
def _residual_block(sd: SD, p: str, x: Tensor, stride: int) -> Tensor:
    y = conv_bn_act(sd, f"{p}.conv1", x, stride=stride)
    y = conv_bn_act(sd, f"{p}.conv2", y, relu=False)
    if stride != 1:
        x = conv_bn_act(sd, f"{p}.downsample", x, stride=stride, relu=False)
    return F.relu(x + y)


def context_net(sd: SD, p: str, img: Tensor, cas: bool) -> Dict[str, Tensor]:
    x = conv_bn_act(sd, f"{p}.conv1", img)
    ctx = {}
    x = _residual_block(sd, f"{p}.layer1.0", x, 2)
    x = _residual_block(sd, f"{p}.layer1.1", x, 1)
    if cas:
        ctx["stage3"] = conv(sd, f"{p}.output3", x, padding=1)
    x = _residual_block(sd, f"{p}.layer2.0", x, 2)
    x = _residual_block(sd, f"{p}.layer2.1", x, 1)
    ctx["stage2"] = conv(sd, f"{p}.output2", x, padding=1)
    x = _residual_block(sd, f"{p}.layer3.0", x, 2)
    x = _residual_block(sd, f"{p}.layer3.1", x, 1)
    ctx["stage1"] = conv(sd, f"{p}.output1", x, padding=1)
    return ctx


# ----------------------------------------------------------------------------------------------
# a3 differentiable_warping (`module.py:181-218`)
# ----------------------------------------------------------------------------------------------
def compose_projection(proj_pair: Tensor) -> Tensor:
    """`[B,2,4,4]` (extrinsic, intrinsic) -> `[B,4,4]` with rows 0-2 = K @ E[:3,:4]
    (`module.py:520-525`, `:635-640`)."""
    P = proj_pair[:, 0].clone()
    P[:, :3, :4] = proj_pair[:, 1, :3, :3] @ proj_pair[:, 0, :3, :4]
    return P


def warp_coordinates(src_proj: Tensor, ref_proj: Tensor, depth: Tensor) -> Tuple[Tensor, Tensor]:
    """Source-image pixel coordinates (u, v), each `[B,D,H,W]`, of every reference pixel lifted to
    every hypothesis depth (`module.py:188-207`).  `p.z == 0` gets +1e-8; negative z is not masked."""
    B, D, H, W = depth.shape
    Hm = src_proj @ torch.linalg.inv(ref_proj)
    R, t = Hm[:, :3, :3], Hm[:, :3, 3]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=depth.dtype, device=depth.device),
                            torch.arange(W, dtype=depth.dtype, device=depth.device), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=depth.dtype, device=depth.device)))
    ray = (R @ pix.unsqueeze(0)).view(B, 3, 1, H, W)              # [B,3,1,H,W]
    p = ray * depth.unsqueeze(1) + t.view(B, 3, 1, 1, 1)          # [B,3,D,H,W]
    z = p[:, 2]
    z = torch.where(z == 0, z + 1e-8, z)
    return p[:, 0] / z, p[:, 1] / z


def differentiable_warping(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth: Tensor) -> Tensor:
    """Bilinear, zero-padded, align_corners=True resampling of `src_fea [B,C,Hs,Ws]` at the warp
    coordinates -> `[B,C,D,H,W]` (`module.py:208-218`)."""
    B, C, Hs, Ws = src_fea.shape
    _, D, H, W = depth.shape
    u, v = warp_coordinates(src_proj, ref_proj, depth)
    gx = u / ((Ws - 1) / 2) - 1
    gy = v / ((Hs - 1) / 2) - 1
    grid = torch.stack((gx, gy), dim=-1).view(B, D * H, W, 2)
    out = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(B, C, D, H, W)


def differentiable_warping_explicit(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth: Tensor) -> Tensor:
    """Same operator written as an explicit 4-tap gather in source pixel units (SURVEY.md
    appendix A.2) - the form the CUDA kernel implements.  Used to cross-check `grid_sample`
    semantics on small inputs."""
    B, C, Hs, Ws = src_fea.shape
    _, D, H, W = depth.shape
    u, v = warp_coordinates(src_proj, ref_proj, depth)
    x0f, y0f = torch.floor(u), torch.floor(v)
    wx1, wy1 = u - x0f, v - y0f
    flat = src_fea.reshape(B, C, Hs * Ws)
    out = torch.zeros(B, C, D, H, W, dtype=src_fea.dtype, device=src_fea.device)
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0f + dx, y0f + dy
            w = (wx1 if dx else 1 - wx1) * (wy1 if dy else 1 - wy1)
            ok = (xi >= 0) & (xi <= Ws - 1) & (yi >= 0) & (yi <= Hs - 1)
            idx = (yi.clamp(0, Hs - 1) * Ws + xi.clamp(0, Ws - 1)).long().view(B, 1, -1).expand(B, C, -1)
            tap = torch.gather(flat, 2, idx).view(B, C, D, H, W)
            out = out + tap * (w * ok.to(w.dtype)).unsqueeze(1)
    return out


# ----------------------------------------------------------------------------------------------
# inverse-depth mapping (`module.py:220-235`, `diffusion.py:140-146`)
# ----------------------------------------------------------------------------------------------
class DepthRange:
    def __init__(self, depth_values: Tensor):
        self.disp_min = depth_values[:, 0].view(-1, 1, 1, 1)     # 1/depth_max
        self.disp_max = depth_values[:, -1].view(-1, 1, 1, 1)    # 1/depth_min
        self.depth_max = 1.0 / self.disp_min
        self.depth_min = 1.0 / self.disp_max

    def to_depth(self, n: Tensor) -> Tensor:
        """normalised inverse depth -> metric depth (`disp_to_depth`, `module.py:220-227`)."""
        min_disp = 1 / self.depth_max
        max_disp = 1 / self.depth_min
        return 1 / (min_disp + (max_disp - min_disp) * n).clamp(min=1e-6)

    def to_norm(self, depth: Tensor) -> Tensor:
        """metric depth -> normalised inverse depth (`depth_to_disp`, `module.py:229-235`)."""
        min_disp = 1 / self.depth_max
        max_disp = 1 / self.depth_min
        return (1 / depth - min_disp) / (max_disp - min_disp)


def upsample_depth(n: Tensor, mask: Tensor, ratio: int) -> Tensor:
    """Convex 9-tap upsampling (`module.py:237-248`): `n [B,1,H,W]`, `mask [B,9*r*r,H,W]` -> `[B,rH,rW]`."""
    B, _, H, W = n.shape
    m = torch.softmax(mask.view(B, 9, ratio, ratio, H, W), dim=1)
    taps = F.unfold(n, 3, padding=1).view(B, 9, 1, 1, H, W)
    up = (m * taps).sum(1)                      # [B,r,r,H,W]
    return up.permute(0, 3, 1, 4, 2).reshape(B, ratio * H, ratio * W)


# ----------------------------------------------------------------------------------------------
# a4-a6 stage-1 depth initialisation (`module.py:422-573`)
# ----------------------------------------------------------------------------------------------
def group_correlation(warped: Tensor, ref: Tensor, G: int) -> Tensor:
    """`(warped * ref).mean over C/G` -> `[B,G,D,H,W]` (`module.py:529-531`, `:644-646`)."""
    B, C, D, H, W = warped.shape
    return (warped.view(B, G, C // G, D, H, W) * ref.view(B, G, C // G, 1, H, W)).mean(2)


def pixel_view_weight(sd: SD, p: str, cor: Tensor) -> Tensor:
    """`PixelViewWeight` (`module.py:450-463`) -> `[B,1,H,W]`."""
    y = conv3d_bn_relu(sd, f"{p}.conv.0", cor)
    y = F.conv3d(y, sd[f"{p}.conv.1.weight"], sd[f"{p}.conv.1.bias"], padding=1)
    return torch.sigmoid(y.squeeze(1)).max(dim=1, keepdim=True)[0]


def cost_reg_net(sd: SD, p: str, x: Tensor) -> Tensor:
    """`CostRegNet_small` (`module.py:422-448`) -> `[B,1,D,H,W]` logits."""
    c1 = conv3d_bn_relu(sd, f"{p}.conv1", conv3d_bn_relu(sd, f"{p}.conv0", x))
    c3 = conv3d_bn_relu(sd, f"{p}.conv3", conv3d_bn_relu(sd, f"{p}.conv2", c1, stride=2))
    y = conv3d_bn_relu(sd, f"{p}.conv5", conv3d_bn_relu(sd, f"{p}.conv4", c3, stride=2))
    y = c3 + deconv3d_bn_relu(sd, f"{p}.conv6", y)
    y = c1 + deconv3d_bn_relu(sd, f"{p}.conv7", y)
    return F.conv3d(y, sd[f"{p}.prob.weight"], None, padding=1)


def depth_regression(logits: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """softmax over D, expected index, window-4 confidence (`module.py:554-571`).
    Returns (expected index [B,1,H,W], floor index (long) [B,1,H,W], confidence [B,1,H,W])."""
    D = logits.shape[1]
    prob = F.softmax(logits, dim=1)
    planes = torch.arange(D, device=logits.device).view(1, D, 1, 1).to(logits.dtype)
    idx = (planes * prob).sum(1, keepdim=True)
    padded = F.pad(prob, (0, 0, 0, 0, 1, 2))                       # D+3 planes, zeros outside
    sum4 = padded[:, 0:D] + padded[:, 1:D + 1] + padded[:, 2:D + 2] + padded[:, 3:D + 3]
    j = idx.long().clamp(0, D - 1)
    return idx, j, torch.gather(sum4, 1, j)


def mask_head(sd: SD, p: str, context: Tensor) -> Tensor:
    """`0.25 * Conv1x1(ReLU(Conv3x3(context)))` (`module.py:481-485,511`, `update.py:335-339,473`)."""
    return 0.25 * conv(sd, f"{p}.2", F.relu(conv(sd, f"{p}.0", context, padding=1)))


def initial_cost(sd: SD, p: str, features: Sequence[Tensor], context: Tensor, proj: Tensor,
                 depth_planes: Tensor, rng: DepthRange, G: int, taps: Optional[dict] = None):
    """`InitialCost.forward` (`module.py:487-573`).  `proj [B,V,2,4,4]`, `depth_planes [B,D,H,W]`.
    Returns (mask, normalised inverse depth [B,1,H,W], depth [B,H,W], view_weights [B,V-1,H,W], conf)."""
    D = depth_planes.shape[1]
    ref_fea, ref_proj = features[0], compose_projection(proj[:, 0])
    mask = mask_head(sd, f"{p}.mask", context)
    wsum, acc, weights = 1e-8, 0, []
    for v in range(1, len(features)):
        warped = differentiable_warping(features[v], compose_projection(proj[:, v]), ref_proj, depth_planes)
        cor = group_correlation(warped, ref_fea, G)
        w = pixel_view_weight(sd, f"{p}.pixel_view_weight", cor)
        weights.append(w)
        wsum = wsum + w.unsqueeze(1)
        acc = acc + w.unsqueeze(1) * cor
        if taps is not None and v == 1:
            taps["stage1_cor_view1"] = cor
    volume = acc / wsum
    logits = cost_reg_net(sd, f"{p}.cost_regularization", volume).squeeze(1)
    idx, j, conf = depth_regression(logits)
    n = idx / (D - 1.0)
    if taps is not None:
        taps.update(stage1_volume=volume, stage1_logits=logits, stage1_index=idx, stage1_floor=j)
    return mask, n, rng.to_depth(n).squeeze(1), torch.cat(weights, dim=1), conf


# ----------------------------------------------------------------------------------------------
# a7 GetCost (`module.py:250-277`, `:575-667`)
# ----------------------------------------------------------------------------------------------
def depth_hypotheses(cur: Tensor, D: int, interval: float, conf: Optional[Tensor],
                     min_radius: float, max_radius: float) -> Tensor:
    """`get_cur_depth_range_samples` (`module.py:250-277`): `cur [B,H,W]` -> `[B,D,H,W]` in [0,1]."""
    if conf is None:
        lo = cur - D // 2 * interval
        hi = cur + D // 2 * interval
    else:
        r = D // 2 * interval
        r_min, r_max = min_radius * r, max_radius * r
        r = r_min + (1 - conf) * (r_max - r_min)
        lo, hi = cur - r, cur + r
    step = (hi - lo) / (D - 1)
    k = torch.arange(D, device=cur.device, dtype=cur.dtype).view(1, D, 1, 1)
    return (k * step.unsqueeze(1) + lo.unsqueeze(1)).clamp(0, 1)


def get_cost(inv_depth: Tensor, features: Sequence[Tensor], proj: Tensor, interval: float, rng: DepthRange,
             D: int, view_weights: Tensor, conf: Optional[Tensor], G: int, min_radius: float, max_radius: float):
    """`GetCost.forward` -> (cost `[B,G*D,H,W]` group-major, samples `[B,D,H,W]`)."""
    if D > 1:
        samples = depth_hypotheses(inv_depth.squeeze(1), D, interval, conf, min_radius, max_radius)
    else:
        samples = inv_depth
    depth = rng.to_depth(samples)
    ref_fea, ref_proj = features[0], compose_projection(proj[:, 0])
    wsum, acc = 1e-8, 0
    for v in range(1, len(features)):
        warped = differentiable_warping(features[v], compose_projection(proj[:, v]), ref_proj, depth)
        cor = group_correlation(warped, ref_fea, G)
        w = view_weights[:, v - 1].unsqueeze(1).unsqueeze(1)
        wsum = wsum + w
        acc = acc + w * cor
    vol = acc / wsum
    B, g, d, H, W = vol.shape
    return vol.reshape(B, g * d, H, W), samples


# ----------------------------------------------------------------------------------------------
# a8-a10 ConditionEncoder, Unet, SepConvGRU (`update.py:50-62,81-297`, `module.py:152-179`)
# ----------------------------------------------------------------------------------------------
def condition_encoder(sd: SD, p: str, inv_depth: Tensor, samples: Tensor, cost: Tensor) -> Tensor:
    c = F.relu(conv(sd, f"{p}.convc1", cost, padding=1))
    c = F.relu(conv(sd, f"{p}.convc2", c, padding=1))
    d = F.relu(conv(sd, f"{p}.convd1", samples, padding=1))
    d = F.relu(conv(sd, f"{p}.convd2", d, padding=1))
    o = F.relu(conv(sd, f"{p}.output", torch.cat((c, d), 1), padding=1))
    return torch.cat((o, inv_depth), 1)


def sep_conv_gru(sd: SD, p: str, h: Tensor, x: Tensor) -> Tensor:
    for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat((h, x), 1)
        z = torch.sigmoid(conv(sd, f"{p}.convz{tag}", hx, padding=pad))
        r = torch.sigmoid(conv(sd, f"{p}.convr{tag}", hx, padding=pad))
        q = torch.tanh(conv(sd, f"{p}.convq{tag}", torch.cat((r * h, x), 1), padding=pad))
        h = (1 - z) * h + z * q
    return h


def time_embedding(sd: SD, p: str, t: Tensor, dim: int) -> Tensor:
    """`SinusoidalPosEmb(dim)` -> Linear -> GELU -> Linear (`update.py:50-62,205-211`)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    ang = t[:, None] * freq[None, :]
    emb = torch.cat((ang.sin(), ang.cos()), dim=-1).to(sd[f"{p}.1.weight"].dtype)
    e = F.linear(emb, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"])
    return F.linear(F.gelu(e), sd[f"{p}.3.weight"], sd[f"{p}.3.bias"])


def ws_conv3x3(sd: SD, p: str, x: Tensor) -> Tensor:
    """`WeightStandardizedConv2d` (`update.py:81-94`), fp32 eps."""
    w = sd[p + ".weight"]
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return F.conv2d(x, (w - mean) * (var + 1e-5).rsqrt(), sd[p + ".bias"], padding=1)


def _block(sd: SD, p: str, x: Tensor, scale_shift=None, groups: int = 4) -> Tensor:
    y = ws_conv3x3(sd, f"{p}.proj", x)
    y = F.group_norm(y, groups, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        y = y * (scale + 1) + shift
    return F.silu(y)


def resnet_block(sd: SD, p: str, x: Tensor, temb: Optional[Tensor]) -> Tensor:
    ss = None
    if temb is not None and f"{p}.mlp.1.weight" in sd:
        e = F.linear(F.silu(temb), sd[f"{p}.mlp.1.weight"], sd[f"{p}.mlp.1.bias"])
        ss = e[:, :, None, None].chunk(2, dim=1)
    h = _block(sd, f"{p}.block1", x, ss)
    h = _block(sd, f"{p}.block2", h)
    res = conv(sd, f"{p}.res_conv", x) if f"{p}.res_conv.weight" in sd else x
    return h + res


def pixel_unshuffle2(x: Tensor) -> Tensor:
    """einops `b c (h p1) (w p2) -> b (c p1 p2) h w` with p1=p2=2 (`update.py:46`)."""
    B, C, H, W = x.shape
    return x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, C * 4, H // 2, W // 2)


def unet(sd: SD, p: str, x: Tensor, hidden: Tensor, t: Tensor, dim: int, levels: int):
    """`Unet.forward` (`update.py:245-274`) -> (hidden, delta, confidence)."""
    x = conv(sd, f"{p}.init_conv", x, padding=3)
    r = x
    temb = time_embedding(sd, f"{p}.time_mlp", t, dim)
    skips = []
    for i in range(levels):
        x = resnet_block(sd, f"{p}.downs.{i}.0", x, temb)
        skips.append(x)
        if i < levels - 1:
            x = conv(sd, f"{p}.downs.{i}.1.1", pixel_unshuffle2(x))
        else:
            x = conv(sd, f"{p}.downs.{i}.1", x, padding=1)
    hidden = sep_conv_gru(sd, f"{p}.gru", hidden, x)
    x = resnet_block(sd, f"{p}.mid", hidden, None)
    for i in range(levels):
        x = resnet_block(sd, f"{p}.ups.{i}.0", torch.cat((x, skips.pop()), 1), temb)
        if i < levels - 1:
            x = conv(sd, f"{p}.ups.{i}.1.1", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
        else:
            x = conv(sd, f"{p}.ups.{i}.1", x, padding=1)
    x = resnet_block(sd, f"{p}.final_res_block", torch.cat((x, r), 1), temb)
    return hidden, conv(sd, f"{p}.final_conv", x), torch.sigmoid(conv(sd, f"{p}.conf", x))


# ----------------------------------------------------------------------------------------------
# a11 diffusion refinement, eval branch (`update.py:26-36,354-405,466-521`)
# ----------------------------------------------------------------------------------------------
def cosine_schedule(timesteps: int, s: float = 0.008) -> Dict[str, Tensor]:
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - ac[1:] / ac[:-1], 0, 0.999).float()
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    return {
        "alphas_cumprod": alphas_cumprod,
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / alphas_cumprod),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / alphas_cumprod - 1),
    }


def refine_stage(sd: SD, p: str, cost_fn: Callable, inv0: Tensor, hidden: Tensor, context: Tensor, *,
                 iters: int, scale: float, dim: int, levels: int, timesteps: int, sampling_timesteps: int,
                 eta: float, randn: Callable[[Tensor], Tensor], taps: Optional[dict] = None, tag: str = ""):
    """Eval branch of `DiffusionUpdateBlockDepth.forward` (`update.py:466-521`).
    `randn(like)` supplies the Gaussian draws (`torch.randn_like` in the reference)."""
    B = inv0.shape[0]
    sched = cosine_schedule(timesteps)
    times = torch.linspace(-1, timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    img = (scale * randn(inv0)).to(inv0.dtype)
    mask = mask_head(sd, f"{p}.mask", context)
    inv_list, conf_list, cur_hidden = [], [], hidden
    for time, time_next in zip(times[:-1], times[1:]):
        t = torch.full((B,), time, device=inv0.device, dtype=torch.long)
        inv_list, conf_list = [], []
        inv = (inv0 + img).clamp(0, 1)
        delta = inv - inv0
        img = delta
        cur_hidden, conf = hidden, None
        for it in range(iters):
            cost, samples = cost_fn(inv, conf)
            enc = condition_encoder(sd, f"{p}.encoder", inv, samples, cost)
            cur_hidden, upd, conf = unet(sd, f"{p}.unet", torch.cat((context, enc), 1), cur_hidden, t, dim, levels)
            if taps is not None:
                taps[f"{tag}_it{it}_cost"] = cost
                taps[f"{tag}_it{it}_samples"] = samples
                taps[f"{tag}_it{it}_update"] = upd
                taps[f"{tag}_it{it}_hidden"] = cur_hidden
            conf = conf.squeeze(1)
            delta = delta + upd
            conf_list.append(conf)
            inv = (inv0 + delta).clamp(0, 1)
            delta = inv - inv0
            inv_list.append(inv)
        if time_next < 0:
            continue
        a, a_next = sched["alphas_cumprod"][time].to(inv0.dtype), sched["alphas_cumprod"][time_next].to(inv0.dtype)
        pred_noise = (sched["sqrt_recip_alphas_cumprod"][time].to(inv0.dtype) * img - delta) / \
            sched["sqrt_recipm1_alphas_cumprod"][time].to(inv0.dtype)
        sigma = eta * ((1 - a / a_next) * (1 - a_next) / (1 - a)).sqrt()
        c = (1 - a_next - sigma ** 2).sqrt()
        img = delta * a_next.sqrt() + c * pred_noise + sigma * (scale * randn(inv0)).to(inv0.dtype)
    return mask, cur_hidden, inv_list, conf_list


# ----------------------------------------------------------------------------------------------
# a13 CasDiffMVS.forward, test mode (`diffusion.py:139-295`)
# ----------------------------------------------------------------------------------------------
UNET_LEVELS = (1, 2, 3)      # len(unet_dim_mults[stage]) (`diffusion.py:33`)
INTERVAL_RATIO = (4, 2, 1)   # depth_interals_ratio (`diffusion.py:15`)


def _hidden_init(sd: SD, p: str, x: Tensor, n_strided: int) -> Tensor:
    """`hidden_init[s-1]` (`diffusion.py:53-58,91-101`): n strided Conv+BN+ReLU then a bias-free 3x3."""
    for i in range(n_strided):
        x = conv_bn_act(sd, f"{p}.{i}", x, stride=2)
    return F.conv2d(x, sd[f"{p}.{n_strided}.weight"], None, padding=1)


def casdiffmvs_forward(sd: SD, args, imgs: Sequence[Tensor], proj_matrices: Dict[str, Tensor],
                       depth_values: Tensor, randn: Callable[[Tensor], Tensor] = torch.randn_like,
                       taps: Optional[dict] = None) -> Dict[str, List[Tensor]]:
    """Test-mode forward.  Returns {"depth": [...], "photometric_confidence": [...], "conf": []}."""
    cas = args.stage_iters[2] != 0
    up_ratio = 2 if cas else 4
    depth_values = depth_values.to(imgs[0].dtype)
    rng = DepthRange(depth_values)
    interval0 = 1.0 / depth_values.size(1)

    feats = [feature_net(sd, "feature", im, cas) for im in imgs]
    ctxs = context_net(sd, "context", imgs[0], cas)
    if taps is not None:
        taps["feat_ref"] = feats[0]
        taps["ctx"] = ctxs

    depths: List[Tensor] = []
    confs: List[Tensor] = []
    view_weights = None
    for s in range(3):
        if args.stage_iters[s] == 0:
            continue
        key = f"stage{s + 1}"
        fs = [f[key] for f in feats]
        proj = proj_matrices[key].to(imgs[0].dtype)
        B, _, H, W = fs[0].shape
        if s == 0:
            D0 = args.numdepth_initial
            planes = torch.arange(D0, device=fs[0].device, dtype=imgs[0].dtype).view(1, D0, 1, 1) / (D0 - 1.0)
            planes = rng.to_depth(planes.repeat(1, 1, H, W))
            context = F.relu(ctxs[key])
            mask, inv, init_depth, view_weights, conf = initial_cost(
                sd, "depthnet", fs, context, proj, planes, rng, args.cost_dim_stage[0], taps)
            depths.append(init_depth)
            confs.append(F.interpolate(conf, scale_factor=8, mode="nearest").squeeze(1))
            depths.append(rng.to_depth(upsample_depth(inv, mask, 2).unsqueeze(1)).squeeze(1))
            if taps is not None:
                taps.update(stage1_mask=mask, stage1_inv=inv, view_weights=view_weights, stage1_conf=conf)
        else:
            inv_cur = rng.to_norm(depths[-1].unsqueeze(1))
            vw = F.interpolate(view_weights, scale_factor=2 ** s, mode="nearest")
            hdim, cdim = args.hidden_dim[s], args.context_dim[s]
            hid_part, ctx_part = torch.split(ctxs[key], [hdim, cdim], dim=1)
            hidden = torch.tanh(_hidden_init(sd, f"hidden_init.{s - 1}", hid_part, s))
            context = F.relu(ctx_part)

            def cost_fn(inv, conf, fs=fs, proj=proj, s=s, vw=vw):
                return get_cost(inv, fs, proj, interval0 * INTERVAL_RATIO[s], rng, args.CostNum[s], vw, conf,
                                args.cost_dim_stage[1], args.min_radius, args.max_radius)

            blk = f"update_block_depth{s + 1}"
            mask, hidden, inv_seq, conf_seq = refine_stage(
                sd, blk, cost_fn, inv_cur, hidden, context, iters=args.stage_iters[s], scale=args.scale[s],
                dim=args.unet_dim[s], levels=UNET_LEVELS[s], timesteps=args.timesteps[s],
                sampling_timesteps=args.sampling_timesteps[s], eta=args.ddim_eta[s], randn=randn,
                taps=taps, tag=key)
            depths.append(rng.to_depth(inv_seq[-1]).squeeze(1))
            confs.append(F.interpolate(conf_seq[-1].unsqueeze(1), scale_factor=2 ** (3 - s), mode="nearest").squeeze(1))
            depths.append(rng.to_depth(upsample_depth(inv_seq[-1], mask, up_ratio).unsqueeze(1)).squeeze(1))
            if taps is not None:
                taps[f"{key}_mask"] = mask
                taps[f"{key}_hidden0"] = hidden
    return {"depth": depths, "conf": [], "photometric_confidence": confs}
