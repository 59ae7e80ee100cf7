"""Host-side weight preparation: reference state-dict entries -> kernel-ready constants.

Done once per (weights, device): eval-mode BatchNorm is folded into the preceding convolution
(`module.py:42-58,88-102,130-144,279-301`), weight standardisation is applied up front
(`update.py:81-94`), pixel-unshuffle + 1x1 becomes a 2x2 stride-2 convolution (`update.py:44-48`),
the GRU z/r gates are concatenated (`module.py:156-162`), the 0.25 mask scale is folded
(`module.py:511`, `update.py:473`), and the time-embedding MLP is evaluated for the (constant)
timestep (`update.py:50-62,138-153,205-211`).  All arithmetic here is plain fp32 torch on the host.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .ops import PackedConv

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def rna_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round fp32 to TF32 (10-bit mantissa), nearest with ties away from zero - bit-identical to PTX
    `cvt.rna.tf32.f32` for finite values."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & -8192).view(torch.float32)


def pack_weight(w: torch.Tensor, bias: Optional[torch.Tensor] = None, pad_cin: int = 0) -> PackedConv:
    """w [Cout,Cin,KH,KW] or [Cout,Cin,KD,KH,KW] -> PackedConv ([KD,KH,KW,cin_pad,cout_pad]).
    `pad_cin` > Cin appends zero input channels (RGB images are staged with a zero fourth channel)."""
    w = w.detach().float().cpu()
    if w.dim() == 4:
        w = w.unsqueeze(2)
    if pad_cin > w.shape[1]:
        w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, pad_cin - w.shape[1]))
    cout, cin, kd, kh, kw = w.shape
    packed = torch.zeros(kd, kh, kw, _pad4(cin), _pad4(cout), dtype=torch.float32)
    packed[:, :, :, :cin, :cout] = w.permute(2, 3, 4, 1, 0)
    b = None if bias is None else bias.detach().float().cpu().contiguous()
    p8 = lambda n: (n + 7) & ~7
    w_t = torch.zeros(kd, kh, kw, p8(cout), p8(cin), dtype=torch.float32)   # tensor-core layout, cin contiguous
    w_t[:, :, :, :cout, :cin] = w.permute(2, 3, 4, 0, 1)
    # tcgen05 layout: hi/lo TF32 planes, [2][KD][KH*KW][cin_pad8/4][cout_pad16][4] (input-channel quad innermost)
    ci8, co16 = p8(cin), (cout + 15) & ~15
    full = torch.zeros(kd, kh * kw, ci8, co16, dtype=torch.float32)
    full[:, :, :cin, :cout] = w.permute(2, 3, 4, 1, 0).reshape(kd, kh * kw, cin, cout)
    quad = full.view(kd, kh * kw, ci8 // 4, 4, co16).permute(0, 1, 2, 4, 3).contiguous()
    hi = rna_tf32(quad)
    lo = rna_tf32(quad - hi)
    w_tc = torch.stack((hi, lo), 0).contiguous()
    w_ws = pack_ws(full.view(kd, kh, kw, ci8, co16), cout) if kw <= 8 else None   # stride-1 slabs
    w_ws16 = pack_ws(full.view(kd, kh, kw, ci8, co16), cout, corr16=True) if kw <= 8 else None
    w_pair = pack_ws_pair(full.view(kd, kh, kw, ci8, co16), cout) if (w_ws is not None and cin <= 4 and kh >= 2) else None
    w_host, bias_host = None, 0.0
    if cout == 1 and cin == 8 and (kd, kh, kw) == (3, 3, 3):     # ops.conv3d_to1: weights travel as launch parameters
        w_host = w[0].permute(1, 2, 3, 0).contiguous()             # [kd][kh][kw][ci]
        bias_host = 0.0 if b is None else float(b[0])
    return PackedConv(packed.contiguous(), b, cin, cout, (kd, kh, kw), w_t.contiguous(), w_tc, w_ws, None, w_pair,
                      w_ws16, w_host, bias_host)


def ws_cc_max(kw: int) -> int:
    """Output channels per launch of the width-stacked kernel (mirrors `ws_cc_max` in csrc/conv_ws.cu)."""
    return min(64, (256 // kw) & ~7)


def ws_extent(k: int, pad: int, stride: int) -> Tuple[int, int]:
    """(smallest shift, extent) of a k-tap kernel in phase-plane shifts (mirrors `ws_extent` in csrc/conv_ws.cu):
    tap t reads input stride*o + t - pad = stride*(o + s) + phase with s = floor((t - pad - phase) / stride)."""
    smin = (-pad) // stride
    return smin, (k - 1 - pad) // stride - smin + 1


def pack_ws(full: torch.Tensor, cout: int, stride: int = 1, pad: Tuple[int, int] = (0, 0), corr16: bool = False) -> torch.Tensor:
    """full [KD,KH,KW,cin_pad8,cout_pad16] -> flat width-stacked slabs (include/diffmvs_b200.h, `w_ws`).  A stride-S
    convolution is S*S stride-1 phases over decimated input planes; with (KHe, KWe) the kernel extent in phase-plane
    shifts, every output-channel chunk (CC channels, N = KWe*CC rounded up to 16) stores the planes
    hi = rna_tf32(w), lo = rna_tf32(w - hi), each [KD][S*S][cin_pad8/8][KHe][2][N][4]; column kw'*CC + c of phase
    (pa, pb) holds tap (S*(kh'+smin_h) + pa + pad_h, S*(kw'+smin_w) + pb + pad_w) of output channel co_base + c, or
    zero where the phase has no such tap.  For stride 1 this is [KD][1][cin/8][KH][2][KW*CC][4] whatever the padding."""
    kd, kh, kw, ci8, co16 = full.shape
    S = stride
    ph, pw = pad if S > 1 else (0, 0)
    smin_h, khe = ws_extent(kh, ph, S)
    smin_w, kwe = ws_extent(kw, pw, S)
    cc_max = ws_cc_max(kwe)
    remaining, co_base, parts = (cout + 7) & ~7, 0, []
    while remaining > 0:
        cc = min(remaining, cc_max)
        n = (kwe * cc + 15) & ~15
        avail = min(cc, co16 - co_base)
        slab = torch.zeros(kd, S * S, ci8 // 8, khe, 2, n, 4, dtype=torch.float32)
        for pa in range(S):
            for pb in range(S):
                for khs in range(khe):
                    th = S * (khs + smin_h) + pa + ph
                    if not 0 <= th < kh:
                        continue
                    for kws in range(kwe):
                        tw = S * (kws + smin_w) + pb + pw
                        if not 0 <= tw < kw:
                            continue
                        # [kd][ci8][avail] -> [kd][chunk][quad][4][avail] -> [kd][chunk][quad][avail][4]
                        tap = full[:, th, tw, :, co_base:co_base + avail].reshape(kd, ci8 // 8, 2, 4, avail)
                        slab[:, pa * S + pb, :, khs, :, kws * cc:kws * cc + avail, :] = tap.permute(0, 1, 2, 4, 3)
        hi = rna_tf32(slab)
        if corr16:
            # `w_ws16`: the lo plane holds the fp16 correction operand of DMVS_PREC_WS2_TF32_F16C - per (kernel row, n) two
            # 16-byte units of 8 halves: unit 0 = fp16(hi) of the chunk's input channels 0..7, unit 1 = fp16(w - hi)
            res = slab - hi                                   # exact in fp32
            c16 = torch.zeros(kd, S * S, ci8 // 8, khe, 2, n, 8, dtype=torch.float16)
            # balanced by exact powers of two against the activation side (conv_ws2.cu: A_lo * 16, A_hi / 16), so that the
            # tiny residual w - hi stays out of fp16's subnormal range
            c16[..., 0, :, 0:4], c16[..., 0, :, 4:8] = (hi[..., 0, :, :] / 16).half(), (hi[..., 1, :, :] / 16).half()
            c16[..., 1, :, 0:4], c16[..., 1, :, 4:8] = (res[..., 0, :, :] * 16).half(), (res[..., 1, :, :] * 16).half()
            lo = c16.contiguous().view(torch.float32)         # [..., 2, n, 4]: the same bytes as a lo plane
        else:
            lo = rna_tf32(slab - hi)
        parts += [hi.reshape(-1), lo.reshape(-1)]
        co_base += cc
        remaining -= cc
    return torch.cat(parts).contiguous()


def pack_ws_pair(full: torch.Tensor, cout: int) -> torch.Tensor:
    """Stride-1 slabs for layers with <= 4 input channels (`w_ws_pair`, include/diffmvs_b200.h): the K = 8 of one MMA
    spans kernel rows (2j, 2j+1) x input channels 0..3 instead of 8 channels of one row.  Per output-channel chunk two
    planes (hi, lo), each [KD][ceil(KH/2)][2][N][4], column kw*CC + c; the odd row past KH-1 is zero."""
    kd, kh, kw, ci8, co16 = full.shape
    assert float(full[:, :, :, 4:, :].abs().max()) == 0.0, "pair packing is for <= 4 input channels"
    khp = (kh + 1) // 2
    cc_max = ws_cc_max(kw)
    remaining, co_base, parts = (cout + 7) & ~7, 0, []
    while remaining > 0:
        cc = min(remaining, cc_max)
        n = (kw * cc + 15) & ~15
        avail = min(cc, co16 - co_base)
        slab = torch.zeros(kd, khp, 2, n, 4, dtype=torch.float32)
        for r in range(kh):
            for kws in range(kw):
                # [kd][4][avail] -> [kd][avail][4]
                slab[:, r // 2, r % 2, kws * cc:kws * cc + avail, :] = full[:, r, kws, :4, co_base:co_base + avail].permute(0, 2, 1)
        hi = rna_tf32(slab)
        lo = rna_tf32(slab - hi)
        parts += [hi.reshape(-1), lo.reshape(-1)]
        co_base += cc
        remaining -= cc
    return torch.cat(parts).contiguous()


def pack_up2_phases(w3: torch.Tensor, bias: Optional[torch.Tensor] = None) -> Tuple[PackedConv, PackedConv]:
    """Phase-collapsed weights of conv3x3(pad 1)(nearest_x2(.)) for ops.conv_up2.  w3 [Cout,C,3,3] -> two PackedConv
    (output-row parity py = 0, 1), each [2*Cout, C, 2, 3]: channel px*Cout + o is output column parity px.
    Row taps: py = 0 reads low-res rows (y-1, y) with kernel rows ({0}, {1,2}) summed; py = 1 reads (y, y+1) with
    ({0,1}, {2}).  Column taps over (x-1, x, x+1): px = 0 uses ({0}, {1,2}, {}), px = 1 uses ({}, {0,1}, {2})."""
    w = w3.detach().double().cpu()
    cout, c = w.shape[:2]
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    cols = {0: ([0], [1, 2], []), 1: ([], [0, 1], [2])}
    out = []
    for py in (0, 1):
        wp = torch.zeros(2 * cout, c, 2, 3, dtype=torch.float64)
        for px in (0, 1):
            for a, khs in enumerate(rows[py]):
                for b, kws in enumerate(cols[px]):
                    for kh in khs:
                        for kw in kws:
                            wp[px * cout:(px + 1) * cout, :, a, b] += w[:, :, kh, kw]
        out.append(pack_weight(wp.float(), None if bias is None else torch.cat((bias.float(), bias.float()))))
    return out[0], out[1]


def compose_1x1_into_3x3(w3: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor]):
    """conv3x3_{w3}(conv1x1_{w1,b1}(x)), zero padding 1, as ONE 3x3 convolution of x plus a frame correction:
    returns (w [Cout,Cin,3,3], interior bias [Cout], table [3,3,Cout]).  The bias b1 reaches an output pixel through the
    taps that fall inside the image only, so on the one-pixel frame the interior bias overshoots by the out-of-image
    taps' share: table[ry][rx] (ry, rx = 0 first, 1 interior, 2 last row / column) is that correction (zero at [1][1]),
    applied by ops.border_bias_add.  Composed in float64."""
    w3d, w1d = w3.detach().double().cpu(), w1.detach().double().cpu().reshape(w1.shape[0], w1.shape[1])
    w = torch.einsum("omhw,mc->ochw", w3d, w1d)
    cout = w3.shape[0]
    s = torch.zeros(3, 3, cout, dtype=torch.float64)          # s[kh][kw][o] = sum_m w3[o,m,kh,kw] * b1[m]
    if b1 is not None:
        s = torch.einsum("omhw,m->hwo", w3d, b1.detach().double().cpu())
    table = torch.zeros(3, 3, cout, dtype=torch.float64)
    oob = {0: [0], 1: [], 2: [2]}
    for ry in range(3):
        for rx in range(3):
            for kh in range(3):
                for kw in range(3):
                    if kh in oob[ry] or kw in oob[rx]:
                        table[ry, rx] -= s[kh, kw]
    return w.float(), s.sum(dim=(0, 1)).float(), table.float().contiguous()


def pack_ws_from_packed(w: torch.Tensor, cout: int, stride: int, pad: Tuple[int, int], corr16: bool = False) -> torch.Tensor:
    """Width-stacked slabs for a given stride / padding from the FFMA layout `PackedConv.w`
    ([KD,KH,KW,cin_pad4,cout_pad4]); built on first use of a strided layer (ops.conv) and cached."""
    w = w.detach().float().cpu()
    kd, kh, kw, ci4, co4 = w.shape
    full = torch.zeros(kd, kh, kw, (ci4 + 7) & ~7, (cout + 15) & ~15, dtype=torch.float32)
    full[..., :ci4, :min(co4, full.shape[-1])] = w[..., :min(co4, full.shape[-1])]
    return pack_ws(full, cout, stride, pad, corr16=corr16)


def bn_scale_shift(sd: SD, p: str) -> Tuple[torch.Tensor, torch.Tensor]:
    scale = sd[p + ".weight"].float() / torch.sqrt(sd[p + ".running_var"].float() + BN_EPS)
    shift = sd[p + ".bias"].float() - sd[p + ".running_mean"].float() * scale
    return scale, shift


def pack_conv_bn(sd: SD, p: str, pad_cin: int = 0) -> PackedConv:
    """`module.Conv2d/Conv3d/ConvBnReLU/ConvBn`: conv (no bias) followed by eval BN."""
    w = sd[p + ".conv.weight"].float()
    scale, shift = bn_scale_shift(sd, p + ".bn")
    w = w * scale.view(-1, *([1] * (w.dim() - 1)))
    if (p + ".conv.bias") in sd:
        shift = shift + sd[p + ".conv.bias"].float() * scale
    return pack_weight(w, shift, pad_cin)


def pack_conv(sd: SD, p: str, gain: float = 1.0, rows: Optional[slice] = None) -> PackedConv:
    """Plain conv `p.weight` [+ `p.bias`], optionally a slice of output channels and a scalar gain."""
    w = sd[p + ".weight"].float()
    b = sd.get(p + ".bias")
    if rows is not None:
        w = w[rows]
        b = None if b is None else b[rows]
    if gain != 1.0:
        w = w * gain
        b = None if b is None else b.float() * gain
    return pack_weight(w, b)


def pack_deconv3d_bn(sd: SD, p: str):
    """ConvTranspose3d weight [Cin,Cout,3,3,3] + BN -> ([27,Cin,Cout] folded, bias[Cout])."""
    w = sd[p + ".conv.weight"].float()
    scale, shift = bn_scale_shift(sd, p + ".bn")
    w = w * scale.view(1, -1, 1, 1, 1)
    packed = w.permute(2, 3, 4, 0, 1).reshape(27, w.shape[0], w.shape[1]).contiguous().cpu()
    return packed, shift.contiguous().cpu()


def standardize_weight(w: torch.Tensor) -> torch.Tensor:
    """`WeightStandardizedConv2d.forward` weight transform, fp32 branch (`update.py:86-92`)."""
    w = w.float()
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + 1e-5).rsqrt()


def pack_unshuffle_conv(sd: SD, p: str) -> PackedConv:
    """Rearrange('b c (h p1) (w p2) -> b (c p1 p2) h w') + 1x1 conv == 2x2 stride-2 conv."""
    w = sd[p + ".weight"].float()          # [Cout, C*4, 1, 1]
    cout, c4 = w.shape[:2]
    return pack_weight(w.view(cout, c4 // 4, 2, 2), sd.get(p + ".bias"))


def pack_gru(sd: SD, p: str, tag: str):
    """(z|r gate conv with 2*hidden outputs, q conv) for direction `tag` in {"1","2"}."""
    pre = p + "." if p else ""
    wz, wr = sd[f"{pre}convz{tag}.weight"].float(), sd[f"{pre}convr{tag}.weight"].float()
    bz, br = sd[f"{pre}convz{tag}.bias"].float(), sd[f"{pre}convr{tag}.bias"].float()
    zr = pack_weight(torch.cat((wz, wr), 0), torch.cat((bz, br), 0))
    q = pack_conv(sd, f"{pre}convq{tag}")
    return zr, q


# --------------------------------------------------------------------------------------------
# time conditioning
# --------------------------------------------------------------------------------------------
def time_embedding(sd: SD, p: str, t: int, dim: int) -> torch.Tensor:
    """SinusoidalPosEmb(dim) -> Linear -> GELU -> Linear for a scalar timestep (`update.py:50-62,205-211`)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    ang = torch.tensor([t], dtype=torch.long)[:, None] * freq[None, :]
    emb = torch.cat((ang.sin(), ang.cos()), dim=-1)
    e = F.linear(emb, sd[f"{p}.1.weight"].float().cpu(), sd[f"{p}.1.bias"].float().cpu())
    return F.linear(F.gelu(e), sd[f"{p}.3.weight"].float().cpu(), sd[f"{p}.3.bias"].float().cpu())


def block_affine(sd: SD, p: str, temb: Optional[torch.Tensor]):
    """(g1, g0) per Block: GN(x)*(scale+1)+shift == (x-mean)*rstd*g1 + g0 (`update.py:124-131,147-153`)."""
    out = {}
    ss = None
    if temb is not None and f"{p}.mlp.1.weight" in sd:
        e = F.linear(F.silu(temb), sd[f"{p}.mlp.1.weight"].float().cpu(), sd[f"{p}.mlp.1.bias"].float().cpu())[0]
        ss = e.chunk(2, dim=0)
    for b in ("block1", "block2"):
        gamma = sd[f"{p}.{b}.norm.weight"].float().cpu()
        beta = sd[f"{p}.{b}.norm.bias"].float().cpu()
        if b == "block1" and ss is not None:
            scale, shift = ss
            out[b] = ((gamma * (scale + 1)).contiguous(), (beta * (scale + 1) + shift).contiguous())
        else:
            out[b] = (gamma.contiguous(), beta.contiguous())
    return out


def cosine_schedule(timesteps: int, s: float = 0.008) -> Dict[str, torch.Tensor]:
    """`cosine_beta_schedule` + derived buffers (`update.py:26-36,354-390`)."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999).float()
    alphas = 1.0 - betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.0)
    return {
        "betas": betas,
        "alphas_cumprod": alphas_cumprod,
        "alphas_cumprod_prev": alphas_cumprod_prev,
        "sqrt_alphas_cumprod": torch.sqrt(alphas_cumprod),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - alphas_cumprod),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - alphas_cumprod),
        "sqrt_recip_alphas": torch.sqrt(1.0 / alphas),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / alphas_cumprod),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / alphas_cumprod - 1),
        "posterior_variance": betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod),
    }
