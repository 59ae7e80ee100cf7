"""Generate `tests/golden/*.npz` by running the REAL reference (build container only).

    python -m oracle.make_golden            # from the repo root, needs /root/reference

For each case the script builds seeded synthetic inputs and weights (`diffmvs_b200/synth.py`),
loads the weights into `/root/reference`'s own `CasDiffMVS` with `strict=True` (which also pins
the state-dict layout in `oracle/spec.py`), replaces `torch.randn_like` by recorded draws, runs
the reference forward on CPU fp32 and stores: the noise draws, every output tensor, and
operator-boundary taps (one `differentiable_warping` call, `InitialCost`, the first `GetCost`
and `Unet` calls, every refinement block).  The weights are not stored: they are regenerated
from the seed and checked against the stored SHA-256.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffmvs_b200 import synth  # noqa: E402
from oracle import spec  # noqa: E402
from oracle.refimport import import_reference_models  # noqa: E402

CASES = {
    # name: (workload, overrides)
    "cfg1": ("cfg1", {}),
    "cas_tiny": ("cas_tiny", {}),
    "cas_tiny_ddim2": ("cas_tiny", dict(sampling_timesteps=[0, 2, 2], ddim_eta=[0, 1.0, 0.5])),
}
WEIGHT_SEED = 123
NOISE_SEED = 7


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def _np(t):
    return t.detach().cpu().numpy()


def run_case(ref_models, case: str):
    workload, over = CASES[case]
    args = synth.workload_args(workload, **over)
    shapes = spec.state_dict_shapes(args)
    sd = synth.synth_state_dict(shapes, WEIGHT_SEED)
    model = ref_models.CasDiffMVS(args, test=True).eval()
    ref_sd = model.state_dict()
    assert set(ref_sd) == set(shapes), "oracle/spec.py key set differs from the reference"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, tuple(v.shape), shapes[k])
    full = dict(ref_sd)
    full.update(sd)
    model.load_state_dict(full, strict=True)

    imgs, proj, depth_values = synth.workload_inputs(workload)
    out = {"digest": np.array(state_dict_digest(sd))}

    gen = torch.Generator().manual_seed(NOISE_SEED)
    draws = []

    def recorded_randn(x):
        t = torch.randn(x.shape, generator=gen, dtype=torch.float32)
        draws.append(t)
        return t

    import models.module as rmod

    taps = {}
    real_warp = rmod.differentiable_warping

    def tap_warp(src_fea, src_proj, ref_proj, depth_values):
        y = real_warp(src_fea, src_proj, ref_proj, depth_values)
        if "warp_out" not in taps:
            taps.update(warp_src=_np(src_fea), warp_src_proj=_np(src_proj), warp_ref_proj=_np(ref_proj),
                        warp_depth=_np(depth_values), warp_out=_np(y))
        return y

    def hook_once(mod, name, fmt):
        def hook(_m, inp, outp):
            if name + "_done" in taps:
                return
            taps[name + "_done"] = True
            fmt(inp, outp)
        return mod.register_forward_hook(hook)

    handles = []

    def fmt_depthnet(inp, outp):
        for n, t in zip(("mask", "inv", "depth", "view_weights", "conf"), outp):
            taps["depthnet_" + n] = _np(t)
    handles.append(hook_once(model.depthnet, "depthnet", fmt_depthnet))

    def fmt_unet(tag):
        def f(inp, outp):
            taps[tag + "_in"] = _np(inp[0])
            taps[tag + "_hidden_in"] = _np(inp[1])
            for n, t in zip(("hidden", "delta", "conf"), outp):
                taps[f"{tag}_{n}"] = _np(t)
        return f

    def fmt_block(tag):
        def f(inp, outp):
            mask, hidden, inv_list, conf_list = outp
            taps[tag + "_mask"] = _np(mask)
            taps[tag + "_hidden"] = _np(hidden)
            taps[tag + "_inv_last"] = _np(inv_list[-1])
            taps[tag + "_conf_last"] = _np(conf_list[-1])
        return f

    handles.append(hook_once(model.update_block_depth2.unet, "unet2", fmt_unet("unet2")))
    handles.append(hook_once(model.update_block_depth2, "block2", fmt_block("block2")))
    if hasattr(model, "update_block_depth3"):
        handles.append(hook_once(model.update_block_depth3.unet, "unet3", fmt_unet("unet3")))
        handles.append(hook_once(model.update_block_depth3, "block3", fmt_block("block3")))

    # GetCost is invoked through functools.partial with kwargs only -> forward hook with kwargs
    def getcost_hook(_m, a, kw, outp):
        if "getcost_cost" in taps:
            return
        taps["getcost_inv"] = _np(a[0])
        taps["getcost_cost"] = _np(outp[0])
        taps["getcost_samples"] = _np(outp[1])
    handles.append(model.GetCost.register_forward_hook(getcost_hook, with_kwargs=True))

    real_randn = torch.randn_like
    torch.randn_like = recorded_randn
    rmod.differentiable_warping = tap_warp
    try:
        with torch.no_grad():
            res = model(imgs, proj, depth_values)
    finally:
        torch.randn_like = real_randn
        rmod.differentiable_warping = real_warp
        for h in handles:
            h.remove()

    for i, t in enumerate(draws):
        out[f"noise_{i}"] = _np(t)
    for i, t in enumerate(res["depth"]):
        out[f"depth_{i}"] = _np(t)
    for i, t in enumerate(res["photometric_confidence"]):
        out[f"photo_conf_{i}"] = _np(t)
    assert len(res["conf"]) == 0
    if not over:  # operator taps only for the shipped-script cases; the DDIM case pins outputs only
        for k, v in taps.items():
            if not k.endswith("_done"):
                out["tap_" + k] = v
    return out


def main():
    ref_models = import_reference_models()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for case in CASES:
        out = run_case(ref_models, case)
        path = os.path.join(ROOT, "tests", "golden", case + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case}: {len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
