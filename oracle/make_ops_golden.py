"""Generate `tests/golden/ops_<case>.npz`: operator-surface inputs AND outputs of the REAL reference.

    python -m oracle.make_ops_golden        # from the repo root, needs /root/reference (build container only)

`oracle/make_golden.py` pins the end-to-end outputs; this script pins every operator of the `models/`
surface that north_star names: for one forward of the reference's own `CasDiffMVS` on CPU
fp32 it records, per operator, the exact tensors the reference passed in and got back, so that a GPU test can
call the drop-in operator with the same NCHW arguments (`tests/test_gpu_operator_surface.py`):

    FeatureNet, ContextNet                  module.py:357-420, 321-355     (first call)
    InitialCost                             module.py:487-573
    PixelViewWeight, CostRegNet_small       module.py:450-463, 422-448     (first call)
    GetCost                                 module.py:583-667              (first call: no confidence; second: with)
    ConditionEncoder, Unet, SepConvGRU      update.py:276-297, 161-274, module.py:152-179 (first call, per stage)
    DiffusionUpdateBlockDepth               update.py:466-521              (whole call, per stage)
    upsample_depth                          module.py:237-248              (every call)

Weights are regenerated from the seed (checked by SHA-256 like the end-to-end fixtures).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffmvs_b200 import synth  # noqa: E402
from oracle import spec  # noqa: E402
from oracle.make_golden import NOISE_SEED, WEIGHT_SEED, state_dict_digest  # noqa: E402
from oracle.refimport import import_reference_models  # noqa: E402

CASES = ("cas_tiny", "cfg1")


def _np(t):
    return t.detach().cpu().numpy().copy()


def run_case(ref_models, case: str):
    args = synth.workload_args(case)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), WEIGHT_SEED)
    model = ref_models.CasDiffMVS(args, test=True).eval()
    full = dict(model.state_dict())
    full.update(sd)
    model.load_state_dict(full, strict=True)
    imgs, proj, depth_values = synth.workload_inputs(case)

    out = {"digest": np.array(state_dict_digest(sd))}
    gen = torch.Generator().manual_seed(NOISE_SEED)
    draws = []

    def recorded_randn(x):
        t = torch.randn(x.shape, generator=gen, dtype=torch.float32)
        draws.append(t)
        return t

    count = {}

    def nth(name):
        count[name] = count.get(name, 0) + 1
        return count[name]

    handles = []

    def hook(mod, fn, kwargs=False):
        handles.append(mod.register_forward_hook(fn, with_kwargs=True) if kwargs else mod.register_forward_hook(fn))

    # ---- FeatureNet / ContextNet ---------------------------------------------------------------------------
    def feat_hook(_m, inp, outp):
        v = nth("feature") - 1
        for k, t in outp.items():
            out[f"feature_v{v}_{k}"] = _np(t)
    hook(model.feature, feat_hook)

    def ctx_hook(_m, inp, outp):
        for k, t in outp.items():
            out[f"context_{k}"] = _np(t)
    hook(model.context, ctx_hook)

    # ---- InitialCost and its children ----------------------------------------------------------------------
    def depthnet_hook(_m, a, kw, outp):
        # called as depthnet(features_stage, context, proj_matrices_stage, depth_values=, scale_inv_depth=)
        out["depthnet_context"] = _np(a[1])
        out["depthnet_proj"] = _np(a[2])
        out["depthnet_depth_values"] = _np(kw["depth_values"])
        p = kw["scale_inv_depth"]
        out["depth_min"] = _np(p.keywords["min_depth"])
        out["depth_max"] = _np(p.keywords["max_depth"])
        for n, t in zip(("mask", "inv", "depth", "view_weights", "conf"), outp):
            out["depthnet_" + n] = _np(t)
    hook(model.depthnet, depthnet_hook, kwargs=True)

    def once(name, fn):
        def h(_m, inp, outp):
            if nth(name) == 1:
                fn(inp, outp)
        return h

    hook(model.depthnet.pixel_view_weight,
         once("pvw", lambda i, o: out.update(pvw_in=_np(i[0]), pvw_out=_np(o))))
    hook(model.depthnet.cost_regularization,
         once("reg", lambda i, o: out.update(costreg_in=_np(i[0]), costreg_out=_np(o))))

    # ---- GetCost: first two calls (confidence None / given) ------------------------------------------------------
    def getcost_hook(_m, a, kw, outp):
        i = nth("getcost")
        if i > 2:
            return
        p = f"getcost{i}_"
        out[p + "inv"] = _np(a[0])
        out[p + "proj"] = _np(kw["proj_matrices"])
        out[p + "interval"] = np.array(float(kw["depth_interval"]))
        out[p + "costnum"] = np.array(int(kw["CostNum"]))
        out[p + "view_weights"] = _np(kw["view_weights"])
        if kw.get("confidence") is not None:
            out[p + "confidence"] = _np(kw["confidence"])
        out[p + "cost"] = _np(outp[0])
        out[p + "samples"] = _np(outp[1])
    hook(model.GetCost, getcost_hook, kwargs=True)

    # ---- refinement blocks ---------------------------------------------------------------------------------
    blocks = [(2, model.update_block_depth2)]
    if hasattr(model, "update_block_depth3"):
        blocks.append((3, model.update_block_depth3))
    for s, blk in blocks:
        def enc_fn(i, o, s=s):
            out.update({f"enc{s}_depth": _np(i[0]), f"enc{s}_samples": _np(i[1]), f"enc{s}_cost": _np(i[2]),
                        f"enc{s}_out": _np(o)})
        hook(blk.encoder, once(f"enc{s}", enc_fn))

        def unet_fn(i, o, s=s):
            out.update({f"unet{s}_in": _np(i[0]), f"unet{s}_hidden_in": _np(i[1]), f"unet{s}_time": _np(i[2]),
                        f"unet{s}_hidden": _np(o[0]), f"unet{s}_delta": _np(o[1]), f"unet{s}_conf": _np(o[2])})
        hook(blk.unet, once(f"unet{s}", unet_fn))

        def gru_fn(i, o, s=s):
            out.update({f"gru{s}_h": _np(i[0]), f"gru{s}_x": _np(i[1]), f"gru{s}_out": _np(o)})
        hook(blk.unet.gru, once(f"gru{s}", gru_fn))

        def blk_hook(_m, a, kw, outp, s=s):
            # called as block(depth_cost_func, inv_cur_depth, hidden, context, gt_inv_depth=, inv_init_depth=)
            out[f"block{s}_inv0"] = _np(a[1])
            out[f"block{s}_hidden0"] = _np(a[2])
            out[f"block{s}_context"] = _np(a[3])
            mask, hidden, inv_list, conf_list = outp
            out[f"block{s}_mask"] = _np(mask)
            out[f"block{s}_hidden"] = _np(hidden)
            out[f"block{s}_inv_last"] = _np(inv_list[-1])
            out[f"block{s}_conf_last"] = _np(conf_list[-1])
            out[f"block{s}_noise_index"] = np.array(len(draws) - 1)
        hook(blk, blk_hook, kwargs=True)

    # ---- upsample_depth (module-level function, resolved through models.diffusion's globals) -----------------------
    import models.diffusion as rdiff
    real_up = rdiff.upsample_depth

    def tap_up(depth, mask, ratio=8):
        y = real_up(depth, mask, ratio=ratio)
        i = nth("upsample")
        out.update({f"upsample{i}_depth": _np(depth), f"upsample{i}_mask": _np(mask),
                    f"upsample{i}_ratio": np.array(int(ratio)), f"upsample{i}_out": _np(y)})
        return y

    real_randn = torch.randn_like
    torch.randn_like = recorded_randn
    rdiff.upsample_depth = tap_up
    try:
        with torch.no_grad():
            res = model(imgs, proj, depth_values)
    finally:
        torch.randn_like = real_randn
        rdiff.upsample_depth = real_up
        for h in handles:
            h.remove()
    for i, t in enumerate(draws):
        out[f"noise_{i}"] = _np(t)
    for i, t in enumerate(res["depth"]):
        out[f"depth_{i}"] = _np(t)
    return out


def main():
    ref_models = import_reference_models()
    for case in CASES:
        out = run_case(ref_models, case)
        path = os.path.join(ROOT, "tests", "golden", f"ops_{case}.npz")
        np.savez_compressed(path, **out)
        print(f"{case}: {len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
