"""CPU-side checks: C-ABI surface, state-dict layout, host packing math, synthetic inputs.  No GPU needed
(nothing here launches a kernel)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

from diffmvs_b200 import _cabi, packing, synth
from diffmvs_b200.models import CasDiffMVS
from oracle import diffmvs_ref as O
from oracle import spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "diffmvs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    """The built .so loads and exports exactly what include/diffmvs_b200.h declares."""
    assert os.path.exists(_cabi.LIB_PATH), "run `python -m diffmvs_b200.build` first"
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(_cabi.SIGNATURES) == declared, "ctypes table and header disagree"
    lib = _cabi.lib()
    assert lib.dmvs_abi_version() == _cabi.ABI_VERSION
    assert b"sm_100a" in lib.dmvs_build_info()


def test_conv_desc_layout_matches_header_field_order():
    text = open(os.path.join(ROOT, "include", "diffmvs_b200.h")).read()
    body = text[text.index("typedef struct dmvs_conv_desc {"):text.index("} dmvs_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip().split("{")[-1]
        if not decl:
            continue
        for part in decl.split(","):
            names.append(part.strip().split()[-1].lstrip("*"))
    assert names == [f[0] for f in _cabi.ConvDesc._fields_]


def test_null_arguments_are_rejected_without_a_gpu():
    lib = _cabi.lib()
    assert lib.dmvs_conv_f32(None, None) == -1
    assert lib.dmvs_compose_homographies(None, None, 1, 2, None) == -1
    d = _cabi.ConvDesc()
    assert lib.dmvs_conv_f32(ctypes.byref(d), None) == -1


@pytest.mark.parametrize("workload", ["cfg1", "cas_tiny", "cfg4"])
def test_module_state_dict_matches_reference_layout(workload):
    args = synth.workload_args(workload)
    model = CasDiffMVS(args, test=True)
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    want = {k: tuple(v) for k, v in spec.state_dict_shapes(args).items()}
    assert got == want
    sd = synth.synth_state_dict(want, 123)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    # aliases share storage like the reference's ModuleList re-registration (diffusion.py:71,128)
    assert model.update_block[0] is model.update_block_depth2
    # schedule buffers equal the oracle's
    sched = O.cosine_schedule(1000)
    assert torch.equal(model.update_block_depth2.alphas_cumprod, sched["alphas_cumprod"])


def test_bn_folding_and_weight_packing():
    g = torch.Generator().manual_seed(0)
    sd = {"c.conv.weight": torch.randn(6, 5, 3, 3, generator=g), "c.bn.weight": torch.rand(6, generator=g) + 0.5,
          "c.bn.bias": torch.randn(6, generator=g), "c.bn.running_mean": torch.randn(6, generator=g),
          "c.bn.running_var": torch.rand(6, generator=g) + 0.5}
    pc = packing.pack_conv_bn(sd, "c")
    assert tuple(pc.w.shape) == (1, 3, 3, 8, 8) and pc.cin == 5 and pc.cout == 6
    x = torch.randn(2, 5, 9, 11, generator=g)
    w = pc.w[0, :, :, :5, :6].permute(3, 2, 0, 1)
    got = F.conv2d(x, w, pc.bias, padding=1)
    assert torch.allclose(got, O.conv_bn_act(sd, "c", x, relu=False), atol=1e-5)
    assert pc.w[..., 5:, :].abs().sum() == 0 and pc.w[..., 6:].abs().sum() == 0


def test_unshuffle_conv_is_a_2x2_stride2_conv():
    g = torch.Generator().manual_seed(1)
    w, b = torch.randn(7, 12, 1, 1, generator=g), torch.randn(7, generator=g)
    x = torch.randn(1, 3, 8, 10, generator=g)
    pc = packing.pack_unshuffle_conv({"d.weight": w, "d.bias": b}, "d")
    w2 = pc.w[0, :, :, :3, :7].permute(3, 2, 0, 1)
    assert torch.allclose(F.conv2d(x, w2, b, stride=2), F.conv2d(O.pixel_unshuffle2(x), w, b), atol=1e-5)


def test_time_embedding_and_affine_match_oracle():
    args = synth.workload_args("cas_tiny")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    p = "update_block_depth3.unet"
    sub = {k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")}
    temb = packing.time_embedding(sub, "time_mlp", 999, 8)
    ref = O.time_embedding(sd, p + ".time_mlp", torch.tensor([999]), 8)
    assert torch.allclose(temb, ref, atol=1e-6)
    aff = packing.block_affine(sub, "downs.0.0", temb)
    e = F.linear(F.silu(ref), sd[p + ".downs.0.0.mlp.1.weight"], sd[p + ".downs.0.0.mlp.1.bias"])[0]
    scale, shift = e.chunk(2)
    gamma, beta = sd[p + ".downs.0.0.block1.norm.weight"], sd[p + ".downs.0.0.block1.norm.bias"]
    assert torch.allclose(aff["block1"][0], gamma * (scale + 1), atol=1e-6)
    assert torch.allclose(aff["block1"][1], beta * (scale + 1) + shift, atol=1e-6)
    assert torch.equal(aff["block2"][0], sd[p + ".downs.0.0.block2.norm.weight"])


def test_synthetic_inputs_shapes_and_determinism():
    imgs, proj, dv = synth.workload_inputs("cfg1")
    assert len(imgs) == 3 and tuple(imgs[0].shape) == (1, 3, 128, 160)
    assert tuple(proj["stage1"].shape) == (1, 3, 2, 4, 4) and tuple(dv.shape) == (1, 384)
    assert torch.allclose(proj["stage1"][0, 0, 1, 0, 0] * 8, proj["stage4"][0, 0, 1, 0, 0])
    assert dv[0, 0] < dv[0, -1]
    imgs2, _, _ = synth.workload_inputs("cfg1")
    assert all(torch.equal(a, b) for a, b in zip(imgs, imgs2))
    with pytest.raises(AssertionError):
        synth.make_inputs(100, 160, 3)


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffmvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
