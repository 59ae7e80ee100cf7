"""The NCHW operator surface of `diffmvs_b200.models` against the REAL reference's recorded operator I/O (`-m gpu`).

north_star: "keeping the models/ operator surface ... so it drops into test.py/train.py unchanged".  Every test here
builds one drop-in operator exactly as the reference's constructors would (`/root/reference/models/diffusion.py:47-136`),
loads the corresponding slice of the seeded state dict, calls it with the reference's own NCHW arguments recorded by
`oracle/make_ops_golden.py` (which ran `/root/reference`'s `CasDiffMVS` on CPU fp32) and compares with what the reference
returned.  Tolerances: depth-like outputs 1e-4 rel-L1 (bar: 1e-3), feature maps 5e-5, gathers 1e-5.
"""
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffmvs_b200 import synth
from diffmvs_b200 import models as M
from oracle import spec
from tests.helpers import GOLDEN_DIR, WEIGHT_SEED, digest, rel_l1

pytestmark = pytest.mark.gpu
DEV = "cuda"
CASES = ("cas_tiny", "cfg1")
_CACHE = {}


def _case(case):
    if case not in _CACHE:
        g = np.load(f"{GOLDEN_DIR}/ops_{case}.npz")
        args = synth.workload_args(case)
        sd = synth.synth_state_dict(spec.state_dict_shapes(args), WEIGHT_SEED)
        assert digest(sd) == str(g["digest"]), "weights differ from the ones the fixture was generated with"
        _CACHE[case] = (g, args, sd)
    return _CACHE[case]


def _t(g, key):
    return torch.from_numpy(g[key]).to(DEV)


def _load(mod, sd, prefix):
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    res = mod.load_state_dict(sub, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.rsplit(".", 1)[-1] in spec.SCHEDULE_BUFFERS for k in res.missing_keys), res.missing_keys
    return mod.to(DEV).eval()


def _views(g):
    return len([k for k in g.files if k.startswith("feature_v") and k.endswith("_stage1")])


def _features(g, stage):
    return [_t(g, f"feature_v{v}_stage{stage}") for v in range(_views(g))]


def _stages(args):
    return [s for s in (2, 3) if args.stage_iters[s - 1] != 0]


@pytest.mark.parametrize("case", CASES)
def test_feature_net_and_context_net(case):
    g, args, sd = _case(case)
    cas = args.stage_iters[2] != 0
    imgs, _, _ = synth.workload_inputs(case)
    net = _load(M.FeatureNet(8, [48, 32, 16 if cas else 0]), sd, "feature.")
    with torch.no_grad():
        for v in (0, len(imgs) - 1):
            out = net(imgs[v].to(DEV))
            for k in out:
                ref = _t(g, f"feature_v{v}_{k}")
                assert tuple(out[k].shape) == tuple(ref.shape)
                assert rel_l1(out[k], ref) < 5e-5, (v, k)
        odim = [args.hidden_dim[i] + args.context_dim[i] for i in range(3)]
        cnet = _load(M.ContextNet(odim), sd, "context.")
        cout = cnet(imgs[0].to(DEV))
        for k in cout:
            assert rel_l1(cout[k], _t(g, f"context_{k}")) < 5e-5, k


@pytest.mark.parametrize("case", CASES)
def test_initial_cost(case):
    """InitialCost.forward (module.py:487-573) with the reference's argument list."""
    g, args, sd = _case(case)
    net = _load(M.InitialCost(args.context_dim[0], args.cost_dim_stage[0]), sd, "depthnet.")
    scale = partial(M.disp_to_depth, min_depth=_t(g, "depth_min"), max_depth=_t(g, "depth_max"))
    with torch.no_grad():
        mask, inv, depth, vw, conf = net(_features(g, 1), _t(g, "depthnet_context"), _t(g, "depthnet_proj"),
                                         depth_values=_t(g, "depthnet_depth_values"), scale_inv_depth=scale)
    for name, got, tol in (("mask", mask, 1e-4), ("inv", inv, 1e-4), ("depth", depth, 1e-4), ("view_weights", vw, 1e-4),
                           ("conf", conf, 1e-4)):
        ref = _t(g, "depthnet_" + name)
        assert tuple(got.shape) == tuple(ref.shape), (name, tuple(got.shape), tuple(ref.shape))
        assert rel_l1(got, ref) < tol, (name, rel_l1(got, ref))


@pytest.mark.parametrize("case", CASES)
def test_pixel_view_weight_and_cost_regularization(case):
    g, args, sd = _case(case)
    G = args.cost_dim_stage[0]
    with torch.no_grad():
        pvw = _load(M.PixelViewWeight(G), sd, "depthnet.pixel_view_weight.")
        w = pvw(_t(g, "pvw_in"))
        assert tuple(w.shape) == tuple(g["pvw_out"].shape)
        assert rel_l1(w, _t(g, "pvw_out")) < 5e-5
        reg = _load(M.CostRegNet_small(G, 8), sd, "depthnet.cost_regularization.")
        logits = reg(_t(g, "costreg_in"))
        assert tuple(logits.shape) == tuple(g["costreg_out"].shape)
        assert rel_l1(logits, _t(g, "costreg_out")) < 5e-5


def _get_cost_module(args):
    return M.GetCost(args.cost_dim_stage[1], min_radius=args.min_radius, max_radius=args.max_radius).to(DEV).eval()


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("call", (1, 2))
def test_get_cost(case, call):
    """GetCost.forward (module.py:583-667): first call of the refinement (confidence None) and second (confidence
    given), keyword arguments as `functools.partial` passes them (diffusion.py:242-251)."""
    g, args, sd = _case(case)
    p = f"getcost{call}_"
    conf = _t(g, p + "confidence") if (p + "confidence") in g.files else None
    assert (conf is None) == (call == 1)
    with torch.no_grad():
        cost, samples = _get_cost_module(args)(
            _t(g, p + "inv"), features=_features(g, 2), proj_matrices=_t(g, p + "proj"),
            depth_interval=float(g[p + "interval"]), depth_max=_t(g, "depth_max"), depth_min=_t(g, "depth_min"),
            CostNum=int(g[p + "costnum"]), view_weights=_t(g, p + "view_weights"), confidence=conf)
    assert tuple(cost.shape) == tuple(g[p + "cost"].shape) and tuple(samples.shape) == tuple(g[p + "samples"].shape)
    assert rel_l1(samples, _t(g, p + "samples")) < 1e-6
    assert rel_l1(cost, _t(g, p + "cost")) < 2e-5


@pytest.mark.parametrize("case", CASES)
def test_condition_encoder(case):
    g, args, sd = _case(case)
    for s in _stages(args):
        ctx = args.context_dim[s - 1]
        enc = M.ConditionEncoder(num_sample=args.CostNum[s - 1], cost_dim=args.cost_dim_stage[s - 1] * args.CostNum[s - 1],
                                 hidden_dim=ctx, out_chs=ctx)
        enc = _load(enc, sd, f"update_block_depth{s}.encoder.")
        with torch.no_grad():
            out = enc(_t(g, f"enc{s}_depth"), _t(g, f"enc{s}_samples"), _t(g, f"enc{s}_cost"))
        ref = _t(g, f"enc{s}_out")
        assert tuple(out.shape) == tuple(ref.shape)
        assert rel_l1(out, ref) < 5e-5, s
        assert torch.equal(out[:, -1:].contiguous(), _t(g, f"enc{s}_depth"))     # the concatenated depth is a copy


def _unet(args, s):
    ctx = args.context_dim[s - 1]
    return M.Unet(dim=args.unet_dim[s - 1], hidden_dim=args.hidden_dim[s - 1], input_dim=2 * ctx, out_dim=1,
                  dim_mults=((1,), (1, 2), (1, 2, 4))[s - 1])


@pytest.mark.parametrize("case", CASES)
def test_unet_and_gru(case):
    g, args, sd = _case(case)
    for s in _stages(args):
        unet = _load(_unet(args, s), sd, f"update_block_depth{s}.unet.")
        with torch.no_grad():
            hidden, delta, conf = unet(_t(g, f"unet{s}_in"), _t(g, f"unet{s}_hidden_in"), _t(g, f"unet{s}_time"))
            for name, got in (("hidden", hidden), ("delta", delta), ("conf", conf)):
                ref = _t(g, f"unet{s}_{name}")
                assert tuple(got.shape) == tuple(ref.shape), (s, name)
                assert rel_l1(got, ref) < 1e-4, (s, name, rel_l1(got, ref))
            h = unet.gru(_t(g, f"gru{s}_h"), _t(g, f"gru{s}_x"))
            assert rel_l1(h, _t(g, f"gru{s}_out")) < 5e-5, s


@pytest.mark.parametrize("case", CASES)
def test_diffusion_update_block(case, monkeypatch):
    """DiffusionUpdateBlockDepth.forward, eval branch (update.py:466-521), driven by the drop-in GetCost through the
    same `functools.partial` closure the reference builds (diffusion.py:242-251)."""
    g, args, sd = _case(case)
    cas = args.stage_iters[2] != 0
    mults = ((1,), (1, 2), (1, 2, 4))
    for s in _stages(args):
        i = s - 1
        blk = M.DiffusionUpdateBlockDepth(args, dim=args.unet_dim[i], dim_mults=mults[i], hidden_dim=args.hidden_dim[i],
                                          num_sample=args.CostNum[i], cost_dim=args.cost_dim_stage[i] * args.CostNum[i],
                                          context_dim=args.context_dim[i], stage_idx=i, iters=args.stage_iters[i],
                                          ratio=2 if cas else 4)
        blk = _load(blk, sd, f"update_block_depth{s}.")
        vw = F.interpolate(_t(g, "depthnet_view_weights"), scale_factor=2 ** i, mode="nearest")
        proj = synth.workload_inputs(case)[1][f"stage{s}"].to(DEV)
        cost_fn = partial(_get_cost_module(args), features=_features(g, s), proj_matrices=proj,
                          depth_interval=(1.0 / args.numdepth) * (4, 2, 1)[i], depth_max=_t(g, "depth_max"),
                          depth_min=_t(g, "depth_min"), CostNum=args.CostNum[i], view_weights=vw)
        noise = _t(g, f"noise_{int(g[f'block{s}_noise_index'])}")
        monkeypatch.setattr(torch, "randn_like", lambda like, **kw: noise.view(like.shape))
        with torch.no_grad():
            mask, hidden, inv_list, conf_list = blk(cost_fn, _t(g, f"block{s}_inv0"), _t(g, f"block{s}_hidden0"),
                                                    _t(g, f"block{s}_context"))
        for name, got in (("mask", mask), ("hidden", hidden), ("inv_last", inv_list[-1]), ("conf_last", conf_list[-1])):
            ref = _t(g, f"block{s}_{name}")
            assert tuple(got.shape) == tuple(ref.shape), (s, name, tuple(got.shape), tuple(ref.shape))
            assert rel_l1(got, ref) < 1e-4, (s, name, rel_l1(got, ref))


@pytest.mark.parametrize("case", CASES)
def test_upsample_depth(case):
    """upsample_depth (module.py:237-248) on every call of the reference forward: ratio 2 (Cas) and 4 (DiffMVS)."""
    g, args, sd = _case(case)
    n = 1
    while f"upsample{n}_out" in g.files:
        with torch.no_grad():
            out = M.upsample_depth(_t(g, f"upsample{n}_depth"), _t(g, f"upsample{n}_mask"), ratio=int(g[f"upsample{n}_ratio"]))
        ref = _t(g, f"upsample{n}_out")
        assert tuple(out.shape) == tuple(ref.shape)
        assert rel_l1(out, ref) < 1e-6, n
        n += 1
    assert n > 2
