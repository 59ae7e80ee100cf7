"""The oracle's per-operator functions against the REAL reference's recorded operator I/O (CPU, `-m "not gpu"`).

`tests/golden/ops_*.npz` (written by `oracle/make_ops_golden.py` from `/root/reference`) hold, for one forward of the
reference's own `CasDiffMVS`, the inputs and outputs of every operator of the `models/` surface.  Here each oracle
function is fed those inputs; it must reproduce the reference's outputs bit-for-bit on the generating host (a small
tolerance is allowed for a different BLAS / thread count).  This pins the oracle operator by operator, not only end
to end (`tests/test_oracle_golden.py`).
"""
import numpy as np
import pytest
import torch

from diffmvs_b200 import synth
from oracle import diffmvs_ref as O
from oracle import spec
from tests.helpers import GOLDEN_DIR, WEIGHT_SEED, digest

CASES = ("cas_tiny", "cfg1")
TOL = 2e-6


def _case(case):
    g = np.load(f"{GOLDEN_DIR}/ops_{case}.npz")
    args = synth.workload_args(case)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), WEIGHT_SEED)
    assert digest(sd) == str(g["digest"])
    return g, args, sd


def _t(g, k):
    return torch.from_numpy(g[k])


def _close(a, b, tol=TOL):
    assert tuple(a.shape) == tuple(b.shape), (tuple(a.shape), tuple(b.shape))
    err = (a.double() - b.double()).abs().mean() / b.double().abs().mean().clamp_min(1e-30)
    assert err.item() <= tol, err.item()


def _rng(g):
    dv = torch.stack((1.0 / _t(g, "depth_max").view(-1), 1.0 / _t(g, "depth_min").view(-1)), 1)
    return O.DepthRange(dv)


def _features(g, stage):
    n = len([k for k in g.files if k.startswith("feature_v") and k.endswith("_stage1")])
    return [_t(g, f"feature_v{v}_stage{stage}") for v in range(n)]


@pytest.mark.parametrize("case", CASES)
def test_oracle_operators_reproduce_the_reference(case):
    g, args, sd = _case(case)
    cas = args.stage_iters[2] != 0
    imgs, proj, dv = synth.workload_inputs(case)
    with torch.no_grad():
        f0 = O.feature_net(sd, "feature", imgs[0], cas)
        for k, v in f0.items():
            _close(v, _t(g, f"feature_v0_{k}"))
        for k, v in O.context_net(sd, "context", imgs[0], cas).items():
            _close(v, _t(g, f"context_{k}"))
        _close(O.pixel_view_weight(sd, "depthnet.pixel_view_weight", _t(g, "pvw_in")), _t(g, "pvw_out"))
        _close(O.cost_reg_net(sd, "depthnet.cost_regularization", _t(g, "costreg_in")), _t(g, "costreg_out"))
        rng = _rng(g)
        mask, n, depth, vw, conf = O.initial_cost(sd, "depthnet", _features(g, 1), _t(g, "depthnet_context"),
                                                  _t(g, "depthnet_proj"), _t(g, "depthnet_depth_values"), rng,
                                                  args.cost_dim_stage[0])
        for name, got in (("mask", mask), ("inv", n), ("depth", depth), ("view_weights", vw), ("conf", conf)):
            _close(got, _t(g, "depthnet_" + name), 1e-5)
        for call in (1, 2):
            p = f"getcost{call}_"
            c = _t(g, p + "confidence") if (p + "confidence") in g.files else None
            cost, samples = O.get_cost(_t(g, p + "inv"), _features(g, 2), _t(g, p + "proj"), float(g[p + "interval"]), rng,
                                       int(g[p + "costnum"]), _t(g, p + "view_weights"), c, args.cost_dim_stage[1],
                                       args.min_radius, args.max_radius)
            _close(samples, _t(g, p + "samples"))
            _close(cost, _t(g, p + "cost"), 1e-5)
        for s in (2, 3):
            if args.stage_iters[s - 1] == 0:
                continue
            pre = f"update_block_depth{s}"
            _close(O.condition_encoder(sd, pre + ".encoder", _t(g, f"enc{s}_depth"), _t(g, f"enc{s}_samples"),
                                       _t(g, f"enc{s}_cost")), _t(g, f"enc{s}_out"))
            _close(O.sep_conv_gru(sd, pre + ".unet.gru", _t(g, f"gru{s}_h"), _t(g, f"gru{s}_x")), _t(g, f"gru{s}_out"))
            hid, delta, cf = O.unet(sd, pre + ".unet", _t(g, f"unet{s}_in"), _t(g, f"unet{s}_hidden_in"),
                                    _t(g, f"unet{s}_time"), args.unet_dim[s - 1], s)
            _close(hid, _t(g, f"unet{s}_hidden"))
            _close(delta, _t(g, f"unet{s}_delta"), 1e-5)
            _close(cf, _t(g, f"unet{s}_conf"))
        i = 1
        while f"upsample{i}_out" in g.files:
            _close(O.upsample_depth(_t(g, f"upsample{i}_depth"), _t(g, f"upsample{i}_mask"), int(g[f"upsample{i}_ratio"])),
                   _t(g, f"upsample{i}_out"))
            i += 1
