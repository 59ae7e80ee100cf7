"""Pin the oracle (`oracle/diffmvs_ref.py`) against outputs of the real reference.

The golden files were produced by `oracle/make_golden.py` running `/root/reference`'s own
`CasDiffMVS` on CPU fp32.  The reference has no tests of its own, so these
recorded runs are the pin.  On the machine that generated them the oracle reproduces the
reference bit-for-bit; the tolerance below only allows for CPU-ISA-dependent conv summation
order on other hosts.
"""
import numpy as np
import pytest
import torch

from oracle import diffmvs_ref as O
from tests.helpers import GOLDEN_CASES, case_setup, digest, load_golden, rel_l1, replay_noise

TOL = 2e-5  # rel-L1; observed 0.0 on the generating host


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_weights_regenerate_bit_exact(case):
    g = load_golden(case)
    _, sd, *_ = case_setup(case)
    assert digest(sd) == str(g["digest"]), "seeded weight recipe is not reproducible on this host"


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_forward_matches_reference(case):
    g = load_golden(case)
    args, sd, imgs, proj, dv = case_setup(case)
    taps = {}
    with torch.no_grad():
        out = O.casdiffmvs_forward(sd, args, imgs, proj, dv, randn=replay_noise(g), taps=taps)
    n_depth = len([k for k in g.files if k.startswith("depth_")])
    n_conf = len([k for k in g.files if k.startswith("photo_conf_")])
    assert len(out["depth"]) == n_depth and len(out["photometric_confidence"]) == n_conf
    assert out["conf"] == []
    for i, d in enumerate(out["depth"]):
        ref = torch.from_numpy(g[f"depth_{i}"])
        assert d.shape == ref.shape
        assert rel_l1(d, ref) <= TOL, (case, "depth", i)
    for i, c in enumerate(out["photometric_confidence"]):
        ref = torch.from_numpy(g[f"photo_conf_{i}"])
        assert c.shape == ref.shape
        assert (c - ref).abs().max().item() <= 1e-4, (case, "photo_conf", i)


@pytest.mark.parametrize("case", ["cfg1", "cas_tiny"])
def test_operator_taps_match_reference(case):
    g = load_golden(case)
    args, sd, imgs, proj, dv = case_setup(case)
    t = lambda k: torch.from_numpy(g["tap_" + k])
    # a3 differentiable_warping, both the grid_sample form and the explicit 4-tap form
    w = O.differentiable_warping(t("warp_src"), t("warp_src_proj"), t("warp_ref_proj"), t("warp_depth"))
    assert rel_l1(w, t("warp_out")) <= 1e-6
    we = O.differentiable_warping_explicit(t("warp_src"), t("warp_src_proj"), t("warp_ref_proj"), t("warp_depth"))
    assert rel_l1(we, t("warp_out")) <= 1e-4
    # a9 Unet, first call of stage 2 (and stage 3)
    for s, blk in ((1, "unet2"), (2, "unet3")):
        if "tap_" + blk + "_in" not in g.files:
            continue
        B = t(blk + "_in").shape[0]
        tt = torch.full((B,), 999, dtype=torch.long)
        h, delta, conf = O.unet(sd, f"update_block_depth{s + 1}.unet", t(blk + "_in"), t(blk + "_hidden_in"), tt,
                                args.unet_dim[s], O.UNET_LEVELS[s])
        assert rel_l1(h, t(blk + "_hidden")) <= TOL
        assert rel_l1(delta, t(blk + "_delta")) <= TOL
        assert rel_l1(conf, t(blk + "_conf")) <= TOL
    # whole-model taps
    taps = {}
    with torch.no_grad():
        O.casdiffmvs_forward(sd, args, imgs, proj, dv, randn=replay_noise(g), taps=taps)
    assert rel_l1(taps["stage1_mask"], t("depthnet_mask")) <= TOL
    assert rel_l1(taps["stage1_inv"], t("depthnet_inv")) <= TOL
    assert rel_l1(taps["view_weights"], t("depthnet_view_weights")) <= TOL
    assert (taps["stage1_conf"] - t("depthnet_conf")).abs().max().item() <= 1e-4
    assert rel_l1(taps["stage2_it0_cost"], t("getcost_cost")) <= TOL
    assert rel_l1(taps["stage2_it0_samples"], t("getcost_samples")) <= TOL
    assert rel_l1(taps["stage2_mask"], t("block2_mask")) <= TOL
    assert rel_l1(taps["stage2_hidden0"], t("block2_hidden")) <= TOL


def test_depth_regression_window_semantics():
    """conf = sum of p over [j-1, j+2], zero outside (`module.py:562-571`)."""
    logits = torch.zeros(1, 6, 1, 3)
    logits[0, 0, 0, 0] = 20.0   # peak at plane 0 -> window [-1,2]
    logits[0, 5, 0, 1] = 5.0    # soft peak at last plane: expected index just below 5
    idx, j, conf = O.depth_regression(logits)
    p = torch.softmax(logits, 1)
    assert j[0, 0, 0, 0] == 0 and torch.isclose(conf[0, 0, 0, 0], p[0, 0:3, 0, 0].sum())
    # expected index just below 5 floors to 4 -> window [3,6]
    assert j[0, 0, 0, 1] == 4 and torch.isclose(conf[0, 0, 0, 1], p[0, 3:6, 0, 1].sum())
    assert j[0, 0, 0, 2] == 2  # uniform -> 2.5 -> 2
