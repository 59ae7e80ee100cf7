"""GPU depth-map filtering / fusion (diffmvs_b200/fusion.py) against the CPU oracle of the reference's `filter.py`
(`oracle/filter_ref.py`, pinned to the reference by `tests/test_filter_oracle.py`).  The oracle computes in float64 numpy
around OpenCV's float32 remap; the kernels use the same dtypes and operation order, so masks must agree exactly (they are
byte work: tolerance 0; thresholds are compared in the reference's own dtypes - the float64 pixel distance against the
float64 threshold, the float32 relative depth difference against the float32-rounded threshold) and values to a few ulps."""
import numpy as np
import pytest
import torch

from diffmvs_b200 import fusion
from oracle import filter_ref as F
from tests.helpers import plane_scene

pytestmark = pytest.mark.gpu
pytest.importorskip("cv2")
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("H,W,seed", [(48, 64, 0), (120, 200, 1)])
def test_geometric_consistency_matches_oracle(H, W, seed):
    sc = plane_scene(H, W, 4, seed)
    for v in range(1, 4):
        mask_r, drep_r, xs_r, ys_r = F.check_geometric_consistency(sc["depth"][0], sc["K"], sc["E"][0], sc["depth"][v], sc["K"],
                                                                   sc["E"][v], sc["depth_max"], sc["depth_min"], 1.0, 0.01)
        mask, drep, xs, ys = fusion.check_geometric_consistency(_t(sc["depth"][0]), sc["K"], sc["E"][0], _t(sc["depth"][v]),
                                                                sc["K"], sc["E"][v], sc["depth_max"], sc["depth_min"], 1.0, 0.01)
        mask, drep, xs, ys = mask.cpu().numpy(), drep.cpu().numpy(), xs.cpu().numpy(), ys.cpu().numpy()
        assert np.array_equal(mask, mask_r), (v, int((mask != mask_r).sum()))     # masks are byte work: exact
        both = mask & mask_r
        assert np.allclose(drep[both], drep_r[both], rtol=2e-6, atol=0)
        assert np.all(drep[~mask] == 0)
        assert np.allclose(xs, xs_r, rtol=1e-6, atol=1e-4) and np.allclose(ys, ys_r, rtol=1e-6, atol=1e-4)


def test_fuse_view_matches_oracle():
    sc = plane_scene(96, 128, 4, 2)
    src_np = [(sc["depth"][v], sc["K"], sc["E"][v]) for v in range(1, 4)]
    ref = F.fuse_view(sc["depth"][0], sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], sc["conf"], [0.3, 0.5, 0.5], src_np,
                      ref_img=sc["img"], geo_mask_thres=2)
    src = [(_t(d), K, E) for d, K, E in src_np]
    out = fusion.fuse_view(_t(sc["depth"][0]), sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], [_t(c) for c in sc["conf"]],
                           [0.3, 0.5, 0.5], src, ref_img=_t(sc["img"]), geo_mask_thres=2)
    for k in ("photo_mask", "geo_mask", "final_mask"):
        assert np.array_equal(out[k].cpu().numpy(), ref[k]), k
    assert np.allclose(out["depth_avg"].cpu().numpy(), ref["depth_avg"], rtol=1e-6, atol=0)
    assert out["points"].shape == ref["points"].shape
    assert np.allclose(out["points"].cpu().numpy(), ref["points"], rtol=1e-5, atol=1e-3)
    assert np.array_equal(out["colors"].cpu().numpy(), ref["colors"])


def test_dynamic_fuse_view_matches_oracle():
    """Tanks & Temples variant (filter.py:230-262, 311-412): thresholds i/dh_dist, i/dh_rel_diff for i = 2..10."""
    sc = plane_scene(96, 128, 4, 3)
    dh = [2, 12, 1600]
    src_np = [(sc["depth"][v], sc["K"], sc["E"][v]) for v in range(1, 4)]
    ref = F.fuse_view_dynamic(sc["depth"][0], sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], sc["conf"], [0.3, 0.5, 0.5],
                              src_np, dh, ref_img=sc["img"])
    out = fusion.fuse_view_dynamic(_t(sc["depth"][0]), sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"],
                                   [_t(c) for c in sc["conf"]], [0.3, 0.5, 0.5], [(_t(d), K, E) for d, K, E in src_np], dh,
                                   ref_img=_t(sc["img"]))
    for k in ("photo_mask", "geo_mask", "final_mask"):
        assert np.array_equal(out[k].cpu().numpy(), ref[k]), (k, int((out[k].cpu().numpy() != ref[k]).sum()))
    assert np.allclose(out["depth_avg"].cpu().numpy(), ref["depth_avg"], rtol=1e-6, atol=0)
    assert np.allclose(out["points"].cpu().numpy(), ref["points"], rtol=1e-5, atol=1e-3)
    assert np.array_equal(out["colors"].cpu().numpy(), ref["colors"])
    assert 0 < ref["final_mask"].mean() < 1


def test_cpu_tensors_are_rejected():
    sc = plane_scene()
    with pytest.raises(ValueError):
        fusion.check_geometric_consistency(torch.from_numpy(sc["depth"][0]), sc["K"], sc["E"][0], torch.from_numpy(sc["depth"][1]),
                                           sc["K"], sc["E"][1], 935.0, 425.0)
