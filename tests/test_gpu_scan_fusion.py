"""Scan-level depth filtering + fusion (`diffmvs_b200.fusion.filter_depth` / `filter_depth_dynamic`) against the recorded
outputs of the REFERENCE's own `filter.py` run on the same directory (`tests/golden/scan_fusion.npz`, written by
`oracle/make_scan_golden.py`).  The directory is regenerated here with the same seeded builder; masks are byte work and
must match exactly, the fused vertices to float32 round-off, the colours exactly."""
import os

import numpy as np
import pytest
import torch

from diffmvs_b200 import fusion, scene_io
from tests.helpers import GOLDEN_DIR, write_scan_dir

pytestmark = pytest.mark.gpu
pytest.importorskip("cv2")
GOLD = np.load(os.path.join(GOLDEN_DIR, "scan_fusion.npz"))


@pytest.mark.parametrize("tag,n_conf,method", [("cas", 3, "casdiffmvs"), ("diff", 2, "diffmvs")])
@pytest.mark.parametrize("mode", ["static", "dynamic"])
def test_scan_filter_matches_the_reference(tmp_path, tag, n_conf, method, mode):
    from PIL import Image
    root = str(tmp_path / tag)
    write_scan_dir(root, n_conf=n_conf)
    ply = str(tmp_path / f"{tag}_{mode}.ply")
    if mode == "static":
        pts, cols = fusion.filter_depth(root, root, ply, geo_mask_thres=3, geo_pixel_thres=0.25, geo_depth_thres=0.0005,
                                        photo_thres=[0.3, 0.2, 0.1], method=method, dataset="dtu", verbose=False)
    else:
        pts, cols = fusion.filter_depth_dynamic("M60", root, root, ply, photo_thres=[0.3, 0.2, 0.1], method=method, verbose=False)
    shape = tuple(GOLD[f"{tag}_{mode}_shape"])
    for view in range(5):
        for kind in ("photo", "geo", "final"):
            got = np.array(Image.open(os.path.join(root, f"mask/{view:0>8}_{kind}.png"))) > 0
            ref = np.unpackbits(GOLD[f"{tag}_{mode}_{kind}_{view}"])[:shape[0] * shape[1]].reshape(shape).astype(bool)
            assert np.array_equal(got, ref), (view, kind, int((got != ref).sum()))
    xyz_ref, rgb_ref = GOLD[f"{tag}_{mode}_xyz"], GOLD[f"{tag}_{mode}_rgb"]
    assert tuple(pts.shape) == xyz_ref.shape
    assert np.allclose(pts.cpu().numpy(), xyz_ref, rtol=1e-5, atol=1e-3)
    assert np.array_equal(cols.cpu().numpy(), rgb_ref)
    xyz_file, rgb_file = scene_io.read_ply(ply)                      # the written cloud is what was returned
    assert np.array_equal(xyz_file, pts.cpu().numpy()) and np.array_equal(rgb_file, cols.cpu().numpy())


def test_fused_view_equals_pairwise_composition():
    """The one-launch `fuse_view` against the same result composed from per-pair `check_geometric_consistency` calls."""
    from tests.helpers import plane_scene
    sc = plane_scene(96, 128, 5, 6)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ref_d = t(sc["depth"][0])
    src = [(t(sc["depth"][v]), sc["K"], sc["E"][v]) for v in range(1, 5)]
    out = fusion.fuse_view(ref_d, sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], [t(c) for c in sc["conf"]],
                           [0.3, 0.5, 0.5], src, geo_mask_thres=2, geo_pixel_thres=0.5, geo_depth_thres=0.002)
    cnt = torch.zeros_like(ref_d, dtype=torch.int32)
    acc = torch.zeros_like(ref_d)
    for d, K, E in src:
        m, drep, _, _ = fusion.check_geometric_consistency(ref_d, sc["K"], sc["E"][0], d, K, E, sc["depth_max"], sc["depth_min"],
                                                           0.5, 0.002)
        cnt += m.int()
        acc = acc + drep
    assert torch.equal(out["geo_mask"], cnt >= 2)
    avg = (acc + ref_d).double() / (cnt + 1).double()
    assert torch.equal(out["depth_avg"], avg)
