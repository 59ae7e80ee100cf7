"""Scan directory in -> depth maps -> fused point cloud, entirely through the package (GPU): `scan.save_scene_depth`
(the per-scene loop of the reference's test.py:91-205) feeding `fusion.filter_depth` (filter.py:88-227)."""
import os

import numpy as np
import pytest
import torch

from diffmvs_b200 import fusion, scan, scene_io, synth
from oracle import spec

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "io", "general")      # 3 views, 96 x 128, cams + pair.txt


def _model():
    from diffmvs_b200.models import CasDiffMVS
    args = synth.workload_args("cas_tiny")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    m = CasDiffMVS(args, test=True)
    m.load_state_dict(sd, strict=False)
    return m.cuda().eval()


def _read(outdir, kind, vid, ext=".pfm"):
    return scene_io.read_pfm(os.path.join(outdir, kind, f"{vid:08d}{ext}"))[0]


def test_scene_loop_writes_what_the_model_returns(tmp_path):
    model = _model()
    out = str(tmp_path / "out")
    torch.manual_seed(11)
    avg = scan.save_scene_depth(model, G, [""], out, num_view=3, numdepth=384, dataset="general")
    assert avg > 0
    metas = scan.scan_metas(G, [""], "general")
    assert [m[1] for m in metas] == [0, 1, 2]
    torch.manual_seed(11)                                      # same draws, same order: the files hold the model's outputs
    for _, ref, srcs in metas:
        s = scene_io.load_sample(G, "", ref, srcs, 3, "general", 384)
        imgs = [torch.from_numpy(i)[None].cuda() for i in s["imgs"]]
        proj = {k: torch.from_numpy(v)[None].cuda() for k, v in s["proj_matrices"].items()}
        dv = torch.from_numpy(s["depth_values"])[None].cuda()
        with torch.no_grad():
            o = model(imgs, proj, dv)
        assert np.array_equal(_read(out, "depth_est", ref), o["depth"][-1][0].cpu().numpy())
        for i, c in enumerate(o["photometric_confidence"]):
            assert np.array_equal(_read(out, f"conf{i}", ref), c[0].cpu().numpy())
        for kind, ext in (("cams", "_cam.txt"), ("images", ".jpg")):
            assert os.path.exists(os.path.join(out, kind, f"{ref:08d}{ext}"))
        K, E, dmax, dmin = scene_io.read_camera_parameters(os.path.join(out, "cams", f"{ref:08d}_cam.txt"))
        assert np.allclose(K, s["proj_matrices"]["stage4"][0, 1, :3, :3]) and np.allclose(E, s["proj_matrices"]["stage4"][0, 0])


def test_sharded_and_cached_scene_loops_agree(tmp_path):
    model = _model()
    a, b, c = str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c")
    torch.manual_seed(3)
    scan.save_scene_depth(model, G, [""], a, num_view=3, dataset="general")
    torch.manual_seed(3)                                       # two ranks, one after the other on this GPU
    scan.save_scene_depth(model, G, [""], b, num_view=3, dataset="general", rank=0, world=2)
    scan.save_scene_depth(model, G, [""], b, num_view=3, dataset="general", rank=1, world=2)
    torch.manual_seed(3)
    scan.save_scene_depth(model, G, [""], c, num_view=3, dataset="general", cache_views=8)
    for v in range(3):
        ref = _read(a, "depth_est", v)
        assert np.array_equal(ref, _read(b, "depth_est", v))                      # sharding changes nothing
        got = _read(c, "depth_est", v)
        assert np.abs(got - ref).mean() / np.abs(ref).mean() < 1e-5               # cached pyramids: same maps


def test_scan_directory_to_point_cloud(tmp_path):
    model = _model()
    out = str(tmp_path / "out")
    torch.manual_seed(5)
    scan.save_scene_depth(model, G, [""], out, num_view=3, dataset="general")
    ply = str(tmp_path / "fused.ply")
    pts, cols = fusion.filter_depth(G, out, ply, geo_mask_thres=1, geo_pixel_thres=4.0, geo_depth_thres=0.5,
                                    photo_thres=(0.0, 0.0, 0.0), method="casdiffmvs", dataset="general", verbose=False)
    assert os.path.exists(ply)
    rp, rc = scene_io.read_ply(ply)
    assert rp.shape == tuple(pts.shape) and rc.shape == tuple(cols.shape) and rp.shape[1] == 3
    for v in range(3):
        assert os.path.exists(os.path.join(out, "mask", f"{v:08d}_final.png"))
