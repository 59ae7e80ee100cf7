"""`diffmvs_b200/scene_io.py` against fixtures produced by the reference's own I/O functions and evaluation loader
(`oracle/make_io_golden.py`, run in the build container): PFM and camera files must be byte-identical when written
and bit-identical when read; `pair.txt` selection, intrinsics rescaling, projection-matrix pyramids and the inverse
depth range must equal what `datasets/mvs.py` hands to the model."""
import filecmp
import json
import os

import numpy as np
import pytest

from diffmvs_b200 import scene_io as data_io

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")
META = json.load(open(os.path.join(G, "meta.json")))


def test_pfm_read_and_write_match_the_reference(tmp_path):
    arrs = np.load(os.path.join(G, "pfm_arrays.npz"))
    depth, scale = data_io.read_pfm(os.path.join(G, "depth.pfm"))
    assert scale == 1.0 and depth.dtype == np.float32 and np.array_equal(depth, arrs["depth"])
    color, scale = data_io.read_pfm(os.path.join(G, "color.pfm"))
    assert scale == 2.0 and np.array_equal(color, arrs["color"])
    data_io.save_pfm(str(tmp_path / "d.pfm"), arrs["depth"])
    data_io.save_pfm(str(tmp_path / "c.pfm"), arrs["color"], scale=2)
    assert filecmp.cmp(tmp_path / "d.pfm", os.path.join(G, "depth.pfm"), shallow=False)
    assert filecmp.cmp(tmp_path / "c.pfm", os.path.join(G, "color.pfm"), shallow=False)
    with pytest.raises(ValueError):
        data_io.save_pfm(str(tmp_path / "x.pfm"), arrs["depth"].astype(np.float64))


def test_camera_files_match_the_reference(tmp_path):
    p = np.load(os.path.join(G, "cam_params.npz"))
    data_io.write_cam(str(tmp_path / "cam.txt"), p["cam"], 905.5, 425.0)
    assert filecmp.cmp(tmp_path / "cam.txt", os.path.join(G, "written_cam.txt"), shallow=False)
    intr, ext, dmax, dmin = data_io.read_camera_parameters(os.path.join(G, "written_cam.txt"))
    assert np.array_equal(intr, p["intr"]) and np.array_equal(ext, p["ext"])
    assert [dmax, dmin] == list(p["rng"]) == [935, 425]          # the reference's hard-coded DTU range
    intr, ext, dmax, dmin = data_io.read_camera_parameters(os.path.join(G, "written_cam_small.txt"))
    assert np.array_equal(intr, p["intr2"]) and np.array_equal(ext, p["ext2"]) and [dmax, dmin] == list(p["rng2"])


def test_pair_files_match_the_reference():
    pair = os.path.join(G, "general", "pair.txt")
    assert [[r, s] for r, s in data_io.read_pair_file(pair, "dtu")] == META["pairs_dtu"]
    assert [[r, s] for r, s in data_io.read_pair_file(pair, "eth3d")] == META["pairs_eth3d"]
    assert [[r, s] for r, s in data_io.read_pairs_for_inference(pair, 0.01)] == META["general_metas"]
    assert [[r, s] for r, s in data_io.read_pairs_for_inference(pair, 0.1)] == META["dtu_metas"]


def test_general_scene_sample_equals_the_reference_loader():
    cv2 = pytest.importorskip("cv2")  # noqa: F841  (the reference resizes with OpenCV)
    ref = np.load(os.path.join(G, "general_sample.npz"))
    r, srcs = META["general_metas"][0]
    s = data_io.load_sample(os.path.join(G, "general"), "", r, srcs, n_views=3, dataset="general", numdepth=384)
    assert s["filename"] == META["general_filename"]
    assert np.array_equal(np.stack(s["imgs"]), ref["imgs"])
    assert np.array_equal(s["depth_values"], ref["depth_values"]) and s["depth_values"].dtype == np.float32
    for k in ("stage1", "stage2", "stage3", "stage4"):
        assert s["proj_matrices"][k].dtype == np.float32 and np.array_equal(s["proj_matrices"][k], ref[k]), k


def test_benchmark_scene_sample_equals_the_reference_loader():
    pytest.importorskip("cv2")
    ref = np.load(os.path.join(G, "dtu_sample.npz"))
    r, srcs = META["dtu_metas"][0]
    s = data_io.load_sample(os.path.join(G, "bench"), "scan1", r, srcs, n_views=2, dataset="dtu", numdepth=192)
    imgs = np.stack(s["imgs"])
    assert s["filename"] == META["dtu_filename"]
    assert list(imgs.shape) == list(ref["shape"]) == [2, 3, 1152, 1600]      # DTU views are resized to 1600 x 1152
    assert np.array_equal(imgs[:, :, ::97, ::131], ref["img_probe"])
    assert np.allclose(imgs.mean(axis=(1, 2, 3)), ref["img_mean"], rtol=0, atol=1e-6)
    assert np.array_equal(s["depth_values"], ref["depth_values"])
    for k in ("stage1", "stage2", "stage3", "stage4"):
        assert np.array_equal(s["proj_matrices"][k], ref[k]), k


def test_assembled_sample_feeds_the_model_contract():
    """Shapes / dtypes of SURVEY.md 8(b): V x [3,H,W], stage1..4 -> [V,2,4,4], ascending inverse depths."""
    imgs = [np.zeros((64, 96, 3), np.float32)] * 3
    K = np.array([[100.0, 0, 48], [0, 100.0, 32], [0, 0, 1]], np.float32)
    s = data_io.assemble_sample(imgs, [(K, np.eye(4, dtype=np.float32))] * 3, 425.0, 935.0, 384)
    assert [i.shape for i in s["imgs"]] == [(3, 64, 96)] * 3
    assert all(s["proj_matrices"][f"stage{k}"].shape == (3, 2, 4, 4) for k in (1, 2, 3, 4))
    assert np.allclose(s["proj_matrices"]["stage1"][0, 1, :2, :3], K[:2] * 0.125) and s["proj_matrices"]["stage1"][0, 1, 2, 2] == 1
    dv = s["depth_values"]
    assert dv.shape == (384,) and np.all(np.diff(dv) > 0) and np.isclose(dv[0], 1 / 935.0) and np.isclose(dv[-1], 1 / 425.0)


def test_save_outputs_layout_round_trips(tmp_path):
    depth = np.linspace(400, 900, 24, dtype=np.float32).reshape(4, 6)
    confs = [np.full((4, 6), 0.25 * i, np.float32) for i in range(3)]
    cam = np.load(os.path.join(G, "cam_params.npz"))["cam"]
    data_io.save_outputs(str(tmp_path), "scan1/{}/00000007{}", depth, confs, cam, 905.5, 425.0)
    back, _ = data_io.read_pfm(str(tmp_path / "scan1/depth_est/00000007.pfm"))
    assert np.array_equal(back, depth)
    assert np.array_equal(data_io.read_pfm(str(tmp_path / "scan1/conf2/00000007.pfm"))[0], confs[2])
    intr, ext, _, _ = data_io.read_camera_parameters(str(tmp_path / "scan1/cams/00000007_cam.txt"))
    assert np.array_equal(ext, cam[0]) and np.array_equal(intr, cam[1, :3, :3])
    assert not (tmp_path / "scan1/images").exists()                       # no image given, none written


def test_save_outputs_writes_the_reference_image(tmp_path):
    """test.py:151-162 also writes the resized reference image to `images/<id>.jpg` (clip*255 as uint8, RGB -> BGR,
    cv2.imwrite); filter.py:113,316 reads it with PIL for the point colours, so it must exist, have the depth map's
    size and decode to the same RGB image the reference's own writer produces."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.random((3, 32, 48), dtype=np.float32) * 1.2 - 0.1              # exercises the clip
    depth = np.full((32, 48), 600.0, np.float32)
    cam = np.load(os.path.join(G, "cam_params.npz"))["cam"]
    data_io.save_outputs(str(tmp_path), "scan1/{}/00000003{}", depth, [depth * 0], cam, 935.0, 425.0, img=img)
    path = str(tmp_path / "scan1/images/00000003.jpg")
    back = data_io.read_img(path)                                          # PIL decode, [H,W,3] in [0,1]
    assert back.shape == (32, 48, 3) and back.dtype == np.float32
    # the reference's writer, verbatim arithmetic (test.py:160-162), on the same array
    ref_path = str(tmp_path / "ref.jpg")
    u8 = np.clip(np.transpose(img, (1, 2, 0)) * 255, 0, 255).astype(np.uint8)
    cv2.imwrite(ref_path, cv2.cvtColor(u8, cv2.COLOR_RGB2BGR))
    assert open(path, "rb").read() == open(ref_path, "rb").read()
    with pytest.raises(ValueError):
        data_io.save_outputs(str(tmp_path), "scan1/{}/00000004{}", depth, [], cam, 935.0, 425.0, img=img[:, :16])


def test_written_scan_directory_is_what_the_reference_filter_reads(tmp_path):
    """`oracle/make_scan_golden.py` ran the reference's `filter_depth` on a directory written by `save_outputs`; the
    fixture it recorded exists and carries masks for every view (the GPU test compares against it)."""
    g = np.load(os.path.join(os.path.dirname(G), "scan_fusion.npz"))
    assert {f"cas_static_final_{v}" for v in range(5)} <= set(g.files)
    assert g["cas_static_xyz"].shape[1] == 3 and len(g["cas_static_xyz"]) == len(g["cas_static_rgb"]) > 1000


def test_ply_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    pts = rng.standard_normal((257, 3)).astype(np.float32) * 100
    col = rng.integers(0, 256, size=(257, 3), dtype=np.uint8)
    path = str(tmp_path / "fused.ply")
    data_io.write_ply(path, pts, col)
    head = open(path, "rb").read(200).decode("ascii", "ignore")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 257\nproperty float x\n")
    assert os.path.getsize(path) == head.index("end_header\n") + len("end_header\n") + 257 * 15
    p2, c2 = data_io.read_ply(path)
    assert np.array_equal(p2, pts) and np.array_equal(c2, col)
    with pytest.raises(ValueError):
        data_io.write_ply(path, pts, col[:-1])


def test_scan_metas_equal_the_reference_loader_order():
    """`scan.scan_metas` = `MVSDataset.build_metas` (mvs.py:41-77), recorded by oracle/make_io_golden.py."""
    import json
    from diffmvs_b200 import scan
    meta = json.load(open(os.path.join(G, "meta.json")))
    assert [[r, s] for _, r, s in scan.scan_metas(os.path.join(G, "general"), [""], "general")] == meta["general_metas"]
    assert [[r, s] for _, r, s in scan.scan_metas(os.path.join(G, "bench"), ["scan1"], "dtu")] == meta["dtu_metas"]
