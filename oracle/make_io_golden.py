#!/usr/bin/env python
"""Golden fixtures for `diffmvs_b200/scene_io.py`, produced by the REFERENCE's own functions
(`/root/reference/datasets/data_io.py`, `/root/reference/datasets/mvs.py`) in the build container.

    python -m oracle.make_io_golden        # rewrites tests/golden/io/

A tiny synthetic scene (3 views, 96x128 JPEGs, MVSNet-style cams, pair.txt with scores) is written once in both
directory layouts the reference knows ("general": images/, cams/, pair.txt; benchmark: <scan>/images, <scan>/cams_1,
<scan>/pair.txt); the reference's loader, PFM / camera writers and pair-file readers are then run on it and their
outputs stored.  Test infrastructure only - nothing here is imported by the product.
"""
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "io")
REF = "/root/reference"


def main():
    from PIL import Image
    sys.path.insert(0, REF)
    from datasets import data_io as R                     # the reference's functions
    from datasets.mvs import MVSDataset
    sys.path.remove(REF)

    shutil.rmtree(OUT, ignore_errors=True)
    rng = np.random.default_rng(7)
    H, W, V = 96, 128, 3
    for layout in ("general", "bench/scan1"):
        base = os.path.join(OUT, layout)
        cam_dir = "cams" if layout == "general" else "cams_1"
        os.makedirs(os.path.join(base, "images"))
        os.makedirs(os.path.join(base, cam_dir))
    rng_imgs = [rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8) for _ in range(V)]
    for v in range(V):
        ext = np.eye(4, dtype=np.float64)
        ext[0, 3] = 30.0 * v - 17.25
        ext[:3, :3] += 0.01 * rng.standard_normal((3, 3))
        intr = np.array([[231.7, 0.0, 63.4], [0.0, 230.9, 47.8], [0.0, 0.0, 1.0]])
        text = "extrinsic\n" + "\n".join(" ".join(repr(float(x)) for x in row) for row in ext) + "\n\nintrinsic\n" + \
               "\n".join(" ".join(repr(float(x)) for x in row) for row in intr) + "\n\n" + "425.0 2.5 192 905.5\n"
        for layout, cam_dir in (("general", "cams"), ("bench/scan1", "cams_1")):
            base = os.path.join(OUT, layout)
            Image.fromarray(rng_imgs[v]).save(os.path.join(base, "images", f"{v:08d}.jpg"), quality=95)
            with open(os.path.join(base, cam_dir, f"{v:08d}_cam.txt"), "w") as f:
                f.write(text)
    pair = "3\n0\n2 1 0.92 2 0.05\n1\n2 0 0.5 2 0.3\n2\n2 0 0.04 1 0.02\n"
    for layout in ("general", "bench/scan1"):
        with open(os.path.join(OUT, layout, "pair.txt"), "w") as f:
            f.write(pair)

    # ---- the reference's evaluation loader on both layouts ------------------------------------------------
    ds = MVSDataset(os.path.join(OUT, "general"), n_views=3, numdepth=384, dataset="general")
    metas = [list(m[1:]) for m in ds.metas]
    s = ds[0]
    np.savez_compressed(os.path.join(OUT, "general_sample.npz"), imgs=np.stack(s["imgs"]), depth_values=s["depth_values"],
                        **{k: v for k, v in s["proj_matrices"].items()})
    ds2 = MVSDataset(os.path.join(OUT, "bench"), n_views=2, numdepth=192, dataset="dtu", scan=["scan1"])
    s2 = ds2[0]
    imgs2 = np.stack(s2["imgs"])                            # 2 x 3 x 1152 x 1600: keep a digest only
    np.savez_compressed(os.path.join(OUT, "dtu_sample.npz"), depth_values=s2["depth_values"], shape=np.array(imgs2.shape),
                        img_mean=imgs2.mean(axis=(1, 2, 3)), img_probe=imgs2[:, :, ::97, ::131].copy(),
                        **{k: v for k, v in s2["proj_matrices"].items()})
    meta = {"general_metas": metas, "general_filename": s["filename"], "dtu_metas": [list(m[1:]) for m in ds2.metas],
            "dtu_filename": s2["filename"]}

    # ---- PFM / camera writers and the fusion-side readers ------------------------------------------------------
    depth = (rng.random((37, 53), dtype=np.float32) * 500 + 400).astype(np.float32)
    color = rng.random((11, 13, 3), dtype=np.float32)
    R.save_pfm(os.path.join(OUT, "depth.pfm"), depth)
    R.save_pfm(os.path.join(OUT, "color.pfm"), color, scale=2)
    back, scale = R.read_pfm(os.path.join(OUT, "depth.pfm"))
    assert np.array_equal(back, depth) and scale == 1.0
    np.savez_compressed(os.path.join(OUT, "pfm_arrays.npz"), depth=depth, color=color)
    cam = np.asarray(s["proj_matrices"]["stage4"][0])
    R.write_cam(os.path.join(OUT, "written_cam.txt"), cam, 905.5, 425.0)
    intr, ext, dmax, dmin = R.read_camera_parameters(os.path.join(OUT, "written_cam.txt"))
    R.write_cam(os.path.join(OUT, "written_cam_small.txt"), cam, 12.5, 0.75)
    intr2, ext2, dmax2, dmin2 = R.read_camera_parameters(os.path.join(OUT, "written_cam_small.txt"))
    np.savez_compressed(os.path.join(OUT, "cam_params.npz"), cam=cam, intr=intr, ext=ext, rng=np.array([dmax, dmin]),
                        intr2=intr2, ext2=ext2, rng2=np.array([dmax2, dmin2]))
    meta["pairs_dtu"] = [[r, s_] for r, s_ in R.read_pair_file(os.path.join(OUT, "general", "pair.txt"), "dtu")]
    meta["pairs_eth3d"] = [[r, s_] for r, s_ in R.read_pair_file(os.path.join(OUT, "general", "pair.txt"), "eth3d")]
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", OUT, sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
