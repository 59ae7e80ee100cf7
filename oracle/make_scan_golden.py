#!/usr/bin/env python
"""Scan-level golden outputs of the REFERENCE's `filter_depth` / `filter_depth_dynamic` (filter.py:88-227, 262-440).

    python -m oracle.make_scan_golden        # rewrites tests/golden/scan_fusion.npz (build container only)

A tiny synthetic scan is written with the PRODUCT's `scene_io.save_outputs` (tests/helpers.write_scan_dir); the
reference's own `filter.py` then reads that directory unmodified - which also proves the written layout (cams, PFM maps,
`images/<id>.jpg`) is what the reference consumes - and its masks (read back from the PNG files it writes) and fused
vertices are recorded.  `plyfile` is not installed: a stub captures the vertex array `filter.py` hands to
`PlyData([el]).write`.  Test infrastructure only."""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CAPTURED = {}


class _PlyElement:
    @staticmethod
    def describe(arr, name):
        return arr


class _PlyData:
    def __init__(self, els):
        self.els = els

    def write(self, filename):
        CAPTURED[os.path.basename(filename)] = np.array(self.els[0])


def run(tmp):
    from PIL import Image
    from tests.helpers import write_scan_dir
    sys.modules["plyfile"] = types.SimpleNamespace(PlyData=_PlyData, PlyElement=_PlyElement)
    sys.path.insert(0, "/root/reference")
    import filter as ref_filter                          # the reference module, unmodified
    sys.path.remove("/root/reference")
    out = {}
    for tag, n_conf, method in (("cas", 3, "casdiffmvs"), ("diff", 2, "diffmvs")):
        root = os.path.join(tmp, tag)
        write_scan_dir(root, n_conf=n_conf)
        for mode in ("static", "dynamic"):
            ply = os.path.join(tmp, f"{tag}_{mode}.ply")
            if mode == "static":
                ref_filter.filter_depth(root, root, ply, geo_mask_thres=3, geo_pixel_thres=0.25, geo_depth_thres=0.0005,
                                        photo_thres=[0.3, 0.2, 0.1], method=method, dataset="dtu")
            else:
                ref_filter.filter_depth_dynamic("M60", root, root, ply, photo_thres=[0.3, 0.2, 0.1], method=method)
            v = CAPTURED[os.path.basename(ply)]
            out[f"{tag}_{mode}_xyz"] = np.stack((v["x"], v["y"], v["z"]), 1)
            out[f"{tag}_{mode}_rgb"] = np.stack((v["red"], v["green"], v["blue"]), 1)
            for view in range(5):
                for kind in ("photo", "geo", "final"):
                    m = np.array(Image.open(os.path.join(root, f"mask/{view:0>8}_{kind}.png"))) > 0
                    out[f"{tag}_{mode}_{kind}_{view}"] = np.packbits(m)
            out[f"{tag}_{mode}_shape"] = np.array(m.shape)
    return out


def main():
    with tempfile.TemporaryDirectory() as tmp:
        out = run(tmp)
    path = os.path.join(ROOT, "tests", "golden", "scan_fusion.npz")
    np.savez_compressed(path, **out)
    for k in sorted(out):
        if k.endswith("_xyz"):
            print(k, out[k].shape)
    print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
