"""Fusion path, CPU side: (1) the oracle restatement of `filter.py` reproduces the reference's own outputs
(`tests/golden/filter.npz`, written by `oracle/make_filter_golden.py` from `/root/reference/filter.py`) bit for bit;
(2) the arithmetic the CUDA kernel uses for `cv2.remap(INTER_LINEAR)` - position rounded to 1/32 pixel, float weight
tables, zero border - is replayed in numpy and compared with OpenCV itself."""
import os

import numpy as np
import pytest

from oracle import filter_ref as F
from tests.helpers import plane_scene

cv2 = pytest.importorskip("cv2")
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter.npz"))


def test_oracle_matches_the_reference_filter():
    sc = plane_scene()
    for v in range(1, len(sc["E"])):
        mask, drep, xs, ys = F.check_geometric_consistency(sc["depth"][0], sc["K"], sc["E"][0], sc["depth"][v], sc["K"],
                                                           sc["E"][v], sc["depth_max"], sc["depth_min"], 1.0, 0.01)
        assert np.array_equal(mask, GOLD[f"mask{v}"])
        assert np.array_equal(drep, GOLD[f"drep{v}"]) and drep.dtype == np.float32
        assert np.array_equal(xs, GOLD[f"xs{v}"]) and np.array_equal(ys, GOLD[f"ys{v}"])


def test_oracle_matches_the_reference_dynamic_filter():
    sc = plane_scene()
    masks, last, drep, _, _ = F.check_geometric_consistency_dynamic(sc["depth"][0], sc["K"], sc["E"][0], sc["depth"][1], sc["K"],
                                                                    sc["E"][1], [2, 12, 1600])
    assert np.array_equal(np.stack(masks), GOLD["dyn_masks"]) and np.array_equal(last, GOLD["dyn_last"])
    assert np.array_equal(drep, GOLD["dyn_drep"])
    assert 0 < GOLD["dyn_masks"][0].mean() < GOLD["dyn_masks"][-1].mean() < 1      # the thresholds actually differ


def emulate_remap(src, mx, my):
    """numpy replay of `remap_linear` in csrc/fusion.cu."""
    Hs, Ws = src.shape
    f32 = np.float32
    fx32, fy32 = (mx * f32(32)).astype(f32), (my * f32(32)).astype(f32)

    def cvround(a):
        ok = (a >= -2147483648.0) & (a < 2147483648.0)
        r = np.where(ok, np.rint(np.where(ok, a, 0)).astype(np.int64), -2147483648)
        return r

    sx, sy = cvround(fx32), cvround(fy32)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    ax, ay = (sx & 31).astype(f32) * f32(1 / 32), (sy & 31).astype(f32) * f32(1 / 32)
    wx0, wy0 = f32(1) - ax, f32(1) - ay
    w = [wy0 * wx0, wy0 * ax, ay * wx0, ay * ax]

    def tap(y, x):
        inside = (x >= 0) & (x < Ws) & (y >= 0) & (y < Hs)
        return np.where(inside, src[np.clip(y, 0, Hs - 1), np.clip(x, 0, Ws - 1)], f32(0)).astype(f32)

    a, b, c, d = tap(iy, ix), tap(iy, ix + 1), tap(iy + 1, ix), tap(iy + 1, ix + 1)
    return ((a * w[0] + b * w[1]) + c * w[2]) + d * w[3]


def test_remap_arithmetic_matches_opencv():
    rng = np.random.default_rng(3)
    src = (rng.random((37, 53), dtype=np.float32) * 500 + 400).astype(np.float32)
    mx = (rng.random((64, 80), dtype=np.float32) * 70 - 8).astype(np.float32)      # partly outside the image
    my = (rng.random((64, 80), dtype=np.float32) * 50 - 6).astype(np.float32)
    mx[0, :4] = [1e9, -1e9, np.nan, 52.999]
    my[1, :3] = [36.0, 36.5, -0.49]
    ref = cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR)
    got = emulate_remap(src, mx, my)
    assert np.allclose(got, ref, rtol=2e-7, atol=0), float(np.abs(got - ref).max())


def test_fuse_view_oracle_shapes():
    sc = plane_scene()
    src = [(sc["depth"][v], sc["K"], sc["E"][v]) for v in range(1, 4)]
    out = F.fuse_view(sc["depth"][0], sc["K"], sc["E"][0], sc["depth_max"], sc["depth_min"], sc["conf"], [0.3, 0.5, 0.5],
                      src, ref_img=sc["img"], geo_mask_thres=2)
    n = int(out["final_mask"].sum())
    assert 0 < n < out["final_mask"].size and out["points"].shape == (n, 3) and out["colors"].shape == (n, 3)
    assert out["depth_avg"].dtype == np.float64
