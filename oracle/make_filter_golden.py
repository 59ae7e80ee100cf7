#!/usr/bin/env python
"""Golden vectors for the fusion path from the REFERENCE's own `filter.py` (`check_geometric_consistency`,
`reproject_with_depth`), run in the build container on the synthetic plane scene of tests/helpers.py.

    python -m oracle.make_filter_golden        # rewrites tests/golden/filter.npz

`filter.py` imports `plyfile` (not installed) only to write the final PLY; a stub module satisfies the import.
Test infrastructure only."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests.helpers import plane_scene
    sys.modules.setdefault("plyfile", types.SimpleNamespace(PlyData=None, PlyElement=None))
    sys.path.insert(0, "/root/reference")
    import filter as ref_filter                          # the reference module
    sys.path.remove("/root/reference")
    sc = plane_scene()
    out = {}
    for v in range(1, len(sc["E"])):
        mask, drep, xs, ys = ref_filter.check_geometric_consistency(
            sc["depth"][0], sc["K"], sc["E"][0], sc["depth"][v], sc["K"], sc["E"][v], sc["depth_max"], sc["depth_min"], 1.0, 0.01)
        out[f"mask{v}"], out[f"drep{v}"], out[f"xs{v}"], out[f"ys{v}"] = mask, drep, xs, ys
    dh = [2, 12, 1600]                                   # 'Family' (filter.py:274-301)
    masks, last, drep, _, _ = ref_filter.check_geometric_consistency_dynamic(
        sc["depth"][0], sc["K"], sc["E"][0], sc["depth"][1], sc["K"], sc["E"][1], dh)
    out["dyn_masks"], out["dyn_last"], out["dyn_drep"] = np.stack(masks), last, drep
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "filter.npz"), **out)
    print({k: (v.shape, float(np.mean(v))) for k, v in out.items() if k.startswith("mask")})


if __name__ == "__main__":
    main()
