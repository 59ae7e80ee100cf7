"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from diffmvs_b200 import synth
from oracle import spec

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = {
    "cfg1": ("cfg1", {}),
    "cas_tiny": ("cas_tiny", {}),
    "cas_tiny_ddim2": ("cas_tiny", dict(sampling_timesteps=[0, 2, 2], ddim_eta=[0, 1.0, 0.5])),
}
WEIGHT_SEED = 123


def load_golden(case: str):
    return np.load(os.path.join(GOLDEN_DIR, case + ".npz"))


def digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def case_setup(case: str):
    """(args, state_dict, imgs, proj, depth_values) of a golden case."""
    workload, over = GOLDEN_CASES[case]
    args = synth.workload_args(workload, **over)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), WEIGHT_SEED)
    imgs, proj, dv = synth.workload_inputs(workload)
    return args, sd, imgs, proj, dv


def replay_noise(golden):
    draws = [torch.from_numpy(golden[k]) for k in sorted((k for k in golden.files if k.startswith("noise_")),
                                                          key=lambda s: int(s.split("_")[1]))]
    it = iter(draws)
    return lambda like: next(it).to(like.device)


def rel_l1(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().mean() / b.abs().mean().clamp_min(1e-30)).item()


def plane_scene(H=48, W=64, V=4, seed=0):
    """Small multi-view depth-map set for the fusion tests: a tilted plane seen by V cameras (shared intrinsics,
    x-translated, slightly rotated), analytic depth per view, plus noise / outliers in some regions so that the
    consistency masks are neither empty nor full.  Returns float32 numpy arrays."""
    import numpy as np
    rng = np.random.default_rng(seed)
    K = np.array([[1.1 * W, 0, W / 2 - 0.5], [0, 1.1 * W, H / 2 - 0.5], [0, 0, 1]], dtype=np.float32)
    n = np.array([0.08, -0.05, 1.0])
    c0 = 600.0                                           # plane n . P = c0 (world)
    Es, depths = [], []
    for v in range(V):
        ang = 0.01 * v
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        t = np.array([-25.0 * v, 3.0 * (v % 2), 0.5 * v])
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, t
        Es.append(E.astype(np.float32))
        # camera-frame ray d = K^-1 (u, v, 1); world point P = R^T (s d - t); n . P = c0  =>  s
        u, w = np.meshgrid(np.arange(W), np.arange(H))
        d = np.linalg.inv(K.astype(np.float64)) @ np.vstack((u.reshape(-1), w.reshape(-1), np.ones(H * W)))
        nr = R @ n                                       # n^T R^T = (R n)^T
        s = (c0 + nr @ t) / (nr @ d)
        depth = s.reshape(H, W)
        depth = depth * (1 + 2e-4 * rng.standard_normal((H, W)))         # within the 1 % consistency band
        depth[H // 3: H // 2, W // 4: W // 2] *= 1.0 + 0.05 * (v % 2)      # inconsistent patch in odd views
        depths.append(depth.astype(np.float32))
    depths[0][:3] = 2000.0                               # rows outside the depth range of the reference view
    confs = [rng.random((H, W), dtype=np.float32) for _ in range(3)]
    img = rng.random((H, W, 3), dtype=np.float32)
    return {"K": K, "E": Es, "depth": depths, "conf": confs, "img": img, "depth_min": 425.0, "depth_max": 935.0}


def write_scan_dir(root: str, H=96, W=128, V=5, seed=4, n_conf=3):
    """A tiny scan in the layout test.py:142-200 writes (through `scene_io.save_outputs`, image included) plus a
    `pair.txt`: the plane scene above seen by V cameras, every view a reference with all others as sources.  Used by the
    scan-level fusion tests and by `oracle/make_scan_golden.py` (which runs the REFERENCE's filter.py on it)."""
    import os
    import numpy as np
    from diffmvs_b200 import scene_io
    sc = plane_scene(H, W, V, seed)
    rng = np.random.default_rng(100 + seed)
    os.makedirs(root, exist_ok=True)
    for v in range(V):
        cam = np.zeros((2, 4, 4), dtype=np.float32)
        cam[0] = sc["E"][v]
        cam[1, :3, :3] = sc["K"]
        confs = [rng.random((H, W), dtype=np.float32) for _ in range(n_conf)]
        img = rng.random((3, H, W), dtype=np.float32)
        scene_io.save_outputs(root, "{}/" + f"{v:0>8}" + "{}", sc["depth"][v], confs, cam, sc["depth_max"], sc["depth_min"], img=img)
    with open(os.path.join(root, "pair.txt"), "w") as f:
        f.write(f"{V}\n")
        for v in range(V):
            others = [u for u in range(V) if u != v]
            f.write(f"{v}\n{len(others)} " + " ".join(f"{u} {100.0 - abs(u - v):.1f}" for u in others) + "\n")
    return sc
