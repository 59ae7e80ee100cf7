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
