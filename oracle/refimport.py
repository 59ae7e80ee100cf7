"""Import the real reference model in the build container (oracle tooling, never on the GPU box).

`/root/reference/models/module.py:7` touches `cuda:0` at import; this shim drops that device
kwarg while the package is imported and leaves `/root/reference`
untouched.  Only `oracle/make_golden.py` and ad-hoc validation scripts use it.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def import_reference_models():
    import torch

    if not reference_available():
        raise RuntimeError("reference tree not mounted at " + REFERENCE_ROOT)
    real_ones = torch.ones

    def ones_no_cuda(*a, **k):
        if not torch.cuda.is_available() and str(k.get("device", "")).startswith("cuda"):
            k = dict(k)
            k.pop("device")
        return real_ones(*a, **k)

    sys.path.insert(0, REFERENCE_ROOT)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        del sys.modules[name]
    torch.ones = ones_no_cuda
    try:
        import models as ref_models  # noqa: F401
        import models.module  # noqa: F401
        import models.update  # noqa: F401
        import models.diffusion  # noqa: F401
    finally:
        torch.ones = real_ones
        sys.path.remove(REFERENCE_ROOT)
    return ref_models
