"""Drop-in for the reference's `models` package (`/root/reference/models/__init__.py:1-2`)."""
from .diffusion import CasDiffMVS  # noqa: F401
from .module import (ContextNet, CostRegNet_small, FeatureNet, GetCost, InitialCost, PixelViewWeight,  # noqa: F401
                     SepConvGRU, depth_to_disp, differentiable_warping, disp_to_depth,
                     get_cur_depth_range_samples, upsample_depth)
from .update import ConditionEncoder, DiffusionUpdateBlockDepth, Unet  # noqa: F401


def compute_inverse_loss(*_a, **_k):
    """Training loss of the reference (`models/loss.py`); training is outside this package's scope."""
    raise NotImplementedError("diffmvs_b200 implements the inference hot path only (SURVEY.md section 2)")
