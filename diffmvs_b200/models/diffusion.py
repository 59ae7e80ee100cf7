"""`CasDiffMVS` - the drop-in for `/root/reference/models/diffusion.py:9-295` (inference path).

    from diffmvs_b200.models import CasDiffMVS            # instead of `from models import *`
    model = CasDiffMVS(args, test=True); model.load_state_dict(ckpt["model"], strict=False)
    model.cuda().eval(); outputs = model(imgs, proj_matrices, depth_values)

Constructor arguments, sub-module names (state-dict keys, including the `update_block.{0,1}` aliases)
and the forward contract are the reference's.  The forward replays `pipeline.CasDiffMVSPlan`.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import pipeline
from .module import Conv2d, ContextNet, FeatureNet, GetCost, InitialCost, _PlannedModule
from .update import DiffusionUpdateBlockDepth


class CasDiffMVS(_PlannedModule):
    """Implementation of DiffMVS and CasDiffMVS"""

    def __init__(self, args, depth_interals_ratio=[4, 2, 1], test=False):
        super().__init__()
        if list(depth_interals_ratio) != [4, 2, 1]:
            raise NotImplementedError("the kernels assume the reference's interval ratios [4,2,1] (diffusion.py:15)")
        self.numdepth_initial = args.numdepth_initial
        self.depth_interals_ratio = depth_interals_ratio
        self.args = args
        self.num_stage = 3
        self.cost_dim_stage = args.cost_dim_stage
        self.unet_dim = args.unet_dim
        self.unet_dim_mults = [(1,), (1, 2), (1, 2, 4)]
        self.test = test
        cas = args.stage_iters[2] != 0
        self.up_ratio = 2 if cas else 4
        self.CostNum = args.CostNum
        self.feat_dim_stage = [48, 32, 16 if cas else 0]
        self.hdim_stage = args.hidden_dim
        self.cdim_stage = args.context_dim
        self.context_dim = [self.hdim_stage[i] + self.cdim_stage[i] for i in range(3)]

        self.feature = FeatureNet(base_channels=8, out_channel=self.feat_dim_stage)
        self.context = ContextNet(self.context_dim, self.hdim_stage)
        inits = [nn.Sequential(Conv2d(self.hdim_stage[1], 32, 3, 2, padding=1),
                               nn.Conv2d(32, self.hdim_stage[1], 3, 1, padding=1, bias=False))]
        if cas:
            inits.append(nn.Sequential(Conv2d(self.hdim_stage[2], 32, 3, 2, padding=1),
                                       Conv2d(32, 32, 3, 2, padding=1),
                                       nn.Conv2d(32, self.hdim_stage[2], 3, 1, padding=1, bias=False)))
        self.hidden_init = nn.ModuleList(inits)

        def block(s):
            return DiffusionUpdateBlockDepth(
                args, dim=self.unet_dim[s], dim_mults=self.unet_dim_mults[s], hidden_dim=self.hdim_stage[s],
                num_sample=self.CostNum[s], cost_dim=self.cost_dim_stage[s] * self.CostNum[s],
                context_dim=self.cdim_stage[s], stage_idx=s, iters=args.stage_iters[s], ratio=self.up_ratio)

        self.update_block_depth2 = block(1)
        if cas:
            self.update_block_depth3 = block(2)
            self.update_block = nn.ModuleList([self.update_block_depth2, self.update_block_depth3])
        else:
            self.update_block = nn.ModuleList([self.update_block_depth2])
        self.depthnet = InitialCost(self.cdim_stage[0], self.cost_dim_stage[0])
        self.GetCost = GetCost(self.cost_dim_stage[1], min_radius=args.min_radius, max_radius=args.max_radius)

    def _build_plan(self, sd, device):
        return pipeline.CasDiffMVSPlan(sd, self.args, device, test=self.test)

    def use_cuda_graph(self, enabled: bool = True) -> "CasDiffMVS":
        """Replay the forward as one CUDA graph per input signature (captured on first use).  Results and
        RNG consumption are those of the eager path; inputs must be CUDA tensors of a fixed shape."""
        self._use_graph = bool(enabled)
        return self

    def forward(self, imgs, proj_matrices, depth_values, depth_gt_ms=None, features=None, return_features=False):
        """`imgs`: list of V `[B,3,H,W]`; `proj_matrices`: dict stage1..4 -> `[B,V,2,4,4]`; `depth_values [B,N]`
        -> {"depth": [...], "conf": [], "photometric_confidence": [...]} (diffusion.py:139-295, test mode).

        Extension (not in the reference): `return_features=True` adds `"features"`, the per-view FeatureNet pyramids;
        passing them back as `features=[pyramid or None per view]` on a later call skips FeatureNet for those views
        (a scan re-uses every image as a source view of its neighbours).  `return_features` may also be a list of view
        indices (e.g. `[0]`: only the new reference image's pyramid).  The returned pyramids are the caller's own
        copies.  Results are unchanged; works with `use_cuda_graph()` (one captured graph per cache pattern)."""
        if not self.test or depth_gt_ms is not None:
            raise NotImplementedError("training-mode outputs (per-iteration lists, ground-truth injection) are out of "
                                      "scope: build with test=True (SURVEY.md section 2)")
        with torch.no_grad():
            plan = self.plan(imgs[0].device)
            if getattr(self, "_use_graph", False):
                return plan.forward_graphed(imgs, proj_matrices, depth_values, features, return_features)
            return plan.forward(imgs, proj_matrices, depth_values, features=features, return_features=return_features)
