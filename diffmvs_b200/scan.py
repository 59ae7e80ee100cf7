"""Scan mode: run the model over the reference views of a scan while keeping each image's FeatureNet pyramid
(SURVEY.md 8(f) row 1 - cross-ref-view feature cache).

`test.py:101-127` processes one reference view per iteration and re-encodes all V images every time
(`diffusion.py:156-157`), although neighbouring reference views of a scan share most of their source images: every
image is encoded about V times per scan.  `ScanRunner` keeps the pyramids of the most recently used images on the
device (LRU by image id, 71.9 MB per 1600x1152 image) and passes them back through the model's `features=` argument,
so FeatureNet runs only on images it has not seen; depth maps are unchanged (`tests/test_gpu_model.py`).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Hashable, List, Optional, Sequence

import torch

Tensor = torch.Tensor


class FeatureCache:
    """LRU cache: image id -> feature pyramid ({"stage1": [B,h,w,C], ...}, channels-last CUDA tensors)."""

    def __init__(self, capacity: int = 16):
        self.capacity = int(capacity)
        self._d: "OrderedDict[Hashable, Dict[str, Tensor]]" = OrderedDict()
        self.hits = self.misses = 0

    def get(self, key: Hashable) -> Optional[Dict[str, Tensor]]:
        f = self._d.get(key)
        if f is None:
            self.misses += 1
            return None
        self._d.move_to_end(key)
        self.hits += 1
        return f

    def put(self, key: Hashable, pyramid: Dict[str, Tensor]) -> None:
        self._d[key] = pyramid
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)

    def __len__(self) -> int:
        return len(self._d)


class ScanRunner:
    """`runner(image_ids, imgs, proj_matrices, depth_values)` = `model(imgs, proj_matrices, depth_values)` with the
    pyramids of already-seen `image_ids` taken from the cache (ids are the caller's image identifiers, e.g. the view
    numbers of `pair.txt`; the same id must always mean the same image at the same size)."""

    def __init__(self, model, capacity: int = 16):
        self.model = model
        self.cache = FeatureCache(capacity)

    def __call__(self, image_ids: Sequence[Hashable], imgs: Sequence[Tensor], proj_matrices, depth_values):
        if len(image_ids) != len(imgs):
            raise ValueError("one id per image")
        feats: List[Optional[Dict[str, Tensor]]] = [self.cache.get(i) for i in image_ids]
        new = [v for v, f in enumerate(feats) if f is None]
        out = self.model(imgs, proj_matrices, depth_values, features=feats if len(new) < len(imgs) else None,
                         return_features=new if new else False)
        for v in new:
            self.cache.put(image_ids[v], out["features"][v])
        out.pop("features", None)
        return out
