"""Scan mode: run the model over the reference views of a scan while keeping each image's FeatureNet pyramid
(SURVEY.md 8(f) row 1 - cross-ref-view feature cache).

`test.py:101-127` processes one reference view per iteration and re-encodes all V images every time
(`diffusion.py:156-157`), although neighbouring reference views of a scan share most of their source images: every
image is encoded about V times per scan.  `ScanRunner` keeps the pyramids of the most recently used images on the
device (LRU by image id, 71.9 MB per 1600x1152 image) and passes them back through the model's `features=` argument,
so FeatureNet runs only on images it has not seen; depth maps are unchanged (`tests/test_gpu_model.py`).
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np
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


def scan_metas(testpath: str, scans: Sequence[str], dataset: str = "dtu") -> List[Tuple[str, int, List[int]]]:
    """(scan, reference view, source views) of every reference view the evaluation loader would visit, in its order
    (`MVSDataset.build_metas`, datasets/mvs.py:41-77)."""
    from . import scene_io
    if dataset == "general":
        return [("", r, s) for r, s in scene_io.read_pairs_for_inference(os.path.join(testpath, "pair.txt"), 0.01)]
    metas = []
    for scan in scans:
        metas += [(scan, r, s) for r, s in scene_io.read_pairs_for_inference(os.path.join(testpath, scan, "pair.txt"), 0.1)]
    return metas


def save_scene_depth(model, testpath: str, scans: Sequence[str], outdir: str, num_view: int = 5, numdepth: int = 384,
                     dataset: str = "dtu", max_h: int = 4800, max_w: int = 6400, rank: int = 0, world: int = 1,
                     cache_views: int = 0, device=None, verbose: bool = False) -> float:
    """The per-scene loop of the reference's evaluation script (`save_scene_depth`, test.py:91-205) around an already
    built model: for every reference view of `scans` read the images / cameras (`scene_io.load_sample` = one
    `MVSDataset` item), run `model(imgs, proj_matrices, depth_values)` at batch 1 and write
    `<outdir>/<scan>/{depth_est,cams,images,conf<i>}/<id>.*` exactly as test.py:142-200 does, i.e. the directory
    `filter.py` / `fusion.filter_depth` reads.  Returns the average forward time per reference view in seconds (the
    value test.py prints).

    `rank` / `world`: this process handles its contiguous block of the reference views (`sharding.shard_views`); no
    exchange is needed, every rank writes its own files.  `cache_views` > 0 keeps that many FeatureNet pyramids on the
    device and re-encodes only images not seen before (`ScanRunner`; same depth maps)."""
    from . import scene_io, sharding
    if device is None:
        device = next(model.parameters()).device
    metas = scan_metas(testpath, scans, dataset)
    mine = sharding.shard_views(len(metas), rank, world)
    runner = ScanRunner(model, capacity=cache_views) if cache_views > 0 else None
    time_sum, n_done = 0.0, 0
    with torch.no_grad():
        for idx in mine:
            scan, ref_view, src_views = metas[idx]
            sample = scene_io.load_sample(testpath, scan, ref_view, src_views, num_view, dataset, numdepth, max_w, max_h)
            imgs = [torch.from_numpy(np.ascontiguousarray(i))[None].to(device) for i in sample["imgs"]]
            proj = {k: torch.from_numpy(v)[None].to(device) for k, v in sample["proj_matrices"].items()}
            dv = torch.from_numpy(sample["depth_values"])[None].to(device)
            depth_max = 1.0 / float(sample["depth_values"][0])          # test.py:116-117
            depth_min = 1.0 / float(sample["depth_values"][-1])
            torch.cuda.synchronize(device)
            t0 = time.time()
            if runner is not None:
                ids = [(scan, v) for v in ([ref_view] + list(src_views)[:num_view - 1])]
                out = runner(ids, imgs, proj, dv)
            else:
                out = model(imgs, proj, dv)
            torch.cuda.synchronize(device)
            time_sum += time.time() - t0
            n_done += 1
            depth = out["depth"][-1][0].cpu().numpy()
            confs = [c[0].cpu().numpy() for c in out["photometric_confidence"]]
            cam = sample["proj_matrices"]["stage4"][0]                   # reference camera (test.py:147,155)
            scene_io.save_outputs(outdir, sample["filename"], depth, confs, cam, depth_max, depth_min, img=sample["imgs"][0])
            if verbose:
                print(f"Iter {n_done}/{len(mine)}, Time:{time.time() - t0:.4f} Res:{tuple(depth.shape)}")
    return time_sum / max(n_done, 1)
