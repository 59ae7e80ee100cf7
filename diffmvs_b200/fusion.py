"""GPU depth-map filtering and fusion - the per-pixel work of the reference's `filter.py`.

`check_geometric_consistency` has the reference's signature (`filter.py:54-87`) on CUDA tensors; `fuse_view` is the
array-level core of `filter_depth` (`:90-227`): photometric mask, geometric consistency against every source view,
averaged depth, final mask and the fused points / colours of one reference view.  File handling (PFM, cams,
pair.txt) is `diffmvs_b200.scene_io`.  The small matrix algebra is done with numpy in float32 exactly as the
reference does, so both implementations feed identical matrices to the per-pixel arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import check

Tensor = torch.Tensor


def _f32(a) -> np.ndarray:
    return np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float32)


def _geo_mats(K_ref, E_ref, K_src, E_src) -> np.ndarray:
    K_ref, E_ref, K_src, E_src = _f32(K_ref), _f32(E_ref), _f32(K_src), _f32(E_src)
    parts = [np.linalg.inv(K_ref), np.matmul(E_src, np.linalg.inv(E_ref)), K_src, np.linalg.inv(K_src),
             np.matmul(E_ref, np.linalg.inv(E_src)), K_ref]          # float32 results, as in filter.py:21-47
    return np.ascontiguousarray(np.concatenate([p.astype(np.float64).reshape(-1) for p in parts]))


def _req(t: Tensor, name: str, dtype=torch.float32) -> Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name}: CUDA tensor required (there is no CPU path)")
    if t.dtype != dtype or t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D {dtype} map, got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device_of(argpos: int = 0):
    """Run the call with the device of its first tensor argument current (kernel launches follow the CUDA current
    device, not the tensors)."""
    def deco(fn):
        def wrapper(*a, **k):
            t = a[argpos]
            if isinstance(t, torch.Tensor) and t.is_cuda and t.device.index != torch.cuda.current_device():
                with torch.cuda.device(t.device):
                    return fn(*a, **k)
            return fn(*a, **k)
        wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
        return wrapper
    return deco


@_on_device_of(0)
def check_geometric_consistency(depth_ref: Tensor, intrinsics_ref, extrinsics_ref, depth_src: Tensor, intrinsics_src,
                                extrinsics_src, ref_depth_max, ref_depth_min, geo_pixel_thres: float = 1.0,
                                geo_depth_thres: float = 0.01, _acc: Optional[Tuple[Tensor, Tensor]] = None):
    """-> (mask bool [H,W], depth_reproj [H,W] (0 where inconsistent), x2d_src, y2d_src), filter.py:54-87."""
    depth_ref, depth_src = _req(depth_ref, "depth_ref"), _req(depth_src, "depth_src")
    H, W = depth_ref.shape
    Hs, Ws = depth_src.shape
    mats = _geo_mats(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src)
    dev = depth_ref.device
    mask = torch.empty((H, W), device=dev, dtype=torch.uint8)
    drep = torch.empty((H, W), device=dev, dtype=torch.float32)
    xs = torch.empty((H, W), device=dev, dtype=torch.float32)
    ys = torch.empty((H, W), device=dev, dtype=torch.float32)
    acc_sum, acc_cnt = _acc if _acc is not None else (None, None)
    check(_cabi.lib().dmvs_geo_consistency(
        depth_ref.data_ptr(), depth_src.data_ptr(), mats.ctypes.data_as(C.c_void_p), float(ref_depth_min),
        float(ref_depth_max), float(geo_pixel_thres), float(np.float32(geo_depth_thres)), mask.data_ptr(), drep.data_ptr(),
        xs.data_ptr(), ys.data_ptr(), None if acc_sum is None else acc_sum.data_ptr(),
        None if acc_cnt is None else acc_cnt.data_ptr(), H, W, Hs, Ws, _stream()), "dmvs_geo_consistency")
    return mask.bool(), drep, xs, ys


@_on_device_of(0)
def fuse_view(ref_depth: Tensor, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences: Sequence[Tensor],
              photo_thres: Sequence[float], src_views: Sequence[Tuple[Tensor, object, object]], ref_img: Optional[Tensor] = None,
              geo_mask_thres: int = 3, geo_pixel_thres: float = 1.0, geo_depth_thres: float = 0.01):
    """One reference view of `filter_depth` (filter.py:105-215).  `src_views`: (depth, intrinsics, extrinsics) per
    source view; `confidences` / `photo_thres`: the 2 (DiffMVS) or 3 (CasDiffMVS) confidence maps and thresholds.
    Returns dict: photo_mask, geo_mask, final_mask (bool [H,W]), depth_avg (float64 [H,W]), points [N,3] float32
    (world), colors [N,3] uint8 (when `ref_img` [H,W,3] in [0,1] is given)."""
    ref_depth = _req(ref_depth, "ref_depth")
    H, W = ref_depth.shape
    dev = ref_depth.device
    photo = torch.ones((H, W), device=dev, dtype=torch.bool)
    for conf, thr in zip(confidences, photo_thres):
        photo &= _req(conf, "confidence") > thr
    acc_sum = torch.zeros((H, W), device=dev, dtype=torch.float32)
    acc_cnt = torch.zeros((H, W), device=dev, dtype=torch.int32)
    for d_src, K_src, E_src in src_views:
        check_geometric_consistency(ref_depth, ref_intrinsics, ref_extrinsics, d_src, K_src, E_src, depth_max, depth_min,
                                    geo_pixel_thres, geo_depth_thres, _acc=(acc_sum, acc_cnt))
    K_ref, E_ref = _f32(ref_intrinsics), _f32(ref_extrinsics)
    mats = np.ascontiguousarray(np.concatenate([np.linalg.inv(K_ref).astype(np.float64).reshape(-1),
                                                np.linalg.inv(E_ref).astype(np.float64).reshape(-1)]))
    depth_avg = torch.empty((H, W), device=dev, dtype=torch.float64)
    geo = torch.empty((H, W), device=dev, dtype=torch.uint8)
    fin = torch.empty((H, W), device=dev, dtype=torch.uint8)
    xyz = torch.empty((H, W, 3), device=dev, dtype=torch.float32)
    photo_u8 = photo.to(torch.uint8)
    check(_cabi.lib().dmvs_fuse_points(ref_depth.data_ptr(), acc_sum.data_ptr(), acc_cnt.data_ptr(), photo_u8.data_ptr(),
                                       int(geo_mask_thres), mats.ctypes.data_as(C.c_void_p), depth_avg.data_ptr(),
                                       geo.data_ptr(), fin.data_ptr(), xyz.data_ptr(), H, W, _stream()), "dmvs_fuse_points")
    final = fin.bool()
    out = {"photo_mask": photo, "geo_mask": geo.bool(), "final_mask": final, "depth_avg": depth_avg,
           "points": xyz[final]}                       # row-major order of the valid pixels, as numpy's boolean index
    if ref_img is not None:
        out["colors"] = (ref_img.to(dev)[final] * 255).to(torch.uint8)
    return out


@_on_device_of(0)
def fuse_view_dynamic(ref_depth: Tensor, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences: Sequence[Tensor],
                      photo_thres: Sequence[float], src_views: Sequence[Tuple[Tensor, object, object]],
                      dh_pixel_dist_num: Sequence[int], ref_img: Optional[Tensor] = None):
    """One reference view of `filter_depth_dynamic` (filter.py:230-262, 311-412; Tanks & Temples).  For every source
    view the consistency test is evaluated for the thresholds i / dh_dist pixels and i / dh_rel_diff relative depth,
    i = dh_view_num .. 10; a pixel is kept when at least i source views pass test i for some i.  The loosest test
    (i = 10) also defines the reprojected depths that are averaged.  `dh_pixel_dist_num` = [dh_view_num, dh_dist,
    dh_rel_diff] (per-scene tables: `oracle.filter_ref.DH_*`, filter.py:274-301)."""
    ref_depth = _req(ref_depth, "ref_depth")
    H, W = ref_depth.shape
    dev = ref_depth.device
    view_num, dh_dist, dh_rel = dh_pixel_dist_num
    photo = torch.ones((H, W), device=dev, dtype=torch.bool)
    for conf, thr in zip(confidences, photo_thres):
        photo &= _req(conf, "confidence") > thr
    acc_sum = torch.zeros((H, W), device=dev, dtype=torch.float32)     # sum of the reprojected depths (test i = 10)
    acc_cnt = torch.zeros((H, W), device=dev, dtype=torch.int32)       # geo_mask_sum (test i = 10)
    level_cnt = {i: torch.zeros((H, W), device=dev, dtype=torch.int32) for i in range(view_num, 10)}
    inf = float("inf")
    for d_src, K_src, E_src in src_views:
        for i in range(view_num, 11):
            last = i == 10
            mask, _, _, _ = check_geometric_consistency(ref_depth, ref_intrinsics, ref_extrinsics, d_src, K_src, E_src, inf, -inf,
                                                        i / dh_dist, i / dh_rel, _acc=(acc_sum, acc_cnt) if last else None)
            if not last:
                level_cnt[i] += mask.to(torch.int32)
    geo = acc_cnt >= 10
    for i in range(view_num, 10):
        geo |= level_cnt[i] >= i
    K_ref, E_ref = _f32(ref_intrinsics), _f32(ref_extrinsics)
    mats = np.ascontiguousarray(np.concatenate([np.linalg.inv(K_ref).astype(np.float64).reshape(-1),
                                                np.linalg.inv(E_ref).astype(np.float64).reshape(-1)]))
    depth_avg = torch.empty((H, W), device=dev, dtype=torch.float64)
    geo_unused = torch.empty((H, W), device=dev, dtype=torch.uint8)
    fin_unused = torch.empty((H, W), device=dev, dtype=torch.uint8)
    xyz = torch.empty((H, W, 3), device=dev, dtype=torch.float32)
    check(_cabi.lib().dmvs_fuse_points(ref_depth.data_ptr(), acc_sum.data_ptr(), acc_cnt.data_ptr(), None, 0,
                                       mats.ctypes.data_as(C.c_void_p), depth_avg.data_ptr(), geo_unused.data_ptr(),
                                       fin_unused.data_ptr(), xyz.data_ptr(), H, W, _stream()), "dmvs_fuse_points")
    in_range = (depth_avg >= depth_min) & (depth_avg <= depth_max)
    final = photo & geo & in_range
    out = {"photo_mask": photo, "geo_mask": geo, "final_mask": final, "depth_avg": depth_avg, "points": xyz[final]}
    if ref_img is not None:
        out["colors"] = (ref_img.to(dev)[final] * 255).to(torch.uint8)
    return out
