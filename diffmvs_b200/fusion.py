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


def pair_matrices(K_ref, E_ref, K_src, E_src) -> np.ndarray:
    """The 68 doubles `dmvs_geo_consistency` / `dmvs_fuse_view` take for one (reference, source) pair: float32 inverses
    and products exactly as numpy computes them in filter.py:21-47, widened to float64."""
    return _geo_mats(K_ref, E_ref, K_src, E_src)


def _fuse_mats(K_ref, E_ref) -> np.ndarray:
    K_ref, E_ref = _f32(K_ref), _f32(E_ref)
    return np.ascontiguousarray(np.concatenate([np.linalg.inv(K_ref).astype(np.float64).reshape(-1),
                                                np.linalg.inv(E_ref).astype(np.float64).reshape(-1)]))


MAX_SOURCE_VIEWS = 16      # kMaxSrc of csrc/fusion.cu


def _fuse(ref_depth: Tensor, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences, photo_thres, src_views, ref_img,
          *, geo_mask_thres=0, geo_pixel_thres=0.0, geo_depth_thres=0.0, dynamic=None, mats: Optional[Tensor] = None):
    """One launch of `dmvs_fuse_view` (all source views, all thresholds) + the boolean-mask compaction."""
    ref_depth = _req(ref_depth, "ref_depth")
    H, W = ref_depth.shape
    dev = ref_depth.device
    S = len(src_views)
    if S > MAX_SOURCE_VIEWS:
        raise ValueError(f"at most {MAX_SOURCE_VIEWS} source views per reference view (got {S})")
    if len(confidences) > 3 or len(confidences) != len(photo_thres):
        raise ValueError("confidences / photo_thres: up to three maps with one threshold each")
    srcs = [_req(dv, "depth_src") for dv, _, _ in src_views]
    Hs, Ws = (srcs[0].shape if S else (H, W))
    if any(tuple(t.shape) != (Hs, Ws) for t in srcs):
        raise ValueError("all source depth maps of one reference view must have the same size")
    if mats is None:      # host matrix algebra per call; a scan driver passes the pre-uploaded block instead
        blocks = [pair_matrices(ref_intrinsics, ref_extrinsics, K, E) for _, K, E in src_views]
        mats = torch.from_numpy(np.stack(blocks) if blocks else np.zeros((1, 68))).to(dev)
    if mats.dtype != torch.float64 or not mats.is_cuda or mats.numel() < max(S, 1) * 68 or not mats.is_contiguous():
        raise ValueError("mats: contiguous CUDA float64 [S,68] block required")
    confs = [_req(c, "confidence") for c in confidences]
    src_ptrs = (C.c_void_p * max(S, 1))(*[t.data_ptr() for t in srcs])
    conf_ptrs = (C.c_void_p * 3)(*[c.data_ptr() for c in confs])
    thr = (C.c_float * 3)(*[float(np.float32(t)) for t in photo_thres])
    fm = _fuse_mats(ref_intrinsics, ref_extrinsics)
    photo = torch.empty((H, W), device=dev, dtype=torch.uint8)
    geo = torch.empty((H, W), device=dev, dtype=torch.uint8)
    fin = torch.empty((H, W), device=dev, dtype=torch.uint8)
    depth_avg = torch.empty((H, W), device=dev, dtype=torch.float64)
    xyz = torch.empty((H, W, 3), device=dev, dtype=torch.float32)
    if dynamic is None:
        dyn_view_num, dyn_dist, dyn_rel, avg_min, avg_max = 0, 0.0, 0.0, 0.0, 0.0
        dmin, dmax = float(depth_min), float(depth_max)
    else:
        dyn_view_num, dyn_dist, dyn_rel = int(dynamic[0]), float(dynamic[1]), float(dynamic[2])
        if not (1 <= dyn_view_num <= 10) or dyn_dist <= 0 or dyn_rel <= 0:
            raise ValueError(f"dh_pixel_dist_num = {list(dynamic)}: expected [1..10, > 0, > 0]")
        avg_min, avg_max = float(depth_min), float(depth_max)
        dmin, dmax = 0.0, 0.0
    check(_cabi.lib().dmvs_fuse_view(
        ref_depth.data_ptr(), C.cast(src_ptrs, C.c_void_p), mats.data_ptr(), S, H, W, Hs, Ws, C.cast(conf_ptrs, C.c_void_p),
        C.cast(thr, C.c_void_p), len(confs), fm.ctypes.data_as(C.c_void_p), dmin, dmax, float(geo_pixel_thres),
        float(np.float32(geo_depth_thres)), int(geo_mask_thres), dyn_view_num, dyn_dist, dyn_rel, avg_min, avg_max,
        photo.data_ptr(), geo.data_ptr(), fin.data_ptr(), depth_avg.data_ptr(), xyz.data_ptr(), _stream()), "dmvs_fuse_view")
    final = fin.bool()
    out = {"photo_mask": photo.bool(), "geo_mask": geo.bool(), "final_mask": final, "depth_avg": depth_avg,
           "points": xyz[final]}                       # row-major order of the valid pixels, as numpy's boolean index
    if ref_img is not None:
        out["colors"] = (ref_img.to(dev)[final] * 255).to(torch.uint8)
    return out


@_on_device_of(0)
def fuse_view(ref_depth: Tensor, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences: Sequence[Tensor],
              photo_thres: Sequence[float], src_views: Sequence[Tuple[Tensor, object, object]], ref_img: Optional[Tensor] = None,
              geo_mask_thres: int = 3, geo_pixel_thres: float = 1.0, geo_depth_thres: float = 0.01,
              mats: Optional[Tensor] = None):
    """One reference view of `filter_depth` (filter.py:105-215) in one kernel launch.  `src_views`: (depth, intrinsics,
    extrinsics) per source view (at most 16); `confidences` / `photo_thres`: the 2 (DiffMVS) or 3 (CasDiffMVS) confidence
    maps and thresholds; `mats` (optional): the [S,68] float64 CUDA block of `pair_matrices` rows when the caller has
    already uploaded it (scan drivers do, once per scan).  Returns dict: photo_mask, geo_mask, final_mask (bool [H,W]),
    depth_avg (float64 [H,W]), points [N,3] float32 (world), colors [N,3] uint8 (when `ref_img` [H,W,3] in [0,1] is given)."""
    return _fuse(ref_depth, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences, photo_thres, src_views, ref_img,
                 geo_mask_thres=geo_mask_thres, geo_pixel_thres=geo_pixel_thres, geo_depth_thres=geo_depth_thres, mats=mats)


@_on_device_of(0)
def fuse_view_dynamic(ref_depth: Tensor, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences: Sequence[Tensor],
                      photo_thres: Sequence[float], src_views: Sequence[Tuple[Tensor, object, object]],
                      dh_pixel_dist_num: Sequence[int], ref_img: Optional[Tensor] = None, mats: Optional[Tensor] = None):
    """One reference view of `filter_depth_dynamic` (filter.py:230-262, 311-412; Tanks & Temples) in one kernel launch.
    For every source view the consistency test is evaluated for the thresholds i / dh_dist pixels and i / dh_rel_diff
    relative depth, i = dh_view_num .. 10; a pixel is kept when at least i source views pass test i for some i.  The
    loosest test (i = 10) also defines the reprojected depths that are averaged.  `dh_pixel_dist_num` = [dh_view_num,
    dh_dist, dh_rel_diff] (per-scene tables: `DH_VIEW_NUM`, `DH_DIST`, `DH_REL_DIFF` below, filter.py:274-301)."""
    return _fuse(ref_depth, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences, photo_thres, src_views, ref_img,
                 dynamic=dh_pixel_dist_num, mats=mats)


# ------------------------------------------------------------------------------------------------------------------
# Scan-level drivers: the reference's `filter_depth` (filter.py:88-227) and `filter_depth_dynamic` (:262-440) over an
# output directory in the layout test.py:142-200 writes (`scene_io.save_outputs`).  Every file is read once (the
# reference re-reads each source depth map and camera for every pair), all depth maps of the scan live in HBM, the
# matrix algebra of all pairs is done up front on the host (numpy float32, as the reference) and uploaded as one
# block, and each reference view costs one kernel launch.
# ------------------------------------------------------------------------------------------------------------------
# per-scene parameters of the Tanks & Temples dynamic filter (filter.py:274-301)
DH_VIEW_NUM = {'Family': 2, 'Francis': 9, 'Horse': 2, 'Lighthouse': 6, 'M60': 4, 'Panther': 3, 'Playground': 6, 'Train': 3,
               'Auditorium': 2, 'Ballroom': 2, 'Courtroom': 2, 'Museum': 2, 'Palace': 2, 'Temple': 1}
DH_DIST = {'Family': 12, 'Francis': 8, 'Horse': 4, 'Lighthouse': 8, 'M60': 8, 'Panther': 4, 'Playground': 8, 'Train': 4,
           'Auditorium': 4, 'Ballroom': 4, 'Courtroom': 4, 'Museum': 4, 'Palace': 4, 'Temple': 4}
DH_REL_DIFF = {'Family': 1600, 'Francis': 1600, 'Horse': 1300, 'Lighthouse': 1600, 'M60': 1600, 'Panther': 1300,
               'Playground': 1600, 'Train': 1600, 'Auditorium': 1300, 'Ballroom': 1300, 'Courtroom': 1300, 'Museum': 1300,
               'Palace': 1300, 'Temple': 1500}


class ScanMaps:
    """Everything `filter.py` reads for one scan, loaded once: cameras, depth and confidence maps (uploaded to `device`),
    reference images (host)."""

    def __init__(self, out_folder: str, views: Sequence[int], n_conf: int, device):
        from . import scene_io
        self.device = torch.device(device)
        self.views = list(views)
        self.cams = {}
        depth, conf = {}, {}
        self.images = {}
        for v in self.views:
            self.cams[v] = scene_io.read_camera_parameters(f"{out_folder}/cams/{v:0>8}_cam.txt")
            depth[v] = scene_io.read_pfm(f"{out_folder}/depth_est/{v:0>8}.pfm")[0]
        self.depth = {v: torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32)).to(self.device) for v, d in depth.items()}
        self._out, self._n_conf, self._conf = out_folder, n_conf, conf

    def conf(self, v: int) -> List[Tensor]:
        from . import scene_io
        if v not in self._conf:
            maps = [scene_io.read_pfm(f"{self._out}/conf{i}/{v:0>8}.pfm")[0] for i in range(self._n_conf)]
            self._conf[v] = [torch.from_numpy(np.ascontiguousarray(m, dtype=np.float32)).to(self.device) for m in maps]
        return self._conf[v]

    def image(self, v: int) -> Tensor:
        from . import scene_io
        return torch.from_numpy(scene_io.read_img(f"{self._out}/images/{v:0>8}.jpg")).to(self.device)


def _filter_scan(pair_data, out_folder, plyfilename, photo_thres, method, device, *, dynamic=None, geo_mask_thres=3,
                 geo_pixel_thres=1.0, geo_depth_thres=0.01, write_masks=True, verbose=True):
    import os
    from . import scene_io
    n_conf = 3 if method == "casdiffmvs" else 2
    views = sorted({r for r, _ in pair_data} | {s for _, src in pair_data for s in src})
    maps = ScanMaps(out_folder, views, n_conf, device)
    # matrix algebra of every (reference, source) pair, once, on the host (numpy float32 as filter.py:21-47)
    blocks, offset = [], {}
    for ref_view, src_views in pair_data:
        K_ref, E_ref = maps.cams[ref_view][0], maps.cams[ref_view][1]
        offset[ref_view] = len(blocks)
        for s in src_views:
            blocks.append(pair_matrices(K_ref, E_ref, maps.cams[s][0], maps.cams[s][1]))
    mats_all = torch.from_numpy(np.stack(blocks) if blocks else np.zeros((1, 68))).to(maps.device)
    points, colors = [], []
    if write_masks:
        os.makedirs(os.path.join(out_folder, "mask"), exist_ok=True)
    for ref_view, src_views in pair_data:
        K_ref, E_ref, depth_max, depth_min = maps.cams[ref_view]
        src = [(maps.depth[s], maps.cams[s][0], maps.cams[s][1]) for s in src_views]
        mats = mats_all[offset[ref_view]:offset[ref_view] + max(len(src_views), 1)]
        kw = dict(ref_img=maps.image(ref_view), mats=mats)
        if dynamic is None:
            out = fuse_view(maps.depth[ref_view], K_ref, E_ref, depth_max, depth_min, maps.conf(ref_view), photo_thres[:n_conf], src,
                            geo_mask_thres=geo_mask_thres, geo_pixel_thres=geo_pixel_thres, geo_depth_thres=geo_depth_thres, **kw)
        else:
            # filter.py:332-342: with two confidence maps the dynamic filter pairs the second one with the LAST threshold
            thr = photo_thres[:3] if n_conf == 3 else [photo_thres[0], photo_thres[2]]
            out = fuse_view_dynamic(maps.depth[ref_view], K_ref, E_ref, depth_max, depth_min, maps.conf(ref_view), thr, src,
                                    dynamic, **kw)
        if write_masks:
            for kind in ("photo", "geo", "final"):
                scene_io.save_mask(os.path.join(out_folder, f"mask/{ref_view:0>8}_{kind}.png"), out[f"{kind}_mask"].cpu().numpy())
        if verbose:
            print("processing {}, ref-view{:0>2}, photo/geo/final-mask:{}/{}/{}".format(
                out_folder, ref_view, out["photo_mask"].float().mean().item(), out["geo_mask"].float().mean().item(),
                out["final_mask"].float().mean().item()))
        points.append(out["points"])
        colors.append(out["colors"])
    pts = torch.cat(points, 0) if points else torch.zeros((0, 3), device=maps.device)
    cols = torch.cat(colors, 0) if colors else torch.zeros((0, 3), dtype=torch.uint8, device=maps.device)
    if plyfilename is not None:
        scene_io.write_ply(plyfilename, pts.cpu().numpy(), cols.cpu().numpy())
        if verbose:
            print("saving the final model to", plyfilename)
    return pts, cols


def filter_depth(pair_folder, out_folder, plyfilename, geo_mask_thres=3, geo_pixel_thres=1.0, geo_depth_thres=0.01,
                 photo_thres=(0.3, 0.5, 0.5), method="casdiffmvs", dataset="dtu", device="cuda", write_masks=True, verbose=True):
    """Drop-in for the reference's `filter_depth` (filter.py:88-227), same arguments: reads `pair_folder/pair.txt` and the
    maps under `out_folder`, writes `mask/<id>_{photo,geo,final}.png` and the fused cloud `plyfilename`.  Returns the fused
    (points [N,3] float32, colours [N,3] uint8) CUDA tensors (the reference returns nothing)."""
    import os
    from . import scene_io
    pair_data = scene_io.read_pair_file(os.path.join(pair_folder, "pair.txt"), dataset)
    return _filter_scan(pair_data, out_folder, plyfilename, list(photo_thres), method, device, geo_mask_thres=geo_mask_thres,
                        geo_pixel_thres=geo_pixel_thres, geo_depth_thres=geo_depth_thres, write_masks=write_masks,
                        verbose=verbose)


def filter_depth_dynamic(scan, pair_folder, out_folder, plyfilename, photo_thres=(0.3, 0.5, 0.5), method="casdiffmvs",
                         dataset="tank", device="cuda", write_masks=True, verbose=True):
    """Drop-in for the reference's `filter_depth_dynamic` (filter.py:262-440): dynamic thresholds of the Tanks & Temples
    scene `scan` (`DH_*` tables)."""
    import os
    from . import scene_io
    dh = [DH_VIEW_NUM[scan], DH_DIST[scan], DH_REL_DIFF[scan]]
    pair_data = scene_io.read_pair_file(os.path.join(pair_folder, "pair.txt"))       # filter.py:306: default dataset rule
    return _filter_scan(pair_data, out_folder, plyfilename, list(photo_thres), method, device, dynamic=dh,
                        write_masks=write_masks, verbose=verbose)
