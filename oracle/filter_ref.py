"""CPU restatement of the per-pixel part of the reference's `filter.py` (test infrastructure; parity pinned against
the reference's own functions by `oracle/make_filter_golden.py` -> `tests/golden/filter.npz`).

`reproject_with_depth` / `check_geometric_consistency` follow `/root/reference/filter.py:8-87` line by line (numpy,
`cv2.remap`); `fuse_view` is the array-level core of `filter_depth` (`:105-215`) without the file handling."""
import numpy as np


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    import cv2
    width, height = depth_ref.shape[1], depth_ref.shape[0]
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x_ref, y_ref = x_ref.reshape([-1]), y_ref.reshape([-1])
    xyz_ref = np.matmul(np.linalg.inv(intrinsics_ref), np.vstack((x_ref, y_ref, np.ones_like(x_ref))) * depth_ref.reshape([-1]))
    xyz_src = np.matmul(np.matmul(extrinsics_src, np.linalg.inv(extrinsics_ref)), np.vstack((xyz_ref, np.ones_like(x_ref))))[:3]
    K_xyz_src = np.matmul(intrinsics_src, xyz_src)
    xy_src = K_xyz_src[:2] / K_xyz_src[2:3]
    x_src = xy_src[0].reshape([height, width]).astype(np.float32)
    y_src = xy_src[1].reshape([height, width]).astype(np.float32)
    sampled_depth_src = cv2.remap(depth_src, x_src, y_src, interpolation=cv2.INTER_LINEAR)
    xyz_src = np.matmul(np.linalg.inv(intrinsics_src), np.vstack((xy_src, np.ones_like(x_ref))) * sampled_depth_src.reshape([-1]))
    xyz_reprojected = np.matmul(np.matmul(extrinsics_ref, np.linalg.inv(extrinsics_src)), np.vstack((xyz_src, np.ones_like(x_ref))))[:3]
    depth_reproj = xyz_reprojected[2].reshape([height, width]).astype(np.float32)
    K_xyz_reprojected = np.matmul(intrinsics_ref, xyz_reprojected)
    K_xyz_reprojected = np.where(K_xyz_reprojected == 0, 1e-5, K_xyz_reprojected)
    xy_reprojected = K_xyz_reprojected[:2] / K_xyz_reprojected[2:3]
    xy_reprojected = np.clip(xy_reprojected, -1e8, 1e8)
    x_reprojected = xy_reprojected[0].reshape([height, width]).astype(np.float32)
    y_reprojected = xy_reprojected[1].reshape([height, width]).astype(np.float32)
    return depth_reproj, x_reprojected, y_reprojected, x_src, y_src


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                ref_depth_max, ref_depth_min, geo_pixel_thres=1.0, geo_depth_thres=0.01):
    width, height = depth_ref.shape[1], depth_ref.shape[0]
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_reproj, x2d_reproj, y2d_reproj, x2d_src, y2d_src = reproject_with_depth(
        depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src)
    dist = np.sqrt((x2d_reproj - x_ref) ** 2 + (y2d_reproj - y_ref) ** 2)
    depth_diff = np.abs(depth_reproj - depth_ref)
    relative_depth_diff = depth_diff / depth_ref
    mask = np.logical_and(dist < geo_pixel_thres, relative_depth_diff < geo_depth_thres)
    mask2 = np.logical_and(depth_ref > ref_depth_min, depth_ref < ref_depth_max)
    mask = np.logical_and(mask, mask2)
    depth_reproj[~mask] = 0
    return mask, depth_reproj, x2d_src, y2d_src


def fuse_view(ref_depth, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences, photo_thres, src_views,
              ref_img=None, geo_mask_thres=3, geo_pixel_thres=1.0, geo_depth_thres=0.01):
    """filter.py:117-215 for one reference view, on arrays."""
    photo_mask = np.ones(ref_depth.shape, dtype=bool)
    for conf, thr in zip(confidences, photo_thres):
        photo_mask = photo_mask & (conf > thr)
    all_depth, geo_mask_sum = [], 0
    for d_src, K_src, E_src in src_views:
        geo_mask, depth_reproj, _, _ = check_geometric_consistency(ref_depth, ref_intrinsics, ref_extrinsics, d_src, K_src,
                                                                   E_src, depth_max, depth_min, geo_pixel_thres, geo_depth_thres)
        geo_mask_sum = geo_mask_sum + geo_mask.astype(np.int32)
        all_depth.append(depth_reproj)
    depth_est_averaged = (sum(all_depth) + ref_depth) / (geo_mask_sum + 1)
    geo_mask = geo_mask_sum >= geo_mask_thres
    final_mask = np.logical_and(photo_mask, geo_mask)
    height, width = depth_est_averaged.shape[:2]
    x, y = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x, y, depth = x[final_mask], y[final_mask], depth_est_averaged[final_mask]
    xyz_ref = np.matmul(np.linalg.inv(ref_intrinsics), np.vstack((x, y, np.ones_like(x))) * depth)
    xyz_world = np.matmul(np.linalg.inv(ref_extrinsics), np.vstack((xyz_ref, np.ones_like(x))))[:3]
    out = {"photo_mask": photo_mask, "geo_mask": geo_mask, "final_mask": final_mask, "depth_avg": depth_est_averaged,
           "points": xyz_world.transpose((1, 0)).astype(np.float32)}
    if ref_img is not None:
        out["colors"] = (ref_img[final_mask] * 255).astype(np.uint8)
    return out


# per-scene parameters of the Tanks & Temples dynamic filter (filter.py:274-301)
DH_VIEW_NUM = {'Family': 2, 'Francis': 9, 'Horse': 2, 'Lighthouse': 6, 'M60': 4, 'Panther': 3, 'Playground': 6, 'Train': 3,
               'Auditorium': 2, 'Ballroom': 2, 'Courtroom': 2, 'Museum': 2, 'Palace': 2, 'Temple': 1}
DH_DIST = {'Family': 12, 'Francis': 8, 'Horse': 4, 'Lighthouse': 8, 'M60': 8, 'Panther': 4, 'Playground': 8, 'Train': 4,
           'Auditorium': 4, 'Ballroom': 4, 'Courtroom': 4, 'Museum': 4, 'Palace': 4, 'Temple': 4}
DH_REL_DIFF = {'Family': 1600, 'Francis': 1600, 'Horse': 1300, 'Lighthouse': 1600, 'M60': 1600, 'Panther': 1300,
               'Playground': 1600, 'Train': 1600, 'Auditorium': 1300, 'Ballroom': 1300, 'Courtroom': 1300, 'Museum': 1300,
               'Palace': 1300, 'Temple': 1500}


def check_geometric_consistency_dynamic(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                        dh_pixel_dist_num):
    """filter.py:230-262 (dynamic filtering for Tanks & Temples, after D2HC-RMVSNet)."""
    width, height = depth_ref.shape[1], depth_ref.shape[0]
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_reproj, x2d_reproj, y2d_reproj, x2d_src, y2d_src = reproject_with_depth(
        depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src)
    dist = np.sqrt((x2d_reproj - x_ref) ** 2 + (y2d_reproj - y_ref) ** 2)
    depth_diff = np.abs(depth_reproj - depth_ref)
    relative_depth_diff = depth_diff / depth_ref
    masks = []
    for i in range(dh_pixel_dist_num[0], 11):
        mask = np.logical_and(dist < i / dh_pixel_dist_num[1], relative_depth_diff < i / dh_pixel_dist_num[2])
        masks.append(mask)
    depth_reproj[~mask] = 0
    return masks, mask, depth_reproj, x2d_src, y2d_src


def fuse_view_dynamic(ref_depth, ref_intrinsics, ref_extrinsics, depth_max, depth_min, confidences, photo_thres, src_views,
                      dh_pixel_dist_num, ref_img=None):
    """filter.py:311-412 for one reference view, on arrays (`photo_thres` already selected per method)."""
    dh_view_num = dh_pixel_dist_num[0]
    photo_mask = np.ones(ref_depth.shape, dtype=bool)
    for conf, thr in zip(confidences, photo_thres):
        photo_mask = photo_mask & (conf > thr)
    all_depth, geo_mask_sum, geo_mask_sums = [], 0, []
    for ct, (d_src, K_src, E_src) in enumerate(src_views):
        masks, geo_mask, depth_reproj, _, _ = check_geometric_consistency_dynamic(
            ref_depth, ref_intrinsics, ref_extrinsics, d_src, K_src, E_src, dh_pixel_dist_num)
        if ct == 0:
            geo_mask_sums = [m.astype(np.int32) for m in masks]
        else:
            for k, m in enumerate(masks):
                geo_mask_sums[k] += m.astype(np.int32)
        geo_mask_sum = geo_mask_sum + geo_mask.astype(np.int32)
        all_depth.append(depth_reproj)
    geo_mask = geo_mask_sum >= 10
    for i in range(dh_view_num, 11):
        geo_mask = np.logical_or(geo_mask, geo_mask_sums[i - dh_view_num] >= i)
    depth_est_averaged = (sum(all_depth) + ref_depth) / (geo_mask_sum + 1)
    maskdepth = np.logical_and(depth_est_averaged >= depth_min, depth_est_averaged <= depth_max)
    final_mask = np.logical_and(np.logical_and(photo_mask, geo_mask), maskdepth)
    height, width = depth_est_averaged.shape[:2]
    x, y = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x, y, depth = x[final_mask], y[final_mask], depth_est_averaged[final_mask]
    xyz_ref = np.matmul(np.linalg.inv(ref_intrinsics), np.vstack((x, y, np.ones_like(x))) * depth)
    xyz_world = np.matmul(np.linalg.inv(ref_extrinsics), np.vstack((xyz_ref, np.ones_like(x))))[:3]
    out = {"photo_mask": photo_mask, "geo_mask": geo_mask, "final_mask": final_mask, "depth_avg": depth_est_averaged,
           "points": xyz_world.transpose((1, 0)).astype(np.float32)}
    if ref_img is not None:
        out["colors"] = (ref_img[final_mask] * 255).astype(np.uint8)
    return out
