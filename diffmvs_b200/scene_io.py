"""File formats on either side of the hot path (SURVEY.md 8(b) input spec, 8(f) row 3): PFM maps, MVSNet-style
camera files, `pair.txt`, and the assembly of the model's inputs from a reference view and its source views.

Behaviour follows `/root/reference/datasets/data_io.py` (PFM :59-122, `write_cam` :124-141,
`read_camera_parameters` :143-163, `read_pair_file` :172-190) and `/root/reference/datasets/mvs.py`
(`build_metas` :41-77, `read_cam_file` :79-91, resizing :99-124, `__getitem__` :129-210) so that files written by
either implementation are read identically by the other; `tests/test_scene_io.py` checks this against fixtures
produced by the reference's own functions (`oracle/make_io_golden.py`).  Pure host code (numpy); image resizing
uses OpenCV exactly like the reference when it is installed.
"""
from __future__ import annotations

import os
import re
import sys
from typing import Optional, Dict, List, Sequence, Tuple

import numpy as np

# sizes the reference resizes each benchmark to (mvs.py:29-34)
FIXED_SIZES = {"dtu": (1600, 1152), "tank": (1920, 1056), "eth3d": (1920, 1280)}


# ------------------------------------------------------------------------------------------------
# PFM (data_io.py:59-122): text header "Pf|PF", "W H", scale (negative = little endian), rows bottom-up
# ------------------------------------------------------------------------------------------------
def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header == "PF":
            color = True
        elif header == "Pf":
            color = False
        else:
            raise ValueError(f"{filename}: not a PFM file")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise ValueError(f"{filename}: malformed PFM header")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        scale = abs(scale)
        data = np.fromfile(f, endian + "f")
    shape = (height, width, 3) if color else (height, width)
    return np.flipud(np.reshape(data, shape)), scale


def save_pfm(filename: str, image: np.ndarray, scale: float = 1) -> None:
    image = np.flipud(image)
    if image.dtype.name != "float32":
        raise ValueError("save_pfm: image dtype must be float32")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise ValueError("save_pfm: image must have H x W x 3, H x W x 1 or H x W dimensions")
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        image.tofile(f)


# ------------------------------------------------------------------------------------------------
# camera files: "extrinsic" + 4x4, blank, "intrinsic" + 3x3, blank, depth range line
# ------------------------------------------------------------------------------------------------
def _cam_lines(filename: str) -> List[str]:
    with open(filename) as f:
        return [line.rstrip() for line in f.readlines()]


def _floats(text: str) -> np.ndarray:
    return np.array(text.split(), dtype=np.float32)


def read_cam_file(filename: str) -> Tuple[np.ndarray, np.ndarray, float, float]:
    """Input cameras as the evaluation loader reads them (mvs.py:79-91): (intrinsics 3x3, extrinsics 4x4,
    depth_min, depth_max) with depth_min = first and depth_max = LAST number of line 11, negative depth_min -> 1."""
    lines = _cam_lines(filename)
    extrinsics = _floats(" ".join(lines[1:5])).reshape(4, 4)
    intrinsics = _floats(" ".join(lines[7:10])).reshape(3, 3)
    depth_min = float(lines[11].split()[0])
    depth_max = float(lines[11].split()[-1])
    if depth_min < 0:
        depth_min = 1.0
    return intrinsics, extrinsics, depth_min, depth_max


def read_camera_parameters(filename: str) -> Tuple[np.ndarray, np.ndarray, float, float]:
    """Cameras written next to the depth maps, as the fusion step reads them (data_io.py:143-163): the range line
    is "depth_max depth_min" and a depth_max above 425 is replaced by the hard-coded DTU range [425, 935]."""
    lines = _cam_lines(filename)
    extrinsics = _floats(" ".join(lines[1:5])).reshape(4, 4)
    intrinsics = _floats(" ".join(lines[7:10])).reshape(3, 3)
    depth_min = float(lines[11].split()[1])
    depth_max = float(lines[11].split()[0])
    if depth_max > 425:
        depth_max, depth_min = 935, 425
    return intrinsics, extrinsics, depth_max, depth_min


def write_cam(filename: str, cam: np.ndarray, depth_max, depth_min) -> None:
    """`cam` [2,4,4] = (extrinsic, intrinsic in the upper-left 3x3), as test.py:159 passes it (data_io.py:124-141)."""
    with open(filename, "w") as f:
        f.write("extrinsic\n")
        for i in range(4):
            for j in range(4):
                f.write(str(cam[0][i][j]) + " ")
            f.write("\n")
        f.write("\n")
        f.write("intrinsic\n")
        for i in range(3):
            for j in range(3):
                f.write(str(cam[1][i][j]) + " ")
            f.write("\n")
        f.write("\n" + str(depth_max) + " " + str(depth_min) + "\n")


# ------------------------------------------------------------------------------------------------
# pair.txt
# ------------------------------------------------------------------------------------------------
def read_pair_file(filename: str, dataset: str = "dtu") -> List[Tuple[int, List[int]]]:
    """View selection as the fusion step reads it (data_io.py:172-190): every listed source view for DTU-style
    files, score > 0.1 and != ref for "eth3d"; reference views without sources are dropped."""
    data = []
    with open(filename) as f:
        num_viewpoint = int(f.readline())
        for _ in range(num_viewpoint):
            ref_view = int(f.readline().rstrip())
            if dataset != "eth3d":
                src_views = [int(x) for x in f.readline().rstrip().split()[1::2]]
            else:
                fields = [float(x) for x in f.readline().rstrip().split()]
                ids, score = [int(x) for x in fields[1::2]], fields[2::2]
                src_views = [v for v, s in zip(ids, score) if s > 0.1 and v != ref_view]
            if len(src_views) > 0:
                data.append((ref_view, src_views))
    return data


def read_pairs_for_inference(filename: str, min_score: float = 0.1) -> List[Tuple[int, List[int]]]:
    """View selection as the evaluation loader reads it (mvs.py:41-77): sources with score > min_score (0.1 for the
    benchmarks, 0.01 for "general" scenes) that differ from the reference view."""
    metas = []
    with open(filename) as f:
        num_viewpoint = int(f.readline())
        for _ in range(num_viewpoint):
            ref_view = int(f.readline().rstrip())
            fields = [float(x) for x in f.readline().rstrip().split()]
            ids, score = [int(x) for x in fields[1::2]], fields[2::2]
            src_views = [v for v, s in zip(ids, score) if s > min_score and v != ref_view]
            if len(src_views) != 0:
                metas.append((ref_view, src_views))
    return metas


# ------------------------------------------------------------------------------------------------
# images and the model's inputs
# ------------------------------------------------------------------------------------------------
def read_img(filename: str) -> np.ndarray:
    """RGB image as float32 in [0, 1], H x W x 3 (mvs.py:93-97); no mean / std normalisation."""
    from PIL import Image
    return np.array(Image.open(filename), dtype=np.float32) / 255.0


def adaptive_size(h: int, w: int, max_w: int = 6400, max_h: int = 4800, base: int = 32) -> Tuple[int, int]:
    """(new_w, new_h) of `scale_img_adaptive` (mvs.py:104-115): shrink to the maxima, round down to multiples of 32."""
    if h > max_h or w > max_w:
        new_w, new_h = (1.0 * max_w / w) * w // base * base, (1.0 * max_h / h) * h // base * base
    else:
        new_w, new_h = 1.0 * w // base * base, 1.0 * h // base * base
    return int(new_w), int(new_h)


def resize_view(img: np.ndarray, intrinsics: np.ndarray, dataset: str, max_w: int = 6400, max_h: int = 4800):
    """Resize one view and rescale its intrinsics like the evaluation loader (mvs.py:99-124,147-153)."""
    import cv2
    h, w = img.shape[:2]
    intrinsics = intrinsics.copy()
    if dataset in FIXED_SIZES:
        wh = FIXED_SIZES[dataset]
        img = cv2.resize(img, wh, interpolation=cv2.INTER_LINEAR)
        intrinsics[0] *= wh[0] / w
        intrinsics[1] *= wh[1] / h
    else:
        new_w, new_h = adaptive_size(h, w, max_w, max_h)
        intrinsics[0, :] *= 1.0 * new_w / w
        intrinsics[1, :] *= 1.0 * new_h / h
        img = cv2.resize(img, (new_w, new_h))
    return img, intrinsics


def assemble_sample(imgs: Sequence[np.ndarray], cams: Sequence[Tuple[np.ndarray, np.ndarray]], depth_min: float,
                    depth_max: float, numdepth: int = 384) -> Dict[str, object]:
    """The model's inputs from V views (reference first): `imgs` H x W x 3 in [0,1] (already resized), `cams` =
    (intrinsics, extrinsics) per view, depth range of the reference view.  Returns the dict of mvs.py:187-203:
    "imgs" list of [3,H,W], "proj_matrices" stage1..stage4 -> [V,2,4,4] (intrinsic rows 0-1 scaled by 1/8, 1/4,
    1/2, 1), "depth_values" = linspace(1/depth_max, 1/depth_min, numdepth) (inverse depth, ascending)."""
    proj = []
    for intrinsics, extrinsics in cams:
        m = np.zeros((2, 4, 4), dtype=np.float32)
        m[0, :4, :4] = extrinsics
        m[1, :3, :3] = intrinsics
        proj.append(m)
    proj = np.stack(proj)
    stages = {}
    for name, s in (("stage1", 0.125), ("stage2", 0.25), ("stage3", 0.5)):
        p = proj.copy()
        p[:, 1, :2, :] = proj[:, 1, :2, :] * s
        stages[name] = p
    stages["stage4"] = proj
    depth_values = np.linspace(1.0 / depth_max, 1.0 / depth_min, numdepth, dtype=np.float32)
    return {"imgs": [im.transpose([2, 0, 1]) for im in imgs], "proj_matrices": stages, "depth_values": depth_values}


def load_sample(datapath: str, scan: str, ref_view: int, src_views: Sequence[int], n_views: int, dataset: str = "dtu",
                numdepth: int = 384, max_w: int = 6400, max_h: int = 4800) -> Dict[str, object]:
    """One item of the reference's evaluation dataset (mvs.py:129-210) read from the same directory layout."""
    view_ids = [ref_view] + list(src_views)[:n_views - 1]
    general = dataset == "general"
    root = datapath if general else os.path.join(datapath, scan)
    cam_folder = "cams" if general else "cams_1"
    imgs, cams, rng = [], [], None
    for i, vid in enumerate(view_ids):
        img = read_img(os.path.join(root, f"images/{vid:08d}.jpg"))
        intrinsics, extrinsics, dmin, dmax = read_cam_file(os.path.join(root, cam_folder, f"{vid:08d}_cam.txt"))
        img, intrinsics = resize_view(img, intrinsics, dataset, max_w, max_h)
        imgs.append(img)
        cams.append((intrinsics, extrinsics))
        if i == 0:
            rng = (dmin, dmax)
    sample = assemble_sample(imgs, cams, rng[0], rng[1], numdepth)
    prefix = "" if general else scan + "/"
    sample["filename"] = prefix + "{}/" + "{:0>8}".format(view_ids[0]) + "{}"
    return sample


def save_mask(filename: str, mask: np.ndarray) -> None:
    """Boolean mask -> 8-bit PNG with 0 / 255 (data_io.py:161-164)."""
    from PIL import Image
    if mask.dtype != np.bool_:
        raise ValueError("save_mask: boolean array expected")
    Image.fromarray(mask.astype(np.uint8) * 255).save(filename)


def save_outputs(outdir: str, filename: str, depth: np.ndarray, confs: Sequence[np.ndarray], cam: np.ndarray,
                 depth_max, depth_min, img: Optional[np.ndarray] = None) -> None:
    """Everything test.py:142-200 writes for one reference view, in the layout `filter.py` consumes:
    `<outdir>/<scan>/depth_est/<id>.pfm`, `cams/<id>_cam.txt`, `conf<i>/<id>.pfm` and - when `img` ([3,H,W] RGB in
    [0,1], the resized reference image the model saw) is given - `images/<id>.jpg` (test.py:160-162: clip(img*255) as
    uint8, RGB -> BGR, `cv2.imwrite`), which `filter.py:113,316` reads for the point colours."""
    def path(kind, ext):
        p = os.path.join(outdir, filename.format(kind, ext))
        os.makedirs(p.rsplit("/", 1)[0], exist_ok=True)
        return p
    save_pfm(path("depth_est", ".pfm"), np.ascontiguousarray(depth, dtype=np.float32))
    write_cam(path("cams", "_cam.txt"), cam, depth_max, depth_min)
    if img is not None:
        import cv2
        img = np.asarray(img)
        if img.ndim != 3 or img.shape[0] != 3:
            raise ValueError(f"save_outputs: img must be [3,H,W] (got {img.shape})")
        if tuple(img.shape[1:]) != tuple(np.asarray(depth).shape[-2:]):
            raise ValueError("save_outputs: image and depth map sizes differ (filter.py indexes one with the other's mask)")
        u8 = np.clip(np.transpose(img, (1, 2, 0)) * 255, 0, 255).astype(np.uint8)
        cv2.imwrite(path("images", ".jpg"), cv2.cvtColor(u8, cv2.COLOR_RGB2BGR))
    for i, c in enumerate(confs):
        save_pfm(path(f"conf{i}", ".pfm"), np.ascontiguousarray(c, dtype=np.float32))


# ------------------------------------------------------------------------------------------------
# fused point cloud (filter.py:214-227 writes it through `plyfile`, which is not installed here: the layout below is
# plyfile's default - binary little endian, one `vertex` element with float x, y, z and uchar red, green, blue - and
# is NOT pinned against plyfile itself)
# ------------------------------------------------------------------------------------------------
_PLY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def write_ply(filename: str, points: np.ndarray, colors: np.ndarray) -> None:
    """points [N,3] float32 (world), colors [N,3] uint8 -> binary PLY."""
    points, colors = np.asarray(points, dtype=np.float32), np.asarray(colors, dtype=np.uint8)
    if points.ndim != 2 or points.shape[1] != 3 or colors.shape != points.shape:
        raise ValueError("write_ply: points and colors must both be [N,3]")
    rec = np.empty(len(points), dtype=_PLY_DTYPE)
    rec["x"], rec["y"], rec["z"] = points[:, 0], points[:, 1], points[:, 2]
    rec["red"], rec["green"], rec["blue"] = colors[:, 0], colors[:, 1], colors[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex {}\nproperty float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").format(len(points))
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        rec.tofile(f)


def read_ply(filename: str) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of `write_ply` (only that layout)."""
    with open(filename, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
            if not line:
                raise ValueError(f"{filename}: malformed PLY header")
        rec = np.fromfile(f, dtype=_PLY_DTYPE, count=n)
    return np.stack((rec["x"], rec["y"], rec["z"]), 1), np.stack((rec["red"], rec["green"], rec["blue"]), 1)
