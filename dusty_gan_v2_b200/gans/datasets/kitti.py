"""Drop-in for the data side of gans/datasets/kitti.py (`KITTIRaw`, reference 222-370): a Velodyne
scan (`.bin`, [N, 4] float32) becomes the dict of [C, H, W] tensors the trainer's `fetch_reals`
consumes, with the image assembly on the GPU.

Split of work: the per-point image cell -- ring index by scan unfolding (a new ring starts where
the azimuth passes from the 4th to the 1st quadrant) or by elevation, column by azimuth -- is
integer arithmetic on float32 `arctan2` / `arcsin` results, so it is evaluated with numpy on the
host exactly as the reference evaluates it; the depth-ordered scatter (the reference's numba
loop, kitti.py:216-220), the NEAREST width reduction and the mask product (275-279) are one
device op, `dusty_scan_project`.  No CPU fallback for that part.
"""
from pathlib import Path

import numpy as np
import torch

from ... import _cabi as K

# KITTI odometry sequence -> (raw drive, first frame, last frame); 03 is not in KITTI raw
_ODOMETRY_TO_RAW = {
    0: ("2011_10_03_drive_0027_sync", 0, 4540), 1: ("2011_10_03_drive_0042_sync", 0, 1100),
    2: ("2011_10_03_drive_0034_sync", 0, 4660), 4: ("2011_09_30_drive_0016_sync", 0, 270),
    5: ("2011_09_30_drive_0018_sync", 0, 2760), 6: ("2011_09_30_drive_0020_sync", 0, 1100),
    7: ("2011_09_30_drive_0027_sync", 0, 1100), 8: ("2011_09_30_drive_0028_sync", 1100, 5170),
    9: ("2011_09_30_drive_0033_sync", 0, 1590), 10: ("2011_09_30_drive_0034_sync", 0, 1200),
}
_SPLITS = {"train": (0, 1, 2, 4, 5, 6, 7, 9, 10), "val": (8,)}


def scan_cells(points: np.ndarray, H: int = 64, W: int = 2048, scan_unfolding: bool = True):
    """(cell_h, cell_w, depth) of every point of a scan, as the reference computes them
    (kitti.py:319-363), without its Python loop over the rings: with D ring starts at indices
    d_0 < .. < d_{D-1}, a point after d_j and before d_{j+1} lies t = D - 1 - j rings below the
    top one and gets ring H - 1 - t; the reference's loop stops one step late, so t = H still
    gets ring -1 (numpy wraps it to H - 1), deeper ones and the points before d_0 keep ring 0."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 4)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    depth = np.linalg.norm(pts[:, :3], ord=2, axis=1)
    n = len(pts)
    if scan_unfolding:
        quad = np.where(x >= 0, np.where(y >= 0, 0, 3), np.where(y >= 0, 1, 2)).astype(np.int32)
        starts = np.flatnonzero(np.roll(quad, 1) - quad == 3)
        seg = np.searchsorted(starts, np.arange(n), side="right")        # ring starts at or before i
        t = len(starts) - seg
        cell_h = np.where((seg >= 1) & (t <= H), H - 1 - t, 0).astype(np.int32)
    else:
        up, down = np.deg2rad(3), np.deg2rad(-25)
        frac = 1 - (np.arcsin(z / depth) + abs(down)) / (up - down)
        cell_h = np.floor(frac * H).clip(0, H - 1).astype(np.int32)
    yaw = -np.arctan2(y, x)
    cell_w = np.floor((yaw / np.pi + 1) / 2 % 1 * W).clip(0, W - 1).astype(np.int32)
    return cell_h, cell_w, depth.astype(np.float32)


def scan_to_image(points, shape=(64, 512), min_depth=0.9, max_depth=120.0, scan_unfolding=True,
                  full_width=2048, device="cuda"):
    """[6, H, W] fp32 device tensor (x, y, z, reflectance, depth, mask), masked: the reference's
    `load_pts_as_img` -> to_tensor -> NEAREST resize -> `*= mask` (kitti.py:275-279,317-370)."""
    H, W_out = shape
    if full_width % W_out:
        raise ValueError("the output width must divide the projection width (NEAREST keeps every k-th column)")
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 4)
    cell_h, cell_w, depth = scan_cells(pts, H, full_width, scan_unfolding)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("scan_to_image assembles the image on a CUDA device (no CPU fallback)")
    p = torch.from_numpy(pts).to(dev)
    d = torch.from_numpy(depth).to(dev)
    ch, cw = torch.from_numpy(cell_h).to(dev), torch.from_numpy(cell_w).to(dev)
    keys = torch.empty(H * full_width, dtype=torch.int64, device=dev)
    out = torch.empty((6, H, W_out), dtype=torch.float32, device=dev)
    K.call("dusty_scan_project", K.ptr(p), K.ptr(d), K.ptr(ch), K.ptr(cw), K.ptr(keys), K.ptr(out), len(pts), H,
           full_width, W_out, float(min_depth), float(max_depth), K.stream_of(p))
    return out


class KITTIRaw(torch.utils.data.Dataset):
    """Same constructor arguments and item layout as the reference's `KITTIRaw`; items are device
    tensors.  `files`: explicit list of `.bin` paths (the reference's "test" split enumerates
    drive lists that are dataset metadata, not part of this path: pass them here)."""

    mean = {"xyz": [-0.01506443, 0.45959818, -0.89225304], "reflectance": 0.24130844, "depth": 9.689281}
    std = {"xyz": [11.224804, 8.237693, 0.88183135], "reflectance": 0.16860831, "depth": 10.08752}

    def __init__(self, root="data/kitti_raw", split="train", shape=(64, 2048), min_depth=0.9, max_depth=120.0,
                 flip=False, scan_unfolding=True, files=None, device="cuda"):
        super().__init__()
        self.root, self.split, self.shape = Path(root), split, tuple(shape)
        self.min_depth, self.max_depth, self.flip, self.scan_unfolding = min_depth, max_depth, flip, scan_unfolding
        self.device = device
        if files is not None:
            self.datalist = [str(f) for f in files]
        elif split in _SPLITS:
            self.datalist = []
            for seq in _SPLITS[split]:
                drive, first, last = _ODOMETRY_TO_RAW[seq]
                base = f"{self.root}/{drive[:10]}/{drive}/velodyne_points/data"
                self.datalist += [f"{base}/{i:010d}.bin" for i in range(first, last + 1)]
        else:
            raise ValueError(f"split {split!r}: pass the scan files of other splits through `files=`")

    def __len__(self):
        return len(self.datalist)

    def __getitem__(self, index):
        pts = np.fromfile(self.datalist[index], dtype=np.float32).reshape(-1, 4)
        img = scan_to_image(pts, self.shape, self.min_depth, self.max_depth, self.scan_unfolding, device=self.device)
        if self.flip and np.random.rand() > 0.5:
            img = img.flip(-1)
        return {"xyz": img[:3], "reflectance": img[3:4], "depth": img[4:5], "mask": img[5:6]}

    def normalize(self, item):
        out = dict(item)
        for key in ("xyz", "reflectance", "depth"):
            if key in out:
                m = torch.as_tensor(self.mean[key], device=out[key].device).reshape(-1, 1, 1)
                s = torch.as_tensor(self.std[key], device=out[key].device).reshape(-1, 1, 1)
                out[key] = (out[key] - m) / s
        return out
