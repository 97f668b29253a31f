"""Golden vectors of the KITTI scan projection produced by the UNMODIFIED reference
(`gans.datasets.kitti.KITTIRaw.load_pts_as_img` + the `__getitem__` post-processing, reference
kitti.py:275-279,317-370) on a synthetic Velodyne-like scan: 66 rings (two more than the image
has rows: the reference's ring walk then assigns index -1, which numpy wraps), points before the
first ring start, depths outside [min_depth, max_depth].

    python tests/golden/make_golden_kitti.py      ->  tests/golden/kitti_scan.npz
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from small_cfgs import synthetic_scan  # noqa: E402


def main():
    from oracle import ref_import
    ref_import.install()
    import torch
    import torchvision.transforms.functional as TF
    from torchvision.transforms.functional import InterpolationMode
    from gans.datasets.kitti import KITTIRaw
    pts = synthetic_scan(rings=66, per_ring=120, seed=3)
    ds = object.__new__(KITTIRaw)
    ds.min_depth, ds.max_depth = 1.45, 80.0
    out = {"points": pts}
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "scan.bin")
        pts.tofile(f)
        for tag, unfold in (("unfold", True), ("elev", False)):
            img = ds.load_pts_as_img(f, unfold, 64, 2048)
            t = TF.to_tensor(img)
            t = TF.resize(t, (64, 512), InterpolationMode.NEAREST)
            t *= t[[5]]
            out[f"{tag}_image"] = t.numpy().astype(np.float32)
    path = os.path.join(HERE, "kitti_scan.npz")
    np.savez_compressed(path, **out)
    print(f"kitti_scan.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
