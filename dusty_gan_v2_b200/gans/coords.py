"""Drop-in for gans/coords.py: CoordBridge (reference 41-199), minus the visualisation
helpers (normal maps, bird's-eye rendering: out of the hot path).

* The angle grid is built once on the host with the same ATen CPU calls as the reference
  (coords.py:59-71), hence bit-identical.
* inv_depth_norm -> point_map / point_set on CUDA tensors is one kernel
  (dusty_point_project) fed by a host-computed sin/cos table of that grid; the validity
  mask count it returns is integer-exact.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as DF


class _CoordType:
    DEPTH = "depth"
    DEPTH_NORM = "depth_norm"
    INV_DEPTH = "inv_depth"
    INV_DEPTH_NORM = "inv_depth_norm"
    POINT_MAP = "point_map"
    POINT_SET = "point_set"
    NORMAL_MAP = "normal_map"

    def __init__(self):
        self.mode = (self.DEPTH, self.DEPTH_NORM, self.INV_DEPTH, self.INV_DEPTH_NORM,
                     self.POINT_MAP, self.POINT_SET, self.NORMAL_MAP)

    def __contains__(self, key):
        return key in self.mode


CoordType = _CoordType()
_POINTS = (CoordType.POINT_MAP, CoordType.POINT_SET)


class CoordBridge(nn.Module):
    def __init__(self, num_ring, num_points, min_depth, max_depth, angle_file, raydrop_const=0):
        super().__init__()
        self.min_depth, self.max_depth = float(min_depth), float(max_depth)
        assert self.max_depth > self.min_depth
        self.H, self.W = num_ring, num_points
        self.raydrop_const = raydrop_const
        raw = torch.from_numpy(np.load(angle_file)).permute(2, 0, 1)[None]       # [1,2,H0,W0]
        per = torch.cat([raw.sin(), raw.cos()], dim=1)
        per = torch.cat([per, per, per], dim=3)                                  # ring continuity
        per = F.interpolate(per, size=(self.H, self.W * 3), mode="bilinear", align_corners=False)
        per = per[..., self.W: 2 * self.W]
        angle = torch.atan2(per[:, :2], per[:, 2:])
        self.register_buffer("angle", angle)
        el, az = angle[0, 0].reshape(-1), angle[0, 1].reshape(-1)
        trig = torch.stack([el.cos(), el.sin(), az.cos(), az.sin()]).contiguous()
        self.register_buffer("_trig", trig, persistent=False)
        self.last_valid_count = None

    # -- masks (coords.py:73-86)
    def get_mask(self, x, coord):
        if coord == CoordType.DEPTH:
            return (x >= self.min_depth) & (x <= self.max_depth) & (x > 0.0)
        if coord == CoordType.INV_DEPTH:
            return (x >= (1 / self.max_depth)) & (x <= (1 / self.min_depth)) & (x > 0.0)
        if coord in (CoordType.DEPTH_NORM, CoordType.INV_DEPTH_NORM):
            return (x > 0.0) & (x <= 1.0)
        raise NotImplementedError(f"{coord}")

    def _points_from_inv_depth_norm(self, x, tol, as_set):
        if not x.is_cuda:
            raise RuntimeError("range->point projection runs on CUDA tensors only (no CPU fallback)")
        if tuple(x.shape[-2:]) != (self.H, self.W) or x.shape[1] != 1:
            raise RuntimeError(f"expected [B,1,{self.H},{self.W}], got {tuple(x.shape)}")
        pts, count = DF.point_project(x, self._trig, self.min_depth, self.max_depth, tol, as_set)
        self.last_valid_count = count       # int64 device tensor: number of valid pixels
        return pts

    def convert(self, x, src, tgt, tol=1e-11):
        assert src in CoordType, src
        assert tgt in CoordType, tgt
        T = CoordType
        if src == tgt:
            return x
        if tgt == T.NORMAL_MAP:
            raise NotImplementedError("normal maps are visualisation-only (gans/geometry.py)")
        if src == T.DEPTH:
            if tgt in (T.INV_DEPTH, T.INV_DEPTH_NORM):
                inv = 1 / x.add(tol) * self.get_mask(x, src).float()
                return inv * self.min_depth if tgt == T.INV_DEPTH_NORM else inv
            if tgt == T.DEPTH_NORM:
                return x / self.max_depth
            if tgt in _POINTS:
                pm = self.depth_to_point_map(x)
                return self.convert(pm, T.POINT_MAP, tgt)
        elif src == T.DEPTH_NORM:
            return self.convert(x * self.max_depth, T.DEPTH, tgt, tol)
        elif src == T.INV_DEPTH:
            if tgt == T.INV_DEPTH_NORM:
                return x * self.min_depth
            if tgt in (T.DEPTH, T.DEPTH_NORM):
                depth = 1 / x.add(tol) * self.get_mask(x, src).float()
                return depth / self.max_depth if tgt == T.DEPTH_NORM else depth
        elif src == T.INV_DEPTH_NORM:
            if tgt == T.INV_DEPTH:
                return x / self.min_depth
            if tgt in (T.DEPTH, T.DEPTH_NORM):
                return self.convert(x / self.min_depth, T.INV_DEPTH, tgt, tol)
            if tgt in _POINTS:
                return self._points_from_inv_depth_norm(x, tol, tgt == T.POINT_SET)
        elif src == T.POINT_MAP:
            if tgt == T.POINT_SET:
                return x.flatten(2).permute(0, 2, 1).contiguous()     # index h*W + w
            depth = torch.norm(x, p=2, dim=1, keepdim=True)
            return depth if tgt == T.DEPTH else self.convert(depth, T.DEPTH, tgt, tol)
        raise NotImplementedError(f"{src} to {tgt}")

    def depth_to_point_map(self, depth):
        assert depth.dim() == 4
        ce, se, ca, sa = (self._trig[i].reshape(1, 1, self.H, self.W) for i in range(4))
        return torch.cat((depth * ce * ca, depth * ce * sa, depth * se), dim=1)

    def extra_repr(self):
        return f'H={self.H}, W={self.W}, min_depth={self.min_depth}, max_depth="{self.max_depth}"'
