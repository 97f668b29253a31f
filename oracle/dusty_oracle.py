"""CPU oracle for the DUSty-v2 G/D hot path.

TEST INFRASTRUCTURE ONLY -- this file is the *checker*, never the product.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The shipped package (dusty_gan_v2_b200) never does; its ops
raise when the CUDA library is missing.

What it is: a plain fp32 PyTorch-on-CPU *restatement* (functional, state_dict
driven, closed-form FIR / modulation algebra) of what the reference computes on
CPU tensors, i.e. of the reference's own "ref" path.  Every
function cites the reference file:line it follows.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4/8c).  The oracle is therefore pinned against outputs of the reference
itself, imported in the build container from /root/reference by
tests/golden/make_golden.py; the resulting small fixtures are committed under
tests/golden/ and checked by tests/test_oracle_golden.py on every CPU run.  Where
/root/reference is present (the build container) tests/test_reference_live.py also
cross-checks it against the live reference on fresh random inputs, up to the
reference's real Trainer.step against `train_iteration`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SQRT2 = 2.0 ** 0.5

# --------------------------------------------------------------------------------------
# a3  fused bias + leaky-ReLU                      gans/models/ops/fused_act/fused_act.py
# --------------------------------------------------------------------------------------


def bias_act(x: Tensor, bias: Optional[Tensor] = None, slope: float = 0.2,
             scale: float = SQRT2) -> Tensor:
    """y = lrelu(x + b_c) * scale, bias broadcast over dim 1.

    fused_act.py:112-124 (CPU branch) and fused_bias_act_kernel.cu:27-60
    (act=3, grad=0).  NOTE the reference CPU branch hard-codes slope 0.2; the CUDA
    kernel honours `alpha`.  Every caller passes 0.2 so the two agree.
    """
    if bias is not None:
        shape = [1, -1] + [1] * (x.ndim - 2)
        x = x + bias.reshape(shape)
    return torch.where(x > 0, x, x * slope) * scale


def bias_act_grad(dy: Tensor, out: Tensor, slope: float = 0.2,
                  scale: float = SQRT2) -> Tuple[Tensor, Tensor]:
    """First-order backward, gated by the saved *output* (fused_act.py:20-44,
    kernel case 31): dx = dy * (out > 0 ? 1 : slope) * scale, db = sum over all
    dims but 1."""
    dx = torch.where(out > 0, dy, dy * slope) * scale
    dims = [0] + list(range(2, dx.ndim))
    return dx, dx.sum(dims)


# --------------------------------------------------------------------------------------
# a5  upfirdn2d                                    gans/models/ops/upfirdn2d/upfirdn2d.py
# --------------------------------------------------------------------------------------


def upfirdn2d_out_size(n_in: int, up: int, down: int, p0: int, p1: int, k: int) -> int:
    """upfirdn2d.py:93-94 / upfirdn2d_kernel.cu:232-235."""
    return (n_in * up + p0 + p1 - k + down) // down


def upfirdn2d(x: Tensor, kernel: Tensor, up=1, down=1, pad=(0, 0)) -> Tensor:
    """pad -> zero-insert upsample -> true convolution with `kernel` -> decimate.

    Restates upfirdn2d_native (upfirdn2d.py:167-208) as an explicit polyphase
    gather: out[my,mx] = sum_{ty,tx} K[kh-1-ty, kw-1-tx] * xu[my*dy+ty-py0, mx*dx+tx-px0]
    where xu is the zero-inserted input (zero outside).  Argument order follows
    the reference wrapper (upfirdn2d.py:148-164): up/down are (x, y), pad is
    (x0, x1[, y0, y1]).
    """
    up_x, up_y = (up, up) if isinstance(up, int) else up
    down_x, down_y = (down, down) if isinstance(down, int) else down
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    px0, px1, py0, py1 = pad
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    oh = upfirdn2d_out_size(h, up_y, down_y, py0, py1, kh)
    ow = upfirdn2d_out_size(w, up_x, down_x, px0, px1, kw)
    kflip = torch.flip(kernel, [0, 1]).to(x.dtype)
    # zero-inserted signal on its own lattice
    xu = x.new_zeros(n, c, h * up_y, w * up_x)
    xu[:, :, ::up_y, ::up_x] = x
    # explicit pad (positive) / crop (negative)
    xu = F.pad(xu, (max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)))
    xu = xu[:, :, max(-py0, 0): xu.shape[2] - max(-py1, 0),
            max(-px0, 0): xu.shape[3] - max(-px1, 0)]
    full = F.conv2d(xu.reshape(n * c, 1, *xu.shape[2:]), kflip[None, None])
    full = full[:, :, ::down_y, ::down_x]
    assert full.shape[2] == oh and full.shape[3] == ow, (full.shape, oh, ow)
    return full.reshape(n, c, oh, ow)


# --------------------------------------------------------------------------------------
# a4  Resample / BlurVH / filter2d / Pad                   gans/models/ops/common.py
# --------------------------------------------------------------------------------------


def resample_geometry(n_taps: int, up: int, down: int) -> Tuple[int, int]:
    """(p0, p1) per axis as derived in Resample.__init__ (common.py:89-101)."""
    if up > 1:
        return (n_taps - up + 1) // 2 + up - 1, (n_taps - up) // 2
    return (n_taps - down + 1) // 2, (n_taps - down) // 2


def _boundary_index(idx: Tensor, n: int, circular: bool) -> Tensor:
    return torch.remainder(idx, n) if circular else idx.clamp(0, n - 1)


# Two evaluations of the same separable pass.  "gather" (default: what every parity test and
# smoke() check against) is the literal tap-by-tap form; "library" runs it as a depthwise
# library correlation over the extended, zero-stuffed signal -- the way the reference's CPU path
# spends its time -- and is what bench.py's CPU baseline legs select, so that the port is a fair
# stand-in for the reference's speed (tests/golden/calibrate_cpu_port.py).  The two agree to
# 1-2 ulp (tests/test_oracle_golden.py::test_fir_axis_forms_agree runs the goldens under both).
_FIR_IMPL = {"mode": "gather"}


def set_fir_impl(mode: str) -> str:
    """Select the FIR evaluation ("gather" | "library"); returns the previous mode."""
    if mode not in ("gather", "library"):
        raise ValueError(mode)
    prev, _FIR_IMPL["mode"] = _FIR_IMPL["mode"], mode
    return prev


class fir_impl:
    """`with fir_impl("library"): ...` -- scoped `set_fir_impl`."""

    def __init__(self, mode: str):
        self.mode = mode

    def __enter__(self):
        self.prev = set_fir_impl(self.mode)
        return self

    def __exit__(self, *exc):
        set_fir_impl(self.prev)
        return False


def _fir_axis(x: Tensor, taps: Tensor, up: int, down: int, p0: int, p1: int,
              axis: int, circular: bool) -> Tensor:
    """One separable pass: out[m] = sum_t taps[t] * xu[m*down + t - p0], where
    xu[q] = X(q/up) if up | q else 0 and X() extends x circularly (ring, W axis)
    or by edge replication (H axis) -- the closed form of margin-pad -> zero-insert ->
    crop -> depthwise correlate -> stride (common.py:105-135).

    Dispatches on `set_fir_impl`: the literal form (`_fir_axis_gather`) or the depthwise
    library correlation below."""
    if x.ndim != 4 or _FIR_IMPL["mode"] != "library":
        return _fir_axis_gather(x, taps, up, down, p0, p1, axis, circular)
    n = x.shape[axis]
    k = taps.numel()
    n_out = (n * up + p0 + p1 - k) // down + 1
    # source samples q = -p0 .. (n_out-1)*down + k-1 - p0 of the zero-stuffed, extended signal
    q_lo, q_hi = -p0, (n_out - 1) * down + k - 1 - p0
    i_lo, i_hi = -((-q_lo) // up) if q_lo < 0 else (q_lo + up - 1) // up, q_hi // up   # ceil / floor
    idx = _boundary_index(torch.arange(i_lo, i_hi + 1), n, circular)
    ext = x.index_select(axis, idx) if (i_lo < 0 or i_hi >= n) else x.narrow(axis, i_lo, i_hi - i_lo + 1)
    if up > 1:
        shape = list(ext.shape)
        shape[axis] = q_hi - q_lo + 1
        stuffed = x.new_zeros(shape)
        first = i_lo * up - q_lo                       # position of sample i_lo inside [q_lo, q_hi]
        stuffed.narrow(axis, first, (i_hi - i_lo) * up + 1).index_copy_(
            axis, torch.arange(0, (i_hi - i_lo) * up + 1, up), ext)
        ext = stuffed
    C = x.shape[1]
    w = taps.to(x.dtype).reshape((1, 1, k, 1) if axis == 2 else (1, 1, 1, k)).repeat(C, 1, 1, 1)
    stride = (down, 1) if axis == 2 else (1, down)
    return F.conv2d(ext, w, stride=stride, groups=C)


def _fir_axis_gather(x: Tensor, taps: Tensor, up: int, down: int, p0: int, p1: int,
                     axis: int, circular: bool) -> Tensor:
    """The literal form of `_fir_axis`: one gathered, weighted term per tap."""
    n = x.shape[axis]
    k = taps.numel()
    n_out = (n * up + p0 + p1 - k) // down + 1
    m = torch.arange(n_out)
    out = None
    for t in range(k):
        q = m * down + t - p0
        hit = (torch.remainder(q, up) == 0)
        src = _boundary_index(torch.div(q, up, rounding_mode="floor"), n, circular)
        term = x.index_select(axis, src)
        shape = [1] * x.ndim
        shape[axis] = n_out
        term = term * (hit.to(x.dtype) * taps[t]).reshape(shape)
        out = term if out is None else out + term
    return out


def resample(x: Tensor, up: int = 1, down: int = 1, window: Sequence[float] = (1, 3, 3, 1),
             ring: bool = True, normalize: bool = True, direction: str = "hw") -> Tensor:
    """Resample.forward (common.py:45-138): W pass first, then H pass
    (common.py:126-128)."""
    taps = torch.tensor(list(window), dtype=torch.float32)
    if normalize:
        taps = taps / taps.sum()
    up_h = up if "h" in direction else 1
    up_w = up if "w" in direction else 1
    down_h = down if "h" in direction else 1
    down_w = down if "w" in direction else 1
    # kernel *= (up_h*up_w) ** (ndim/2), ndim == 1  (common.py:82)
    taps = taps * float(up_h * up_w) ** 0.5
    taps = taps.to(x.dtype)
    y = x
    if "w" in direction:
        p0, p1 = resample_geometry(len(window), up_w, down_w)
        y = _fir_axis(y, taps, up_w, down_w, p0, p1, axis=3, circular=ring)
    if "h" in direction:
        p0, p1 = resample_geometry(len(window), up_h, down_h)
        y = _fir_axis(y, taps, up_h, down_h, p0, p1, axis=2, circular=False)
    return y


def blur_vh(x: Tensor, window: Sequence[float] = (1, 2, 1), ring: bool = True) -> Tensor:
    """BlurVH (common.py:141-155): cat(vertical blur, horizontal blur) on dim 1."""
    v = resample(x, window=window, ring=ring, direction="h")
    h = resample(x, window=window, ring=ring, direction="w")
    return torch.cat([v, h], dim=1)


def pad2d(x: Tensor, padding, ring: bool = False, mode: str = "replicate") -> Tensor:
    """Pad (common.py:10-24): W padded first (circular when ring), then H (`mode`)."""
    if isinstance(padding, int):
        padding = (padding,) * 4
    left, right, top, bottom = padding
    x = F.pad(x, (left, right, 0, 0), mode="circular" if ring else mode)
    return F.pad(x, (0, 0, top, bottom), mode=mode)


def filter2d(x: Tensor, kernel: Tensor, gain: float = 1.0) -> Tensor:
    """filter2d (common.py:27-42): same-size separable blur, circular W / replicate H."""
    k = kernel / kernel.sum()
    k = k * (gain ** 0.5)
    n = k.numel()
    y = _fir_axis(x, k.to(x.dtype), 1, 1, n // 2, (n - 1) // 2, axis=3, circular=True)
    return _fir_axis(y, k.to(x.dtype), 1, 1, n // 2, (n - 1) // 2, axis=2, circular=False)


# --------------------------------------------------------------------------------------
# a8  EqualLR / PixelNorm / MappingNetwork
# --------------------------------------------------------------------------------------


def equal_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], gain: float = 1.0,
                 lr_mul: float = 1.0) -> Tensor:
    """EqualLR(nn.Linear) (common.py:158-184): scales the *input* by 1/sqrt(fan_in),
    applies the module (bias included), then multiplies by gain*lr_mul."""
    scale = 1.0 / math.sqrt(weight[0].numel())
    return F.linear(x * scale, weight, bias) * (gain * lr_mul)


def equal_conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1,
                 gain: float = 1.0) -> Tensor:
    scale = 1.0 / math.sqrt(weight[0].numel())
    return F.conv2d(x * scale, weight, bias, stride=stride) * gain


def pixel_norm(x: Tensor, alpha: float = 1e-8) -> Tensor:
    """common.py:213-223."""
    return x / x.pow(2.0).mean(dim=1, keepdim=True).add(alpha).sqrt()


def mapping_network(sd: Dict[str, Tensor], z: Tensor, prefix: str = "mapping_network.",
                    depth: int = 2) -> Tensor:
    """MappingNetwork (dusty_v2.py:13-29): PixelNorm, then `depth` x
    [EqualLR(Linear, gain sqrt2, lr_mul 0.01) + LeakyReLU(0.2)]."""
    h = pixel_norm(z)
    for i in range(1, depth + 1):
        h = equal_linear(h, sd[f"{prefix}{i}.0.module.weight"], sd[f"{prefix}{i}.0.module.bias"],
                         gain=SQRT2, lr_mul=0.01)
        h = F.leaky_relu(h, 0.2)
    return h


# --------------------------------------------------------------------------------------
# a2  FourierFeature                                   gans/models/ops/fourier.py:77-82
# --------------------------------------------------------------------------------------


def fourier_feature(angle: Tensor, freqs: Tensor, phase: Tensor) -> Tensor:
    """coords = f_h*elev + f_w*azim + phase (a 1x1 conv), out = cat(sin, cos)."""
    f = freqs.reshape(-1, 2).to(angle.dtype)
    coords = (angle[:, 0:1] * f[:, 0].reshape(1, -1, 1, 1)
              + angle[:, 1:2] * f[:, 1].reshape(1, -1, 1, 1)
              + phase.reshape(1, -1, 1, 1))
    return torch.cat([coords.sin(), coords.cos()], dim=1)


# --------------------------------------------------------------------------------------
# a1  ModConv2d (1x1)                                   gans/models/ops/style.py:68-126
# --------------------------------------------------------------------------------------


def modconv_weights(style_w: Tensor, weight: Tensor, mod_w: Tensor, mod_b: Tensor,
                    ema_var: Tensor, demod: bool) -> Tensor:
    """Per-sample effective 1x1 weights Wb[B,O,I] (never built as [B,O,I,1,1]):

      s   = EqualLR-linear(style)                         style.py:72
      W'  = scale*W / max|scale*W|   (demod only)         style.py:73-78
      s'  = s / max_i|s_i| + 1 (demod)  or  s + 1         style.py:79-83
      Wb  = W' * s'                                       style.py:93
      Wb *= rsqrt(sum_i Wb^2 + 1e-8)  (demod)             style.py:96-98
      Wb /= sqrt(ema_var) + 1e-8                          style.py:103
    """
    o, i = weight.shape[1], weight.shape[2]
    w = weight.reshape(o, i) * (1.0 / math.sqrt(i))
    s = equal_linear(style_w, mod_w, mod_b)
    if demod:
        w = w / w.abs().max()
        s = s / s.abs().amax(dim=1, keepdim=True)
    s = s + 1.0
    wb = w[None] * s[:, None, :]
    if demod:
        wb = wb * torch.rsqrt(wb.pow(2).sum(dim=2, keepdim=True) + 1e-8)
    return wb / (torch.sqrt(ema_var) + 1e-8)


def modconv(x: Tensor, style_w: Tensor, weight: Tensor, mod_w: Tensor, mod_b: Tensor,
            ema_var: Tensor, demod: bool = True, bias: Optional[Tensor] = None,
            training: bool = False, ema_decay: float = 0.9989) -> Tuple[Tensor, Tensor]:
    """Modulated 1x1 conv as a per-sample contraction Y_b = Wb_b @ X_b.
    In training the EMA of mean(x^2) is updated *before* it is used
    (style.py:99-103).  Returns (y, new_ema_var)."""
    b, i, h, w_ = x.shape
    if training:
        var = x.detach().pow(2).mean()
        ema_var = torch.lerp(ema_var, var, 1 - ema_decay)
    wb = modconv_weights(style_w, weight, mod_w, mod_b, ema_var, demod)
    y = torch.bmm(wb, x.reshape(b, i, h * w_)).reshape(b, -1, h, w_)
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y, ema_var


# --------------------------------------------------------------------------------------
# a10  Gumbel-sigmoid raydrop               gans/models/ops/gumbel.py, gans/models/dusty_v1.py
# --------------------------------------------------------------------------------------

_EPS32 = float(torch.finfo(torch.float32).eps)
_TINY32 = float(torch.finfo(torch.float32).tiny)


def gumbel_sigmoid(logits: Tensor, u: Tensor, temperature: float = 1.0) -> Tuple[Tensor, Tensor]:
    """Closed form of RelaxedBernoulli(T, logits).rsample() given the uniform draw
    `u` (torch.distributions relaxed_bernoulli.py rsample + SigmoidTransform with
    clamped probs), then the straight-through hard threshold (gumbel.py:23-29).
    Returns (mask_with_soft_gradient, soft)."""
    probs = torch.sigmoid(logits).clamp(_EPS32, 1 - _EPS32)
    uu = u.clamp(_EPS32, 1 - _EPS32)
    y = (uu.log() - (-uu).log1p() + probs.log() - (-probs).log1p()) / temperature
    soft = torch.sigmoid(y).clamp(_TINY32, 1 - _EPS32)
    hard = (soft > 0.5).to(logits.dtype)
    return (hard - soft).detach() + soft, soft


def raydrop(image: Tensor, logits: Tensor, u: Tensor, raydrop_const: float = -1.0,
            temperature: float = 1.0) -> Dict[str, Tensor]:
    """RayDropModel.forward (dusty_v1.py:20-25)."""
    mask, _ = gumbel_sigmoid(logits, u, temperature)
    const = torch.tensor(float(raydrop_const), dtype=image.dtype)
    return {"raydrop_mask": mask, "image_orig": image,
            "image": torch.lerp(image, const, 1 - mask)}


# --------------------------------------------------------------------------------------
# a14  CoordBridge                                                   gans/coords.py
# --------------------------------------------------------------------------------------


def angle_grid(angle_hw2: np.ndarray, num_ring: int, num_points: int) -> Tensor:
    """CoordBridge.__init__ (coords.py:59-71): sin/cos -> tile x3 along W -> bilinear
    resize -> centre crop -> atan2.  Bit-exact restatement (same torch CPU ops)."""
    a = torch.from_numpy(np.ascontiguousarray(angle_hw2)).permute(2, 0, 1)[None]
    per = torch.cat([a.sin(), a.cos()], dim=1)
    per = torch.cat([per, per, per], dim=3)
    per = F.interpolate(per, size=(num_ring, num_points * 3), mode="bilinear",
                        align_corners=False)
    per = per[..., num_points: 2 * num_points]
    return torch.atan2(per[:, :2], per[:, 2:])


def inv_depth_norm_to_depth(x: Tensor, min_depth: float, max_depth: float,
                            tol: float = 1e-11) -> Tuple[Tensor, Tensor]:
    """coords.py:143-147 + 73-81: valid = (x>tol) & (1/max <= x/min <= 1/min) & (x/min>0);
    depth = valid / (x/min + tol).  Returns (depth, valid)."""
    inv = x / min_depth
    valid = (x > tol).float()
    valid = valid * ((inv >= 1 / max_depth) & (inv <= 1 / min_depth) & (inv > 0.0)).float()
    depth = 1 / (inv + tol) * valid
    return depth, valid


def depth_to_point_map(depth: Tensor, angle: Tensor) -> Tensor:
    """coords.py:178-185: xyz = depth * [cos el cos az, cos el sin az, sin el]."""
    c, s = torch.cos(angle), torch.sin(angle)
    return torch.cat([depth * c[:, [0]] * c[:, [1]], depth * c[:, [0]] * s[:, [1]],
                      depth * s[:, [0]]], dim=1)


def inv_depth_norm_to_points(x: Tensor, angle: Tensor, min_depth: float, max_depth: float):
    """convert(x, 'inv_depth_norm', 'point_map'/'point_set') (coords.py:139-155).
    Returns (point_map[B,3,H,W], point_set[B,HW,3], valid_count int)."""
    depth, valid = inv_depth_norm_to_depth(x, min_depth, max_depth)
    pm = depth_to_point_map(depth, angle)
    ps = pm.flatten(2).permute(0, 2, 1).contiguous()
    return pm, ps, int(valid.sum().item())


def depth_to_inv_depth_norm(depth: Tensor, min_depth: float, max_depth: float,
                            tol: float = 1e-11) -> Tensor:
    """convert(depth,'depth','inv_depth_norm') (coords.py:92-98,130-132)."""
    valid = ((depth >= min_depth) & (depth <= max_depth) & (depth > 0.0)).float()
    return (1 / (depth + tol) * valid) * min_depth


def fetch_reals(depth: Tensor, mask: Tensor, min_depth: float, max_depth: float,
                raydrop_const: float = -1.0) -> Tensor:
    """Trainer.fetch_reals (trainer.py:211-217)."""
    x = depth_to_inv_depth_norm(depth, min_depth, max_depth) * 2.0 - 1.0
    return mask * x + (1 - mask) * raydrop_const


# --------------------------------------------------------------------------------------
# a12  MinibatchStdDev                                    gans/models/ops/common.py:226-253
# --------------------------------------------------------------------------------------


def minibatch_stddev(x: Tensor, group: int = 4, alpha: float = 1e-8) -> Tensor:
    """features == 1.  Groups are *strided*: members {m, m+B/G, m+2B/G, ...}."""
    b, c, h, w = x.shape
    g = min(b, group)
    y = x.reshape(g, b // g, c, h, w)
    y = (y - y.mean(0, keepdim=True)).pow(2).mean(0)
    y = torch.sqrt(y + alpha).mean(dim=(1, 2, 3))          # [B/G]
    y = y.reshape(1, b // g, 1, 1, 1).expand(g, b // g, 1, h, w).reshape(b, 1, h, w)
    return torch.cat([x, y], dim=1)


# --------------------------------------------------------------------------------------
# a7  aug-coords unshift                                 gans/models/dusty_v2.py:290-297
# --------------------------------------------------------------------------------------


def circular_unshift(v: Tensor, shift_rad: Tensor) -> Tensor:
    """cat([v,v],W) -> affine_grid(translation shift/2pi on x) -> bilinear grid_sample
    -> first W columns.  Same ATen calls as the reference so the fp32 coordinate
    arithmetic is reproduced exactly."""
    b, _, _, w = v.shape
    t = shift_rad / (2 * np.pi)
    theta = torch.zeros(b, 2, 3, dtype=v.dtype)
    theta[:, 0, 0] = 1
    theta[:, 1, 1] = 1
    theta[:, 0, 2] = t
    vv = torch.cat([v, v], dim=3)
    grid = F.affine_grid(theta, vv.shape, align_corners=False)
    return F.grid_sample(vv, grid, mode="bilinear", align_corners=False)[..., :w]


# --------------------------------------------------------------------------------------
# a7/a9  dusty_v2 generator                     gans/models/dusty_v2.py, gans/models/base.py
# --------------------------------------------------------------------------------------


def downsample_angle(angle: Tensor) -> Tensor:
    """SynthesisBlock.downsample_angle (dusty_v2.py:135-140)."""
    c = angle.shape[1]
    per = resample(torch.cat([angle.sin(), angle.cos()], dim=1), down=2)
    return torch.atan2(per[:, :c], per[:, c:])


def _num_levels(sd: Dict[str, Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}layers.{n}.conv1.weight" in sd:
        n += 1
    return n


def synthesis_network(sd: Dict[str, Tensor], ws: Tensor, angle: Tensor, training: bool = False,
                      shifts_rad: Optional[Tensor] = None, output_scale: float = 0.25,
                      prefix: str = "synthesis_network.",
                      new_buffers: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """SynthesisNetwork.forward (dusty_v2.py:261-308) with the per-block body of
    SynthesisBlock.forward (dusty_v2.py:142-180).  `shifts_rad` [B] is the
    aug-coords azimuth shift (already in radians) that the reference draws with
    uniform_ (dusty_v2.py:266-274); None disables it (eval mode)."""
    n_lv = _num_levels(sd, prefix)
    if shifts_rad is not None:
        sh = torch.zeros(ws.shape[0], 2, dtype=angle.dtype)
        sh[:, 1] = shifts_rad
        angle = angle + sh[..., None, None]
    pyramid = [angle]
    for _ in range(n_lv - 1):
        pyramid.insert(0, downsample_angle(pyramid[0]))

    def mc(x, style, name, demod, with_bias):
        p = f"{prefix}{name}."
        y, ev = modconv(x, style, sd[p + "weight"], sd[p + "mod.module.weight"],
                        sd[p + "mod.module.bias"], sd[p + "ema_var"], demod=demod,
                        bias=sd[p + "bias"] if with_bias else None, training=training)
        if new_buffers is not None:
            new_buffers[p + "ema_var"] = ev
        return y

    h, skip, i = None, None, 0
    for l, ang in enumerate(pyramid):
        lp = f"layers.{l}."
        pe = fourier_feature(ang, sd[f"{prefix}{lp}pe.freqs"], sd[f"{prefix}{lp}pe.phase"])
        if h is None:
            h = pe
        else:
            h = torch.cat([resample(h, up=2), pe], dim=1)
        h = bias_act(mc(h, ws[:, i], lp + "conv1", True, False), sd[f"{prefix}{lp}bias_act1.bias"])
        n_conv = 1
        if l > 0:
            h = bias_act(mc(h, ws[:, i + 1], lp + "conv2", True, False),
                         sd[f"{prefix}{lp}bias_act2.bias"])
            n_conv = 2
        out = {}
        for name in ("image", "raydrop_logit"):
            o = mc(h, ws[:, i + n_conv], f"{lp}head.heads.{name}", False, True)
            if skip is not None:
                o = o + resample(skip[name], up=2)
            out[name] = o
        skip = out
        i += n_conv

    if shifts_rad is not None:
        skip = {k: circular_unshift(v, shifts_rad) for k, v in skip.items()}
    skip = {k: v * output_scale for k, v in skip.items()}
    skip["image"] = torch.tanh(skip["image"])
    return skip


def generator(sd: Dict[str, Tensor], z: Tensor, angle: Tensor, u: Tensor, training: bool = False,
              shifts_rad: Optional[Tensor] = None, truncation_psi: float = 1.0,
              input_w: bool = False, raydrop_const: float = -1.0, temperature: float = 1.0,
              new_buffers: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """base.Generator.forward (base.py:26-63) for dusty_v2.  `u` is the uniform
    tensor RelaxedBernoulli would draw (shape of raydrop_logit)."""
    n_styles = 2 * _num_levels(sd, "synthesis_network.")
    if input_w:
        w = z
    else:
        w = mapping_network(sd, z)
        w = torch.stack([w] * n_styles, dim=1)
    if training:
        if new_buffers is not None:
            new_buffers["w_avg"] = torch.lerp(sd["w_avg"], w[:, 0].mean(0, keepdim=True).detach(),
                                              1 - 0.995)
    elif truncation_psi != 1.0:
        w = torch.lerp(sd["w_avg"][None].expand_as(w), w, truncation_psi)
    o = synthesis_network(sd, w, angle, training, shifts_rad, new_buffers=new_buffers)
    o["w"] = w
    o.update(raydrop(o["image"], o["raydrop_logit"], u, raydrop_const, temperature))
    return o


# --------------------------------------------------------------------------------------
# a11  dusty_v2 discriminator                          gans/models/dusty_v2.py:325-396
# --------------------------------------------------------------------------------------


def _conv_ring(x: Tensor, w: Tensor, stride: int, padding: int) -> Tensor:
    """ops.Conv2d(bias=False, ring=True, equal_lr=True) (common.py:187-210)."""
    if padding:
        x = pad2d(x, padding, ring=True)
    return equal_conv2d(x, w, None, stride)


def discriminator(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    h = blur_vh(x)
    h = bias_act(_conv_ring(h, sd["layers.1.0.module.weight"], 1, 0), sd["layers.2.bias"])
    l = 3
    while f"layers.{l}.conv1.1.module.weight" in sd:
        p = f"layers.{l}."
        r = bias_act(_conv_ring(h, sd[p + "conv1.1.module.weight"], 1, 1), sd[p + "bias_act1.bias"])
        r = bias_act(_conv_ring(resample(r), sd[p + "conv2.1.module.weight"], 2, 1),
                     sd[p + "bias_act2.bias"])
        s = _conv_ring(resample(h), sd[p + "skip.0.module.weight"], 2, 0)
        h = (r + s) / SQRT2
        l += 1
    h = minibatch_stddev(h)
    h = bias_act(_conv_ring(h, sd["epilogue.1.1.module.weight"], 1, 1), sd["epilogue.2.bias"])
    h = h.flatten(1)
    h = bias_act(equal_linear(h, sd["epilogue.4.module.weight"], None), sd["epilogue.5.bias"])
    return equal_linear(h, sd["epilogue.6.module.weight"], sd["epilogue.6.module.bias"])


# --------------------------------------------------------------------------------------
# a13  losses                                   gans/models/loss.py, gans/trainer.py:426-447
# --------------------------------------------------------------------------------------


def nsgan_g(y_fake: Tensor) -> Tensor:
    return F.softplus(-y_fake).mean()


def nsgan_d(y_real: Tensor, y_fake: Tensor) -> Tensor:
    return F.softplus(-y_real).mean() + F.softplus(y_fake).mean()


def r1_penalty(grads: Tensor) -> Tensor:
    return grads.pow(2).sum(dim=[1, 2, 3]).mean()


# --------------------------------------------------------------------------------------
# a6  AdaptiveAugment.forward, deterministic part       gans/augment/adaptive_augment.py:471-545
# --------------------------------------------------------------------------------------

SYM6 = (0.015404109327027373, 0.0034907120842174702, -0.11799011114819057,
        -0.048311742585633, 0.4910559419267466, 0.787641141030194, 0.3379294217276218,
        -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148)


def _m3(rows):
    return torch.tensor(rows, dtype=torch.float32)


def ada_padding(G_inv: Tensor, height: int, width: int, ksize: int):
    """get_padding (adaptive_augment.py:271-291) -> (x1, x2, y1, y2) python ints."""
    cx, cy = (width - 1) / 2, (height - 1) / 2
    cp = G_inv @ _m3([(-cx, -cy, 1), (cx, -cy, 1), (cx, cy, 1), (-cx, cy, 1)]).T
    pad_k = ksize // 4
    pad = cp[:, :2, :].permute(1, 0, 2).flatten(1)
    pad = torch.cat((-pad, pad)).max(1).values
    pad = pad + _m3([pad_k * 2 - cx, pad_k * 2 - cy] * 2)
    pad = pad.max(_m3([0, 0] * 2)).min(_m3([width - 1, height - 1] * 2))
    x1, y1, x2, y2 = (int(v) for v in pad.ceil().to(torch.int32))
    return x1, x2, y1, y2


def bilinear_warp(img: Tensor, grid: Tensor) -> Tensor:
    """F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=False) written
    with gather so that it is differentiable to any order (ATen's grid_sampler backward
    is not; the reference wraps it in GridSampleForward/Backward,
    adaptive_augment.py:49-96, for the same reason)."""
    B, C, H, W = img.shape
    ix = ((grid[..., 0] + 1) * W - 1) / 2
    iy = ((grid[..., 1] + 1) * H - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    fx, fy = ix - x0, iy - y0
    flat = img.reshape(B, C, H * W)
    out = 0
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi, yi = x0 + dx, y0 + dy
            ok = ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).to(img.dtype)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long().reshape(B, 1, -1)
            val = torch.gather(flat, 2, idx.expand(B, C, -1)).reshape(B, C, *grid.shape[1:3])
            out = out + val * (wy * wx * ok).unsqueeze(1)
    return out


def ada_apply(img: Tensor, G_inv: Tensor, C: Tensor) -> Tensor:
    """The deterministic body of AdaptiveAugment.forward for a given inverse geometric
    transform G_inv [B,3,3] and colour matrix C [B,4,4]: circular-W / reflect-H pad,
    SYM6 2x upsampling (W then H), bilinear affine warp, SYM6 2x downsampling, colour."""
    B, ch, H, W = img.shape
    k = torch.tensor(SYM6)
    nk = k.numel()
    x1, x2, y1, y2 = ada_padding(G_inv, H, W, nk)
    img = F.pad(img, (x1, x2, 0, 0), mode="circular")
    img = F.pad(img, (0, 0, y1, y2), mode="reflect")
    G_inv = _m3([(1, 0, (x1 - x2) / 2), (0, 1, (y1 - y2) / 2), (0, 0, 1)]) @ G_inv
    u0, u1 = (nk + 1) // 2, (nk - 2) // 2
    img = upfirdn2d(img, k[None], up=(2, 1), pad=(u0, u1, 0, 0))
    img = upfirdn2d(img, k[:, None], up=(1, 2), pad=(0, 0, u0, u1))
    G_inv = _m3([(2, 0, 0), (0, 2, 0), (0, 0, 1)]) @ G_inv @ _m3([(.5, 0, 0), (0, .5, 0), (0, 0, 1)])
    G_inv = (_m3([(1, 0, -.5), (0, 1, -.5), (0, 0, 1)]) @ G_inv
             @ _m3([(1, 0, .5), (0, 1, .5), (0, 0, 1)]))
    pad_k = nk // 4
    shape = (B, ch, (H + pad_k * 2) * 2, (W + pad_k * 2) * 2)
    G_inv = (_m3([(2 / img.shape[3], 0, 0), (0, 2 / img.shape[2], 0), (0, 0, 1)]) @ G_inv
             @ _m3([(shape[3] / 2, 0, 0), (0, shape[2] / 2, 0), (0, 0, 1)]))
    grid = F.affine_grid(G_inv[:, :2, :], shape, align_corners=False)
    img = bilinear_warp(img, grid)
    d = -pad_k * 2
    d0, d1 = d + (nk - 1) // 2, d + (nk - 2) // 2
    kf = k.flip(0)
    img = upfirdn2d(img, kf[None], down=(2, 1), pad=(d0, d1, 0, 0))
    img = upfirdn2d(img, kf[:, None], down=(1, 2), pad=(0, 0, d0, d1))
    img = img.reshape(B, ch, H * W)
    if ch == 3:
        img = C[:, :3, :3] @ img + C[:, :3, 3:]
    else:
        Cm = C[:, :3, :].mean(dim=1, keepdim=True)
        img = img * Cm[:, :, :3].sum(dim=2, keepdim=True) + Cm[:, :, 3:]
    return img.reshape(B, ch, H, W)


# --------------------------------------------------------------------------------------
# a16  one training iteration on the CPU (timing baseline + step-level parity)
#      gans/trainer.py:247-451
# --------------------------------------------------------------------------------------


def warmup_dropout(x: Tensor, keep: Optional[Tensor], raydrop_const: float = -1.0) -> Tensor:
    """Trainer.warmup with blur sigma 0 (trainer.py:241-244); `keep` is the Bernoulli draw."""
    if keep is None:
        return x
    return keep * x + (1 - keep) * raydrop_const


def train_iteration(sdG: Dict[str, Tensor], sdD: Dict[str, Tensor], x_real: Tensor, angle: Tensor,
                    rnd: Dict[str, Tensor], with_r1: bool = True, gp_weight: float = 16.0,
                    arch: str = "dusty_v2"):
    """Losses and gradients of one iteration: G step, D step and (optionally) the R1 step,
    with every random draw supplied in `rnd` (z_g, z_d, shift_g, shift_d, u_g, u_d, keep_*,
    Ginv_*, C_*).  Returns dict(loss_G, loss_D, r1, grads_G, grads_D, grads_R1).
    arch "dusty_v2" (default) or "dusty_v1" / "vanilla" (BASELINE config 3: transposed-conv
    generator, strided-conv discriminator; no angle input, no azimuth shift; the Gumbel uniforms
    are used by dusty_v1 only)."""
    def aug(x, tag):
        return ada_apply(warmup_dropout(x, rnd.get(f"keep_{tag}")), rnd[f"Ginv_{tag}"], rnd[f"C_{tag}"])

    if arch != "dusty_v2":
        def generator(sd, z, _angle, u, training=True, shifts_rad=None):      # noqa: F811
            return vanilla_generator(sd, z, u if arch == "dusty_v1" else None)

        discriminator = vanilla_discriminator                                # noqa: F811
    else:
        generator, discriminator = globals()["generator"], globals()["discriminator"]

    import time
    pG = {k: v for k, v in sdG.items() if v.requires_grad}
    pD = {k: v for k, v in sdD.items() if v.requires_grad}
    out = {}
    t0 = time.perf_counter()
    # G step
    fake = generator(sdG, rnd["z_g"], angle, rnd["u_g"], training=True,
                     shifts_rad=rnd["shift_g"] * (2 * np.pi))["image"]
    loss_g = nsgan_g(discriminator(sdD, aug(fake, "g_fake")))
    out["loss_G"] = loss_g.detach()
    out["grads_G"] = dict(zip(pG, torch.autograd.grad(loss_g, list(pG.values()), allow_unused=True)))
    out["t_G"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    # D step
    with torch.no_grad():
        fake = generator(sdG, rnd["z_d"], angle, rnd["u_d"], training=True,
                         shifts_rad=rnd["shift_d"] * (2 * np.pi))["image"]
    y_real = discriminator(sdD, aug(x_real, "d_real").detach())
    y_fake = discriminator(sdD, aug(fake, "d_fake").detach())
    loss_d = nsgan_d(y_real, y_fake)
    out["loss_D"] = loss_d.detach()
    out["grads_D"] = dict(zip(pD, torch.autograd.grad(loss_d, list(pD.values()), allow_unused=True)))
    out["t_D"] = time.perf_counter() - t0
    if with_r1:
        t0 = time.perf_counter()
        x_gp = x_real.detach().clone().requires_grad_()
        y = discriminator(sdD, aug(x_gp, "r1"))
        (gx,) = torch.autograd.grad(y.sum(), x_gp, create_graph=True)
        r1 = r1_penalty(gx)
        loss = (gp_weight / 2) * r1 + 0.0 * y.squeeze()[0]
        out["r1"] = r1.detach()
        out["grads_R1"] = dict(zip(pD, torch.autograd.grad(loss, list(pD.values()), allow_unused=True)))
        out["t_R1"] = time.perf_counter() - t0
    return out


# --------------------------------------------------------------------------------------
# a15  vanilla / dusty_v1 baselines                 gans/models/vanilla.py, dusty_v1.py:31-41
# --------------------------------------------------------------------------------------


def _equal_convT(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int, padding: int) -> Tensor:
    """EqualLR(nn.ConvTranspose2d): the runtime scale is 1/sqrt(weight[0].numel()) -- for a
    transposed conv weight [in, out, kh, kw] that is out*kh*kw (common.py:171-172)."""
    scale = 1.0 / math.sqrt(w[0].numel())
    return F.conv_transpose2d(x * scale, w, b, stride=stride, padding=padding)


def vanilla_synthesis(sd: Dict[str, Tensor], w: Tensor, prefix: str = "synthesis_network.",
                      image_tanh: bool = False) -> Dict[str, Tensor]:
    """vanilla.SynthesisNetwork (vanilla.py:49-69): Projection, 3x Upsample, Head."""
    h = w.reshape(w.shape[0], -1, 1, 1)                                    # "B 1 C -> B C 1 1"
    h = bias_act(_equal_convT(h, sd[f"{prefix}0.1.module.weight"], None, 1, 0), sd[f"{prefix}0.2.bias"])
    for i in (1, 2, 3):
        h = pad2d(h, 1, ring=True, mode="reflect")
        h = bias_act(_equal_convT(h, sd[f"{prefix}{i}.1.module.weight"], None, 2, 3),
                     sd[f"{prefix}{i}.2.bias"])
    out = {}
    for name in ("image", "raydrop_logit"):
        key = f"{prefix}4.heads.{name}.1.module.weight"
        if key in sd:
            o = _equal_convT(pad2d(h, 1, ring=True, mode="reflect"), sd[key],
                             sd[f"{prefix}4.heads.{name}.1.module.bias"], 2, 3)
            out[name] = torch.tanh(o) if (image_tanh and name == "image") else o
    return out


def vanilla_generator(sd: Dict[str, Tensor], z: Tensor, u: Optional[Tensor] = None,
                      raydrop_const: float = -1.0, image_tanh: bool = False) -> Dict[str, Tensor]:
    """vanilla.Generator / dusty_v1.Generator in eval mode (mapping = Identity, one style)."""
    o = vanilla_synthesis(sd, z[:, None, :], image_tanh=image_tanh)
    o["w"] = z[:, None, :]
    if "raydrop_logit" in o and u is not None:
        o.update(raydrop(o["image"], o["raydrop_logit"], u, raydrop_const))
    return o


def vanilla_discriminator(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """vanilla.Discriminator (vanilla.py:94-105): BlurVH, 4x [Pad(1,reflect) + conv4x4 s2 +
    FusedLeakyReLU], full-size conv to one logit."""
    h = blur_vh(x)
    for i in (1, 2, 3, 4):
        h = pad2d(h, 1, ring=True, mode="reflect")
        h = bias_act(equal_conv2d(h, sd[f"{i}.1.module.weight"], None, 2), sd[f"{i}.2.bias"])
    return equal_conv2d(h, sd["5.module.weight"], sd["5.module.bias"], 1)


# ---- BASELINE config 5: GAN inversion losses (reference gans/inversion.py) -----------------
def masked_loss(img_ref: Tensor, img_gen: Tensor, mask: Tensor, loss: str = "l1",
                relative: bool = True) -> Tensor:
    """inversion.py:23-30 -- per-sample masked (relative) L1 / L2."""
    d = (img_ref - img_gen).abs() if loss == "l1" else (img_ref - img_gen).square()
    if relative:
        d = (d * mask) / (img_ref + 1e-11)
    d = (d * mask).sum(dim=(1, 2, 3))
    return d / (mask.sum(dim=(1, 2, 3)) + 1e-8)


def _blurpool3(x: Tensor, taps: Tensor) -> Tensor:
    """Pad(1, replicate H, circular W) + depthwise 3x3 stride-2 correlation
    (inversion.py:52-64)."""
    xp = pad2d(x, 1, ring=True, mode="replicate")
    C = x.shape[1]
    return F.conv2d(xp, taps[None, None].repeat(C, 1, 1, 1), stride=2, groups=C)


def multiscale_masked_loss(gen: Tensor, ref: Tensor, mask: Tensor, level: Optional[int] = None,
                           loss: str = "l1", relative: bool = True) -> Tensor:
    """MultiScaleMaskedLoss.forward (inversion.py:66-80): masked loss on a blur-pool pyramid;
    the mask is pooled with a box filter, renormalised by 9 / (number of valid pixels) and
    re-binarised at every level."""
    H = gen.shape[2]
    level = int(np.log2(H)) if level is None else level
    k1 = torch.tensor([1.0, 2.0, 1.0])
    blur = torch.outer(k1, k1)
    blur = blur / blur.sum()
    ones = torch.ones(3, 3)
    total = 0
    for _ in range(max(1, level)):
        total = total + masked_loss(ref, gen, mask, loss, relative)
        cnt = _blurpool3(mask, ones)
        norm = 9.0 / cnt.masked_fill(cnt == 0, 1.0)
        new_mask = torch.ones_like(cnt).masked_fill(cnt == 0, 0.0)
        gen = _blurpool3(gen * mask, blur) * norm
        ref = _blurpool3(ref * mask, blur) * norm
        mask = new_mask
    return total


def geocross_loss(latents: Tensor) -> Tensor:
    """inversion.py:83-91 (PULSE geodesic cross loss on [B, N, D] latents)."""
    Bn, N, D = latents.shape
    X, Y = latents.view(Bn, 1, N, D), latents.view(Bn, N, 1, D)
    A = ((X - Y).pow(2).sum(-1) + 1e-9).sqrt()
    Bm = ((X + Y).pow(2).sum(-1) + 1e-9).sqrt()
    Dm = 2 * torch.atan2(A, Bm)
    return (Dm.pow(2) * Dm).mean((1, 2)) / 8.0


def spherical_project_(param: Tensor) -> Tensor:
    """SphericalOptimizer.step's projection (inversion.py:17-19): unit RMS along the last axis."""
    with torch.no_grad():
        param.div_(param.pow(2).mean(dim=-1, keepdim=True).add(1e-9).sqrt())
    return param


def inv_depth_norm_to_depth_norm(x: Tensor, min_depth: float, max_depth: float,
                                 tol: float = 1e-11) -> Tensor:
    """convert(x, 'inv_depth_norm', 'depth_norm') (coords.py:139-144 -> 127-135 -> 101-103):
    inv = x/min; depth = mask(inv) / (inv + tol); depth / max.  (The `x > tol` flag computed at
    coords.py:141 is unused on this branch.)"""
    inv = x / min_depth
    valid = ((inv >= 1 / max_depth) & (inv <= 1 / min_depth) & (inv > 0.0)).float()
    return (1 / (inv + tol) * valid) / max_depth


def inversion_targets(depth: Tensor, mask: Tensor, min_depth: float, max_depth: float):
    """demo_inversion.py:89-94: metric depth + mask -> (t_depth [depth_norm], t_inv_depth
    [inv_depth_norm, masked])."""
    t_depth = depth / max_depth
    t_inv = depth_to_inv_depth_norm(t_depth * max_depth, min_depth, max_depth) * mask
    return t_depth, t_inv


def inversion_lr_schedule(iteration: int, num_steps: int, rampup: float = 0.05,
                          rampdown: float = 0.25) -> float:
    """demo_inversion.py:140-146 (StyleGAN2 projector schedule)."""
    t = iteration / num_steps
    gamma = min(1.0, (1.0 - t) / rampdown)
    gamma = 0.5 - 0.5 * math.cos(gamma * math.pi)
    return gamma * min(1.0, t / rampup)


def inversion_forward(sdG: Dict[str, Tensor], z: Tensor, angle: Tensor, t_depth: Tensor,
                      t_inv_depth: Tensor, t_mask: Tensor, min_depth: float, max_depth: float,
                      latent_type: str = "w", phase: Optional[Tensor] = None,
                      u: Optional[Tensor] = None):
    """One evaluation of the inversion objective, demo_inversion.py:149-191 (perturb_z off):
    eval-mode G driven by styles (input_w=True) at angle + phase, tanh_to_sigmoid, the
    multi-scale masked relative-L1 loss (level 2) on depth_norm and on inv_depth_norm, plus
    5e-3 * geocross for 'w+'.  Returns (outputs, per-sample loss [B])."""
    n_styles = 2 * _num_levels(sdG, "synthesis_network.")
    w = torch.stack([z] * n_styles, dim=1) if latent_type == "w" else z
    B = w.shape[0]
    ang = angle if phase is None else angle + phase
    if ang.shape[0] != B:
        ang = ang.expand(B, -1, -1, -1)
    if u is None:
        u = torch.full((B, 1) + tuple(angle.shape[-2:]), 0.5)
    imgs = generator(sdG, w, ang, u, training=False, input_w=True)
    inv_orig = (imgs["image_orig"] + 1.0) / 2.0
    g_depth = inv_depth_norm_to_depth_norm(inv_orig, min_depth, max_depth)
    loss = 0
    if latent_type == "w+":
        loss = loss + 5e-3 * geocross_loss(w)
    loss = loss + multiscale_masked_loss(g_depth, t_depth, t_mask, level=2)
    loss = loss + multiscale_masked_loss(inv_orig, t_inv_depth, t_mask, level=2)
    out = dict(inv_depth=(imgs["image"] + 1.0) / 2.0, inv_depth_orig=inv_orig,
               raydrop_prob=torch.sigmoid(imgs["raydrop_logit"]), g_depth=g_depth,
               raydrop_logit=imgs["raydrop_logit"], raydrop_mask=imgs["raydrop_mask"])
    return out, loss


# --------------------------------------------------------------------------------------
# f4  KITTI scan -> range image                        gans/datasets/kitti.py:216-220,275-279,317-370
# --------------------------------------------------------------------------------------


def scan_cells(points: np.ndarray, H: int = 64, W: int = 2048, scan_unfolding: bool = True):
    """Per-point image cell and depth (kitti.py:319-363), ring assignment as the sequential walk
    the reference does: rings are the runs between 4th->1st quadrant transitions, numbered from
    the LAST run (H - 1) downwards; the walk assigns one run more than H (index -1) before it
    stops; runs further down and the points before the first transition stay at 0."""
    pts = np.asarray(points, dtype=np.float32).reshape(-1, 4)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    depth = np.linalg.norm(pts[:, :3], ord=2, axis=1)
    n = len(pts)
    if scan_unfolding:
        quad = np.zeros(n, dtype=np.int32)
        quad[(x < 0) & (y >= 0)] = 1
        quad[(x < 0) & (y < 0)] = 2
        quad[(x >= 0) & (y < 0)] = 3
        starts = [i for i in range(n) if quad[i - 1] - quad[i] == 3]
        bounds = starts + [n]
        cell_h = np.zeros(n, dtype=np.int32)
        ring = H - 1
        for j in range(len(starts) - 1, -1, -1):
            cell_h[bounds[j]:bounds[j + 1]] = ring
            if ring < 0:
                break
            ring -= 1
    else:
        up, down = np.deg2rad(3), np.deg2rad(-25)
        pitch = np.arcsin(z / depth) + abs(down)
        cell_h = np.floor((1 - pitch / (up - down)) * H).clip(0, H - 1).astype(np.int32)
    yaw = -np.arctan2(y, x)
    cell_w = np.floor(((yaw / np.pi + 1) / 2 % 1) * W).clip(0, W - 1).astype(np.int32)
    return cell_h, cell_w, depth


def scan_to_image(points: np.ndarray, H: int = 64, W: int = 2048, W_out: int = 512, min_depth: float = 0.9,
                  max_depth: float = 120.0, scan_unfolding: bool = True) -> np.ndarray:
    """[6, H, W_out] (x, y, z, reflectance, depth, mask) * mask: points written far-to-near into
    the [H, W] image (kitti.py:216-220,364-368: the nearest point of a cell survives), every
    (W / W_out)-th column kept (NEAREST resize, 276-277), product with the mask channel (278)."""
    pts = np.asarray(points, dtype=np.float32).reshape(-1, 4)
    cell_h, cell_w, depth = scan_cells(pts, H, W, scan_unfolding)
    mask = ((depth >= min_depth) & (depth <= max_depth)).astype(np.float32)
    rows = np.concatenate([pts, depth[:, None].astype(np.float32), mask[:, None]], axis=1)
    img = np.zeros((H, W, 6), dtype=np.float32)
    for i in np.argsort(-depth, kind="stable"):
        img[cell_h[i], cell_w[i]] = rows[i]
    out = img[:, :: W // W_out].transpose(2, 0, 1).copy()
    return out * out[5:6]

