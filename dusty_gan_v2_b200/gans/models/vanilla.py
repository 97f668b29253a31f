"""Drop-in for gans/models/vanilla.py (reference 7-105): the transposed-convolution generator and
strided-convolution discriminator used as baselines (BASELINE config 3), with the reference's
module tree and state_dict keys:

  synthesis_network.0        Projection   z -> [B, C3, H/16, W/16]   (ConvT with the full-map kernel)
  synthesis_network.{1,2,3}  Upsample     Pad(1, reflect / ring) -> ConvT 4x4 stride 2 pad 3 -> bias + lrelu
  synthesis_network.4        Head         one Pad + ConvT per output map (image, raydrop_logit)
  discriminator 0            BlurVH
  discriminator {1..4}       Downsample   Pad(1) -> Conv 4x4 stride 2 -> bias + lrelu
  discriminator 5            Conv with the full-map kernel -> one logit

Every layer runs on dusty_b200 kernels: the dense (transposed) convolutions on the tcgen05
implicit-GEMM family in bf16 mode (ops.EqualLR._conv_low_precision: a transposed convolution is
the data-gradient kernel) and on the CUDA-core family in fp32 parity mode; the padding, the blur
pair, the fused bias + leaky-ReLU and everything downstream of the heads likewise.
"""
from torch import nn

from . import base, ops

_LEVELS = 4          # three 2x stages after the projection, one more inside the head -> 16x overall


def _widths(ch_base, ch_max):
    """Channel plan, finest level first: ch_base * 2^i capped at ch_max."""
    return [min(ch_base << i, ch_max) for i in range(_LEVELS)]


def _activation(spec):
    """Head activations arrive as None, a class, or the class's dotted name (YAML configs)."""
    if spec is None:
        return nn.Identity()
    return (eval(spec) if isinstance(spec, str) else spec)()


def _ring_pad(ring):
    return ops.Pad(padding=1, ring=ring, mode="reflect")


def _up_conv(in_ch, out_ch, bias):
    # 4x4 / stride 2 / padding 3 on the 1-pixel-padded map: exactly 2x the un-padded size
    return ops.EqualLR(nn.ConvTranspose2d(in_ch, out_ch, 4, 2, 3, bias=bias))


class _ToMap(nn.Module):
    """[B, 1, C] style -> [B, C, 1, 1] feature map (parameter-free; index 0 of Projection)."""

    def forward(self, w):
        return w.transpose(1, 2).unsqueeze(-1)


class Projection(nn.Sequential):
    def __init__(self, in_ch, out_ch, kernel):
        first = ops.EqualLR(nn.ConvTranspose2d(in_ch, out_ch, kernel, 1, 0, bias=False))
        super().__init__(_ToMap(), first, ops.FusedLeakyReLU(out_ch))


class Upsample(nn.Sequential):
    def __init__(self, in_ch, out_ch, ring=True):
        super().__init__(_ring_pad(ring), _up_conv(in_ch, out_ch, bias=False), ops.FusedLeakyReLU(out_ch))


class Head(nn.Module):
    """`heads[name]` = Pad -> ConvT(in_ch -> ch) -> activation, for every output with ch > 0."""

    def __init__(self, in_ch, out_ch, ring=True):
        super().__init__()
        self.in_ch = in_ch
        self.heads = nn.ModuleDict({
            spec["name"]: nn.Sequential(_ring_pad(ring), _up_conv(in_ch, spec["ch"], bias=True),
                                        _activation(spec["act"]))
            for spec in out_ch if spec["ch"] != 0})

    def forward(self, x):
        # head maps leave the network in fp32 whatever the trunk's precision (as dusty_v2's do)
        return {name: branch(x).float().contiguous() for name, branch in self.heads.items()}


class SynthesisNetwork(nn.Sequential):
    def __init__(self, in_ch, out_ch, ch_base=64, ch_max=512, resolution=(64, 256), ring=True):
        self.in_ch, self.out_ch, self.num_styles = in_ch, out_ch, 1
        fine_to_coarse = _widths(ch_base, ch_max)
        seed_map = (resolution[0] >> _LEVELS, resolution[1] >> _LEVELS)
        stages = [Projection(in_ch, fine_to_coarse[-1], seed_map)]
        for level in range(_LEVELS - 1, 0, -1):
            stages.append(Upsample(fine_to_coarse[level], fine_to_coarse[level - 1], ring))
        stages.append(Head(fine_to_coarse[0], out_ch, ring))
        super().__init__(*stages)


class Generator(base.Generator):
    def __init__(self, synthesis_kwargs):
        super().__init__(mapping_network=nn.Identity(), synthesis_network=SynthesisNetwork(**synthesis_kwargs),
                         measurement_model=nn.Identity())

    def forward_synthesis(self, w, angles=None):
        return self.synthesis_network(w)


class Downsample(nn.Sequential):
    def __init__(self, in_ch, out_ch, ring=True):
        conv = ops.EqualLR(nn.Conv2d(in_ch, out_ch, 4, 2, 0, bias=False))
        super().__init__(_ring_pad(ring), conv, ops.FusedLeakyReLU(out_ch))


class Discriminator(nn.Sequential):
    def __init__(self, in_ch, ch_base=64, ch_max=512, resolution=(64, 256), ring=True):
        widths = [in_ch * 2] + _widths(ch_base, ch_max)          # BlurVH doubles the input channels
        trunk = [Downsample(a, b, ring) for a, b in zip(widths[:-1], widths[1:])]
        full_map = (resolution[0] >> _LEVELS, resolution[1] >> _LEVELS)
        logit = ops.EqualLR(nn.Conv2d(widths[-1], 1, full_map, 1, 0))
        super().__init__(ops.BlurVH(window=[1, 2, 1], ring=ring), *trunk, logit)
