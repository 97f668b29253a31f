"""Drop-in for gans/models/vanilla.py: transposed-conv generator / strided-conv
discriminator baselines (reference 7-105).  Dense (transposed) convolutions are library
calls; padding, blur and bias+activation are dusty_b200 kernels."""
from torch import nn

from . import base, ops


def _act(spec):
    if spec is None:
        return nn.Identity()
    return eval(spec)() if isinstance(spec, str) else spec()


class _ToMap(nn.Module):
    def forward(self, w):                 # [B, 1, C] -> [B, C, 1, 1]
        return w.transpose(1, 2).unsqueeze(-1)


class Projection(nn.Sequential):
    def __init__(self, in_ch, out_ch, kernel):
        super().__init__(_ToMap(),
                         ops.EqualLR(nn.ConvTranspose2d(in_ch, out_ch, kernel, 1, 0, bias=False)),
                         ops.FusedLeakyReLU(out_ch))


class Upsample(nn.Sequential):
    def __init__(self, in_ch, out_ch, ring=True):
        super().__init__(ops.Pad(padding=1, ring=ring, mode="reflect"),
                         ops.EqualLR(nn.ConvTranspose2d(in_ch, out_ch, 4, 2, 3, bias=False)),
                         ops.FusedLeakyReLU(out_ch))


class Head(nn.Module):
    def __init__(self, in_ch, out_ch, ring=True):
        super().__init__()
        self.in_ch = in_ch
        self.heads = nn.ModuleDict()
        for o in out_ch:
            if o["ch"] == 0:
                continue
            self.heads[o["name"]] = nn.Sequential(
                ops.Pad(padding=1, ring=ring, mode="reflect"),
                ops.EqualLR(nn.ConvTranspose2d(in_ch, o["ch"], 4, 2, 3, bias=True)),
                _act(o["act"]))

    def forward(self, x):
        return {name: head(x) for name, head in self.heads.items()}


class SynthesisNetwork(nn.Sequential):
    def __init__(self, in_ch, out_ch, ch_base=64, ch_max=512, resolution=(64, 256), ring=True):
        self.in_ch, self.out_ch, self.num_styles = in_ch, out_ch, 1
        ch = [min(ch_base << i, ch_max) for i in range(4)]
        super().__init__(Projection(in_ch, ch[3], (resolution[0] >> 4, resolution[1] >> 4)),
                         Upsample(ch[3], ch[2], ring), Upsample(ch[2], ch[1], ring),
                         Upsample(ch[1], ch[0], ring), Head(ch[0], out_ch, ring))


class Generator(base.Generator):
    def __init__(self, synthesis_kwargs):
        super().__init__(mapping_network=nn.Identity(),
                         synthesis_network=SynthesisNetwork(**synthesis_kwargs),
                         measurement_model=nn.Identity())

    def forward_synthesis(self, w, angles=None):
        return self.synthesis_network(w)


class Downsample(nn.Sequential):
    def __init__(self, in_ch, out_ch, ring=True):
        super().__init__(ops.Pad(padding=1, ring=ring, mode="reflect"),
                         ops.EqualLR(nn.Conv2d(in_ch, out_ch, 4, 2, 0, bias=False)),
                         ops.FusedLeakyReLU(out_ch))


class Discriminator(nn.Sequential):
    def __init__(self, in_ch, ch_base=64, ch_max=512, resolution=(64, 256), ring=True):
        ch = [min(ch_base << i, ch_max) for i in range(4)]
        super().__init__(ops.BlurVH(window=[1, 2, 1], ring=ring),
                         Downsample(in_ch * 2, ch[0], ring), Downsample(ch[0], ch[1], ring),
                         Downsample(ch[1], ch[2], ring), Downsample(ch[2], ch[3], ring),
                         ops.EqualLR(nn.Conv2d(ch[3], 1, (resolution[0] >> 4, resolution[1] >> 4),
                                               1, 0)))
