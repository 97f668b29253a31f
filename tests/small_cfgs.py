"""Small model configs shared by CPU and GPU tests (must match tests/golden/make_golden.py)."""
G_SMALL = dict(
    arch="dusty_v2",
    mapping_kwargs=dict(in_ch=16, out_ch=16, depth=2),
    synthesis_kwargs=dict(
        in_ch=16,
        out_ch=[dict(name="image", ch=1, act="nn.Tanh"), dict(name="raydrop_logit", ch=1, act=None)],
        ch_base=4, ch_max=16, resolution=[16, 64], layers=[2, 2, 2, 2], ring=True,
        num_fp16_layers=-1, use_noise=False, pe_type="random", pe_scale_offset=[3, -1],
        aug_coords=True, aug_coords_blitting=False),
    measurement_kwargs=dict(raydrop_const=-1, gumbel_temperature=1),
)
D_SMALL = dict(arch="dusty_v2", layer_kwargs=dict(in_ch=1, ring=True, ch_base=4, ch_max=8,
                                                  resolution=[16, 64], mbdis_group=4, mbdis_feat=1,
                                                  num_fp16_layers=-1, pre_blur=True))

V_SYN = dict(in_ch=16, ch_base=4, ch_max=16, resolution=[32, 64], ring=True)
V1_SMALL = dict(arch="dusty_v1", synthesis_kwargs=dict(V_SYN, out_ch=[
    dict(name="image", ch=1, act=None), dict(name="raydrop_logit", ch=1, act=None)]),
    measurement_kwargs=dict(raydrop_const=-1, gumbel_temperature=1))
VD_SMALL = dict(arch="vanilla", layer_kwargs=dict(in_ch=1, ring=True, ch_base=4, ch_max=16,
                                                  resolution=[32, 64]))
V_SMALL = dict(arch="vanilla", synthesis_kwargs=dict(V_SYN, out_ch=[dict(name="image", ch=1, act=None)]),
               measurement_kwargs={})

# Mid-size dusty_v2 pair whose channel counts qualify for the tcgen05 kernels (G: 64-wide
# features so a 64-channel K block never straddles the feature / Fourier sources, 256-pixel
# lowest level; D: 16..32 channels, multiples of 8) -- tests/golden/trainer_step_mid.npz.
G_MID = dict(
    arch="dusty_v2",
    mapping_kwargs=dict(in_ch=64, out_ch=64, depth=2),
    synthesis_kwargs=dict(
        in_ch=64,
        out_ch=[dict(name="image", ch=1, act="nn.Tanh"), dict(name="raydrop_logit", ch=1, act=None)],
        ch_base=32, ch_max=64, resolution=[32, 128], layers=[2, 2], ring=True,
        num_fp16_layers=-1, use_noise=False, pe_type="random", pe_scale_offset=[3, -1],
        aug_coords=True, aug_coords_blitting=False),
    measurement_kwargs=dict(raydrop_const=-1, gumbel_temperature=1),
)
D_MID = dict(arch="dusty_v2", layer_kwargs=dict(in_ch=1, ring=True, ch_base=16, ch_max=32,
                                                resolution=[32, 128], mbdis_group=4, mbdis_feat=1,
                                                num_fp16_layers=-1, pre_blur=True))


def sample_flat(t, n=2048):
    """Fixed-stride sample of a tensor (all of it when it has at most `n` elements): what the
    compact fixtures store of large gradient / weight tensors, next to their L2 norm."""
    f = t.detach().reshape(-1)
    if f.numel() <= n:
        return f.clone()
    stride = f.numel() // n
    return f[::stride][:n].clone()


def synthetic_scan(rings=66, per_ring=120, seed=0):
    """Velodyne-like [N, 4] float32 scan: `rings` sweeps of increasing azimuth (a new ring starts
    where the azimuth passes from the 4th to the 1st quadrant), elevation falling from +3 to -25
    degrees, ranges 0.5 .. 130 m (some outside any [min_depth, max_depth]); the first points are
    dropped so that the scan starts before any ring start."""
    import numpy as np
    g = np.random.default_rng(seed)
    rows = []
    for r in range(rings):
        phi = np.sort(g.uniform(0, 2 * np.pi, per_ring))
        rng = g.uniform(0.5, 130, per_ring)
        el = np.deg2rad(3 - 28 * r / rings) + g.normal(0, 0.002, per_ring)
        rows.append(np.stack([rng * np.cos(el) * np.cos(phi), rng * np.cos(el) * np.sin(phi), rng * np.sin(el),
                              g.uniform(0, 1, per_ring)], 1))
    return np.concatenate(rows).astype(np.float32)[5:]
