"""Hyper-parameter presets equivalent to the reference's configs/gans/{dusty_v2,dusty_v1,
vanilla}.yaml (same key layout, so a config loaded from the reference's YAML files with
`load_config` is interchangeable).  Dataset paths / checkpoint cadence / validation keys
are omitted: they belong to the data and logging side, which is out of the hot path."""
from .config import to_attr


def _heads(image_act):
    return [dict(name="image", ch=1, act=image_act), dict(name="raydrop_logit", ch=1, act=None)]


def _training(batch_size):
    adam = dict(alpha=0.002, beta1=0, beta2=0.99)
    policy = dict.fromkeys(["lr_flip", "ud_flip", "int_trans", "iso_scale", "frac_trans",
                            "brightness", "contrast", "luma_flip", "hue", "saturation"], 1)
    policy.update(imgfilter=0, noise=0, cutout=0)
    return dict(
        random_seed=0, total_kimg=25000, ema_kimg=10, ema_rampup=0.05, batch_size=batch_size,
        gan_objective="nsgan", loss=dict(gan=1, gp=1, pl=0), lazy=dict(gp=16, pl=4, ada=4),
        lr=dict(generator=dict(adam), discriminator=dict(adam)),
        augment=dict(p_init=0.0, p_target=0.6, kimg=500, policy=policy),
        warmup=dict(fade_kimg=200, blur_init_sigma=0, dropout_init_ratio=0.5),
        amp=dict(main=False, reg=False))


def preset(arch: str = "dusty_v2", batch_size: int = 32, resolution=(64, 512)):
    res = list(resolution)
    dataset = dict(name="kitti_raw", min_depth=1.45, max_depth=80, raydrop_const=-1)
    meas = dict(raydrop_const=-1, gumbel_temperature=1)
    if arch == "dusty_v2":
        gen = dict(arch="dusty_v2",
                   mapping_kwargs=dict(in_ch=512, out_ch=512, depth=2),
                   synthesis_kwargs=dict(in_ch=512, out_ch=_heads("nn.Tanh"), ch_base=32,
                                         ch_max=512, resolution=res, layers=[2, 2, 2, 2],
                                         ring=True, num_fp16_layers=-1, use_noise=False,
                                         pe_type="random", pe_scale_offset=[3, -1],
                                         aug_coords=True, aug_coords_blitting=False),
                   measurement_kwargs=meas)
        disc = dict(arch="dusty_v2",
                    layer_kwargs=dict(in_ch=1, ring=True, ch_base=32, ch_max=512, resolution=res,
                                      mbdis_group=4, mbdis_feat=1, num_fp16_layers=-1,
                                      pre_blur=True))
    elif arch in ("dusty_v1", "vanilla"):
        syn = dict(in_ch=512, ch_base=64, ch_max=512, resolution=res, ring=True)
        if arch == "dusty_v1":
            gen = dict(arch="dusty_v1", mapping_kwargs=dict(in_ch=512, out_ch=512),
                       synthesis_kwargs=dict(syn, out_ch=_heads(None)), measurement_kwargs=meas)
        else:
            gen = dict(arch="vanilla", mapping_kwargs=dict(in_ch=512, out_ch=512),
                       synthesis_kwargs=dict(syn, out_ch=[dict(name="image", ch=1, act=None)]),
                       measurement_kwargs={})
        disc = dict(arch="vanilla", layer_kwargs=dict(in_ch=1, ring=True, ch_base=64, ch_max=512,
                                                      resolution=res))
    else:
        raise ValueError(arch)
    return to_attr(dict(dataset=dataset, training=_training(batch_size), random_seed=0,
                        model=dict(generator=gen, discriminator=disc)))
