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
