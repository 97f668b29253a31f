"""GPU parity against the REFERENCE'S OWN CUDA kernels: `fused.fused_bias_act`
(gans/models/ops/fused_act/fused_bias_act_kernel.cu) and `upfirdn2d_op.upfirdn2d`
(gans/models/ops/upfirdn2d/upfirdn2d_kernel.cu), compiled in place from /root/reference by
oracle/build_ref.py into oracle/_ref/ (the built modules travel to the GPU box, the sources
never enter the repo).  Same pybind-level signatures on both sides."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import build_ref  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def ref_fused():
    m = build_ref.load_built("dusty_ref_fused")
    if m is None:
        pytest.skip("oracle/_ref/dusty_ref_fused not built (python oracle/build_ref.py)")
    return m


@pytest.fixture(scope="module")
def ref_ufd():
    m = build_ref.load_built("dusty_ref_upfirdn2d")
    if m is None:
        pytest.skip("oracle/_ref/dusty_ref_upfirdn2d not built (python oracle/build_ref.py)")
    return m


def close(a, b, rtol, atol_rel):
    a, b = a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_rel * max(float(np.abs(b).max()), 1e-12))


@pytest.mark.parametrize("shape", [(3, 5, 7, 9), (4, 32, 64, 512), (6, 16), (2, 8, 33)])
def test_fused_bias_act_vs_reference_cuda(ref_fused, shape):
    """act=3 (leaky ReLU) forward with bias, first-order gradient (grad=1, gated by `refer`) and
    the grad=2 form, empty tensors meaning "absent" (fused_act.py:27,65-72)."""
    from dusty_gan_v2_b200.functional import fused_bias_act
    g = torch.Generator().manual_seed(7)
    x = torch.randn(shape, generator=g).to(DEV)
    b = torch.randn(shape[1], generator=g).to(DEV)
    empty = x.new_empty(0)
    ours = fused_bias_act(x, b, empty, 3, 0, 0.2, 2 ** 0.5)
    ref = ref_fused.fused_bias_act(x, b, empty, 3, 0, 0.2, 2 ** 0.5)
    close(ours, ref, rtol=1e-6, atol_rel=1e-7)
    gy = torch.randn(shape, generator=g).to(DEV)
    for grad in (1, 2):
        ours_g = fused_bias_act(gy, empty, ref, 3, grad, 0.2, 2 ** 0.5)
        ref_g = ref_fused.fused_bias_act(gy, empty, ref, 3, grad, 0.2, 2 ** 0.5)
        close(ours_g, ref_g, rtol=1e-6, atol_rel=1e-7)
    # act=1 (linear) with bias
    close(fused_bias_act(x, b, empty, 1, 0, 0.2, 1.0), ref_fused.fused_bias_act(x, b, empty, 1, 0, 0.2, 1.0),
          rtol=1e-6, atol_rel=1e-7)


SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
        0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252,
        -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148]


@pytest.mark.parametrize("name,in_hw,kshape,up,down,pad", [
    ("ada_up_x", (76, 524), (1, 12), (2, 1), (1, 1), (6, 5, 0, 0)),
    ("ada_up_y", (76, 1048), (12, 1), (1, 2), (1, 1), (0, 0, 6, 5)),
    ("ada_down_x", (140, 1036), (1, 12), (1, 1), (2, 1), (-1, -1, 0, 0)),
    ("ada_down_y", (140, 512), (12, 1), (1, 1), (1, 2), (0, 0, -1, -1)),
    ("general_2d", (17, 23), (4, 3), (2, 2), (1, 1), (2, 1, 1, 2)),
    ("down_2d", (32, 40), (4, 4), (1, 1), (2, 2), (1, 1, 1, 1)),
])
def test_upfirdn2d_vs_reference_cuda(ref_ufd, name, in_hw, kshape, up, down, pad):
    """pybind-level `upfirdn2d(input[major,in_h,in_w,1], kernel, up_x, up_y, down_x, down_y, pad_x0,
    pad_x1, pad_y0, pad_y1)` (upfirdn2d.cpp:17-31) at AdaptiveAugment's shapes and two 2-D cases."""
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d, upfirdn2d_op
    g = torch.Generator().manual_seed(8)
    x = torch.randn(6, in_hw[0], in_hw[1], 1, generator=g).to(DEV)
    if kshape in ((1, 12), (12, 1)):
        k = torch.tensor(SYM6).reshape(kshape).to(DEV)
    else:
        k = torch.randn(kshape, generator=g).to(DEV)
    args = (up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3])
    ref = ref_ufd.upfirdn2d(x, k, *args)
    ours = upfirdn2d_op.upfirdn2d(x, k, *args)
    assert ours.shape == ref.shape
    close(ours, ref, rtol=1e-5, atol_rel=1e-6)
    # the Python-level entry (single-axis fast path for the ADA shapes)
    ours2 = upfirdn2d(x.reshape(2, 3, *in_hw), k, up=up, down=down, pad=pad)
    close(ours2.reshape(ref.shape), ref, rtol=1e-5, atol_rel=1e-6)
