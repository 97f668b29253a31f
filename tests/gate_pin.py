"""Gate pinning for low-precision gradient parity.

The networks of this path are piecewise linear (leaky ReLU, slope 0.2): their gradient is a
DISCONTINUOUS function of the forward activations.  A bf16 forward perturbs pre-activations by
~0.3-0.5 % (measured, tools/debug/d_layer_error.py), which flips the gate of the ~0.2 % of units
whose pre-activation is that close to zero; every flipped gate changes its gradient entry by a
factor of five, so the gradient of a bf16 forward differs from the gradient of an fp32 forward by
~sqrt(fraction flipped) -- 3 % per activation layer, 5-7 % after a dozen -- however exact the
backward kernels are.  That is a property of the function, not of an implementation.

To test the KERNELS at the tolerance the task states for bf16 (2e-2), the comparison is made in
the same linear region: `GateRecorder` collects, in call order, the gate pattern each leaky-ReLU
site of the device forward actually used, and `pinned_oracle_gates` makes the CPU oracle's
`bias_act` use those patterns instead of its own sign test.  At a flipped unit the
pre-activation is ~0, so the oracle's forward values barely move; its backward is then the
derivative of the same linear piece the device differentiated."""
import contextlib

import torch


class GateRecorder:
    """Context manager: records (pre-activation > 0) of every leaky-ReLU site evaluated through
    the package's functional layer, as CPU bool tensors in logical NCHW shape."""

    def __init__(self):
        self.gates = []
        self._saved = []

    def _patch(self, obj, name, fn):
        self._saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, fn)

    def __enter__(self):
        import dusty_gan_v2_b200.functional as DF
        from dusty_gan_v2_b200.gans.models import ops
        rec = self.gates

        def keep(mask):
            rec.append(mask.detach().contiguous().cpu())

        o_bias_act, o_tail, o_bmm, o_stem, o_cba = (DF.bias_act, DF.residual_tail, DF.modconv_bmm, DF.stem,
                                                     ops.conv_bias_act)

        def bias_act(x, bias=None, negative_slope=0.2, scale=2 ** 0.5):
            y = o_bias_act(x, bias, negative_slope, scale)
            keep(y > 0)
            return y

        def residual_tail(pre, bias, skip, *a, **k):
            # the kernel gates on fp32 (pre + bias), pre being the stored (bf16) convolution output
            b = bias.detach().float().reshape(1, -1, 1, 1)
            keep((pre.detach().float() + b) > 0)
            return o_tail(pre, bias, skip, *a, **k)

        def modconv_bmm(wb, x1, x2=None, bias=None, act=1, alpha=0.2, scale=1.0, **kw):
            y = o_bmm(wb, x1, x2, bias, act, alpha, scale, **kw)
            if act == 3:
                keep(y > 0)
            return y

        def stem(*a, **k):
            y = o_stem(*a, **k)
            keep(y > 0)
            return y

        def conv_bias_act(*a, **k):
            # the gate is read off the ACTIVATION's output: with blur_taps the op would return the
            # blurred / padded tensor, so the recorder runs the two stages separately (the fused
            # backward kernel has its own test, test_conv_bias_act_blur_pad_fused_backward)
            taps = k.pop("blur_taps", None)
            y = o_cba(*a, **k)
            keep(y > 0)
            return y if taps is None else DF.blur_pad_cl(y, taps)

        self._patch(DF, "bias_act", bias_act)
        self._patch(DF, "residual_tail", residual_tail)
        self._patch(DF, "modconv_bmm", modconv_bmm)
        self._patch(DF, "stem", stem)
        self._patch(ops, "conv_bias_act", conv_bias_act)
        return self

    def __exit__(self, *exc):
        for obj, name, fn in reversed(self._saved):
            setattr(obj, name, fn)
        self._saved.clear()
        return False


@contextlib.contextmanager
def pinned_oracle_gates(O, gates):
    """`O.bias_act` takes its gates from `gates` (consumed in call order); every site must be
    used, shapes must agree."""
    it = iter(gates)
    used = [0]
    orig = O.bias_act

    def bias_act(x, bias=None, slope=0.2, scale=2 ** 0.5):
        if bias is not None:
            x = x + bias.reshape([1, -1] + [1] * (x.ndim - 2))
        g = next(it)
        assert tuple(g.shape) == tuple(x.shape), (used[0], tuple(g.shape), tuple(x.shape))
        used[0] += 1
        own = x > 0
        frac = float((own != g).float().mean())
        assert frac < 2e-2, f"site {used[0]}: {frac:.4f} of the gates differ -- not the same network state"
        return x * torch.where(g, 1.0, slope) * scale

    O.bias_act = bias_act
    try:
        yield used
    finally:
        O.bias_act = orig
    assert used[0] == len(gates), (used[0], len(gates))
