"""Adam on this package's multi-tensor kernel (csrc/optim.cu), state-compatible with
torch.optim.Adam -- `state_dict()` / `load_state_dict()` carry the same `step` / `exp_avg` /
`exp_avg_sq` entries, so checkpoints written by the reference's trainer (gans/trainer.py:128-171,
551-567) resume here and vice versa -- with two additions of the training step folded into the
same pass over the weights: a gradient scale (1 / world_size after a summing all-reduce of a
flat bucket) and the generator's EMA lerp towards the new weights (reference ema_inplace,
trainer.py:30-41: G does not change between its optimiser step and the end of the iteration,
so lerping there is the same arithmetic as lerping at the exit).
"""
import ctypes as C

import torch

from .. import _cabi as K


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def multi_copy(dst, src, scale: float = 1.0):
    """dst[i] <- src[i] * scale for lists of fp32 CUDA tensors (one launch per 32 tensors)."""
    if not dst:
        return
    for d, s in zip(dst, src):
        if (d.dtype != torch.float32 or s.dtype != torch.float32 or not d.is_contiguous()
                or not s.is_contiguous() or d.numel() != s.numel() or not d.is_cuda or not s.is_cuda):
            raise RuntimeError("multi_copy: contiguous fp32 CUDA tensors of equal size only")
    n = (C.c_longlong * len(dst))(*[d.numel() for d in dst])
    K.call("dusty_multi_copy", _ptr_array(dst), _ptr_array(src), n, len(dst), float(scale), K.stream_of(dst[0]))


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, ema_params=None):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        plist = [p for g in self.param_groups for p in g["params"]]
        self._ema = None
        if ema_params is not None:
            ema_params = list(ema_params)
            if len(ema_params) != len(plist):
                raise ValueError("ema_params must pair one-to-one with params")
            self._ema = {id(p): e for p, e in zip(plist, ema_params)}
        self._cache = {}

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None, grads=None, grad_scale: float = 1.0, ema_weight: float = 0.0):
        """grads: optional list aligned with the parameters that have a gradient, in parameter
        order (views of a reduced flat bucket); default: each parameter's `.grad`."""
        if closure is not None:
            raise NotImplementedError("closures are not used on this path")
        gi = 0
        for group in self.param_groups:
            todo = {}                                   # update count -> tensors
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad if grads is None else grads[gi]
                gi += 1
                if not p.is_cuda:
                    raise RuntimeError("dusty Adam runs on CUDA tensors only (no CPU fallback)")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("dusty Adam: contiguous fp32 parameters only")
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                st = self._state_of(p)
                st["step"] += 1
                e = self._ema.get(id(p)) if (self._ema is not None and ema_weight != 0.0) else None
                todo.setdefault(int(st["step"]), []).append((p, g, st["exp_avg"], st["exp_avg_sq"], e))
            b1, b2 = group["betas"]
            for step, items in todo.items():
                ps, gs, ms, vs, es = zip(*items)
                key = (id(group), step > 0, tuple(id(p) for p in ps), ema_weight != 0.0)
                cached = self._cache.get(key)
                if cached is None:                      # parameter / state pointers are stable
                    ema_arr = None
                    if any(e is not None for e in es):
                        ema_arr = (C.c_void_p * len(es))(*[None if e is None else e.data_ptr() for e in es])
                    cached = (_ptr_array(ps), _ptr_array(ms), _ptr_array(vs), ema_arr,
                              (C.c_longlong * len(ps))(*[p.numel() for p in ps]))
                    self._cache[key] = cached
                pa, ma, va, ea, na = cached
                K.call("dusty_multi_adam", pa, _ptr_array(gs), ma, va, ea, na, len(ps), float(group["lr"]),
                       float(b1), float(b2), float(group["eps"]), step, float(ema_weight), float(grad_scale),
                       K.stream_of(ps[0]))
        return None

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._cache.clear()
        for st in self.state.values():                 # torch moves `step` next to the parameter
            if isinstance(st.get("step"), torch.Tensor):
                st["step"] = st["step"].detach().float().cpu()
