"""ctypes binding of libdusty_b200.so (the C ABI declared in include/dusty_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails the
caller gets an exception.  PyTorch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DUSTY_LIB: tools may point at an instrumented build of the same library (csrc/Makefile `prof`)
LIB_PATH = os.environ.get("DUSTY_LIB") or os.path.join(_HERE, "libdusty_b200.so")

F32, BF16 = 0, 1
PAD_ZERO, PAD_CIRCULAR, PAD_REPLICATE, PAD_REFLECT = 0, 1, 2, 3
ABI_VERSION = 1

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every entry returns int status unless listed in _RESTYPE
SIGNATURES = {
    "dusty_abi_version": [],
    "dusty_last_error": [],
    "dusty_query_sm": [],
    "dusty_launch_count": [],
    "dusty_bias_act": [_vp, _vp, _vp, _vp, _i64, _i, _i64, _i, _i, _f, _f, _i, _vp],
    "dusty_bias_act_bwd": [_vp, _vp, _vp, _vp, _i64, _i, _i64, _f, _f, _i, _vp],
    "dusty_fir2d": [_vp, _vp, _vp, _i, _i, _i, _i64] + [_i] * 13 + [_vp],
    "dusty_fir2d_adj": [_vp, _vp, _vp, _i, _i, _i, _i64] + [_i] * 13 + [_vp],
    "dusty_upfirdn2d": [_vp, _vp, _vp, _i64] + [_i] * 13 + [_vp],
    "dusty_resample4": [_vp, _vp, _f, _f, _f, _f, _i64, _i, _i, _i, _i, _i, _vp],
    "dusty_up2_sumsq": [_vp, _vp, _vp, _f, _f, _f, _f, _i64, _i, _i, _i, _vp],
    "dusty_pad2d": [_vp, _vp, _i64] + [_i] * 10 + [_vp],
    "dusty_bias_act_cl": [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _f, _i, _vp],
    "dusty_bias_act_bwd_cl": [_vp, _vp, _vp, _vp, _i64, _i, _f, _f, _i, _vp],
    "dusty_bias_act_add_cl": [_vp, _vp, _vp, _vp, _i64, _i, _f, _f, _f, _i, _vp],
    "dusty_bias_act_add_bwd_cl": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _f, _f, _i, _vp],
    "dusty_pad2d_cl": [_vp, _vp] + [_i] * 12 + [_vp],
    "dusty_blur4_cl": [_vp, _vp, _f, _f, _f, _f] + [_i] * 7 + [_vp],
    "dusty_blur4_down2_cl": [_vp, _vp, _f, _f, _f, _f] + [_i] * 6 + [_vp],
    "dusty_residual_fork_bwd_cl": [_vp, _vp, _vp, _f, _f, _f, _f] + [_i] * 5 + [_vp],
    "dusty_fir1d": [_vp, _vp, _vp, _i, _i, _i64] + [_i] * 7 + [_vp],
    "dusty_affine_warp": [_vp, _vp, _vp] + [_i] * 7 + [_vp],
    "dusty_fourier": [_vp, _vp, _vp, _vp, _i, _i, _i64, _i, _vp],
    "dusty_angle_down2": [_vp, _vp, _i, _i, _i, _vp],
    "dusty_modconv_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i, _f, _f, _i, _i, _i, _vp, _vp, _vp, _vp],
    "dusty_modconv_bwd_dx": [_vp, _vp, _vp, _i, _i, _i, _i, _i64, _i, _i, _i, _vp, _vp, _vp],
    "dusty_modconv_bwd_dw": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i, _i, C.c_longlong, _vp],
    "dusty_modprep_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _i, _vp, _i, _i, _vp],
    "dusty_modprep_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _i, _i, _vp, _vp],
    "dusty_gumbel_raydrop_fwd": [_vp] * 7 + [_i64, _f, _f, _vp],
    "dusty_gumbel_raydrop_bwd": [_vp] * 7 + [_i64, _f, _vp],
    "dusty_point_project": [_vp, _vp, _vp, _vp, _i, _i64, _f, _f, _f, _i, _vp],
    "dusty_minibatch_std_fwd": [_vp, _vp, _vp, _i, _i, _i64, _i, _f, _i, _vp],
    "dusty_minibatch_std_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i64, _i, _f, _i, _vp],
    "dusty_sumsq_rows": [_vp, _vp, _i64, _i64, _i, _i, _vp],
    "dusty_ema_lerp": [_vp, _vp, _vp, _f, _f, _f, _vp],
    "dusty_circular_shift": [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp],
    "dusty_weight_prep": [_vp, _vp, _vp, _i, _i, _i, _f, _i, _vp],
    "dusty_weight_prep_adj": [_vp, _vp, _i, _i, _i, _f, _i, _i, _vp],
    "dusty_filter_rsco_to_ohwi": [_vp, _vp, _i, _i, _i, _i, _vp],
    "dusty_stem_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _i, _vp],
    "dusty_stem_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _i, _vp],
    "dusty_stem_dx": [_vp, _vp, _i, _i, _i, _f, _f, _f, _vp],
    "dusty_conv2d_tc": [_vp, _vp, _vp, _vp] + [_i] * 9 + [_vp, _vp, _i, _i, _i]
                       + [C.c_longlong] * 4 + [_i, _f, _f, C.c_longlong, C.c_longlong, _vp, _i, _i, _vp],
    "dusty_conv2d_tc_classes": [_vp, _vp, _vp] + [_i] * 6 + [_vp] * 7 + [C.c_longlong] * 5 + [_i, _i, _vp],
    "dusty_residual_fork_fwd_cl": [_vp, _vp, _vp, _f, _f, _f, _f, _i, _i, _i, _i, _i, _vp],
    "dusty_blur4_cl_adj_act": [_vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _i, _i, _i, _f, _f, _i, _vp],
    "dusty_ada_apply": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "dusty_ada_apply_smem": [_i, _i],
    "dusty_ada_sample": [_vp, _vp, C.c_ulonglong, _vp, _i, _i, _i, _vp, _vp],
    "dusty_split_bf16x3": [_vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong, _vp, _vp, _i, _vp],
    "dusty_conv_role_prof": [_vp, _i],
    "dusty_gemm_tf32": [_vp, _vp, _vp, _i, _i, _i, C.c_longlong, C.c_longlong, C.c_longlong, _f, _i, _vp],
    "dusty_gemm_bf16": [_vp, _vp, _vp, _i, _i, _i, _i, _i, C.c_longlong, C.c_longlong, C.c_longlong, _f, _i, _vp],
    "dusty_gemm_simt": [_vp, _vp, _vp, _i, _i, _i] + [C.c_longlong] * 6 + [_f, _vp],
    "dusty_scan_project": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp],
    "dusty_multi_adam": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _i, _f, _f, _vp],
    "dusty_multi_copy": [_vp, _vp, _vp, _i, _f, _vp],
    "dusty_conv2d_simt": [_i, _vp, _vp, _vp, _vp] + [_i] * 13 + [_vp, _vp, _vp, _f, _i, _vp],
    "dusty_conv2d_wgrad_tc_workspace": [_i] * 7,
    "dusty_conv2d_halo_supported": [_i] * 4,
    "dusty_conv2d_halo_tc": [_vp, _vp, _vp, _vp] + [_i] * 11 + [C.c_longlong] * 4
                            + [_i, _f, _f, C.c_longlong, C.c_longlong, _i, _vp],
    "dusty_conv2d_wgrad_tc": [_vp, _vp, _vp, _vp, C.c_longlong] + [_i] * 11 + [_vp],
}
_RESTYPE = {"dusty_last_error": C.c_char_p, "dusty_launch_count": C.c_int64,
            "dusty_conv2d_wgrad_tc_workspace": C.c_longlong, "dusty_ada_apply_smem": C.c_longlong}

_lib = None
_fns = {}


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dusty_gan_v2_b200/csrc`.  There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    ver = lib.dusty_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libdusty_b200 ABI version {ver} != expected {ABI_VERSION}")
    _lib = lib
    for name in SIGNATURES:
        _fns[name] = getattr(lib, name)
    return lib


def last_error() -> str:
    return (load().dusty_last_error() or b"").decode()


def launch_count() -> int:
    return int(load().dusty_launch_count())


def check(status: int, what: str):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_of(t: torch.Tensor):
    """Raw cudaStream_t of torch's current stream on t's device (fast path: no Stream object)."""
    if _raw_stream is not None:
        return _raw_stream(t.device.index or 0)
    return torch.cuda.current_stream(t.device).cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"dusty_b200 kernels support float32 and bfloat16 tensors, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "dusty_b200 ops run on CUDA tensors only (no CPU fallback; the CPU restatement "
                "lives in oracle/ and is test infrastructure)")


def call(name: str, *args):
    fn = _fns.get(name)
    if fn is None:
        load()
        fn = _fns[name]
    status = fn(*args)
    if status != 0:
        check(status, name)
