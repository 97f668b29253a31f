"""Builds the REFERENCE's own two CUDA extensions (fused bias+activation, upfirdn2d) from the
sources where they lie under /root/reference into oracle/_ref/ -- a second, GPU-side checker
for the two native ops and the kernel-level baseline our kernels are timed against
(tools/ref_kernel_bench.py, tests/test_gpu_ref_kernels.py).

TEST INFRASTRUCTURE ONLY: nothing in the product package imports these modules.  No reference
source is copied: torch.utils.cpp_extension compiles the files in place (the same `load` call
the reference makes, gans/models/ops/fused_act/fused_act.py:10-17 and
gans/models/ops/upfirdn2d/upfirdn2d.py:10-17) and only build products land in oracle/_ref/
(git-ignored, shipped to the GPU box with the snapshot).

    python oracle/build_ref.py          (build container only: needs /root/reference)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_OPS = os.path.join(os.environ.get("DUSTY_REFERENCE_ROOT", "/root/reference"), "gans", "models", "ops")

MODULES = {
    "dusty_ref_fused": [os.path.join(REF_OPS, "fused_act", "fused_bias_act.cpp"),
                        os.path.join(REF_OPS, "fused_act", "fused_bias_act_kernel.cu")],
    "dusty_ref_upfirdn2d": [os.path.join(REF_OPS, "upfirdn2d", "upfirdn2d.cpp"),
                            os.path.join(REF_OPS, "upfirdn2d", "upfirdn2d_kernel.cu")],
}


def build(verbose=False):
    if not all(os.path.exists(f) for fs in MODULES.values() for f in fs):
        print("oracle/build_ref.py: reference sources not found, nothing built")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")      # no GPU in the build container
    from torch.utils.cpp_extension import load
    for name, sources in MODULES.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(name, sources=sources, extra_cuda_cflags=["--use_fast_math"], build_directory=bdir,
             verbose=verbose, is_python_module=True)
        print(f"oracle/_ref/{name}/{name}.so built")
    return True


def load_built(name):
    """Import a module built by build() from oracle/_ref (None when it is not there)."""
    import importlib.util
    path = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if build(verbose="-v" in sys.argv) else 1)
