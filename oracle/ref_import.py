"""Import shim for the *reference* (kazuto1011/dusty-gan-v2) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (and by optional
cross-check tests that skip when /root/reference is absent) to import the
reference's own Python modules from /root/reference so that golden vectors can be
generated from the real thing.  Never imported by the product package.

The reference imports visualisation / config packages that are not installed here
(kornia, imageio, matplotlib, seaborn, omegaconf, polyscope).  None of their
symbols is touched by the hot path, so they are replaced
by inert placeholder modules.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DUSTY_REFERENCE_ROOT", "/root/reference")
_STUB_ROOTS = ("kornia", "imageio", "matplotlib", "seaborn", "omegaconf", "polyscope")


class _Inert(types.ModuleType):
    """Placeholder module: any attribute resolves to an inert callable/class."""

    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None,
                               "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


class _InertFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            try:  # prefer a real install when there is one
                sys.meta_path.remove(self)
                real = importlib.util.find_spec(fullname) if "." not in fullname else None
            except Exception:
                real = None
            finally:
                sys.meta_path.insert(0, self)
            if real is not None:
                return None
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Inert(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gans"))


def install():
    """Make `import gans...` resolve to the reference tree (CPU path only)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util  # noqa: F401

    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/torch_ext")
    os.makedirs(os.environ["TORCH_EXTENSIONS_DIR"], exist_ok=True)
    sys.meta_path.insert(0, _InertFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class AttrDict(dict):
    """dict with attribute access, standing in for OmegaConf nodes."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_attr(v) for v in obj]
    return obj
