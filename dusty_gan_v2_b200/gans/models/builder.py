"""Drop-in for gans/models/builder.py (reference 4-32): `build_generator(cfg)` /
`build_discriminator(cfg)` keyed on `cfg.arch`.  `cfg` may be an OmegaConf node, the package's
attribute dict or a plain dict; keyword groups are converted to plain containers before they reach
the module constructors (which store them).
"""
from . import dusty_v1, dusty_v2, vanilla

# arch -> (constructor, keyword groups read from the config node)
_GENERATORS = {
    "vanilla": (vanilla.Generator, ("synthesis_kwargs",)),
    "dusty_v1": (dusty_v1.Generator, ("synthesis_kwargs", "measurement_kwargs")),
    "dusty_v2": (dusty_v2.Generator, ("mapping_kwargs", "synthesis_kwargs", "measurement_kwargs")),
}
_DISCRIMINATORS = {"vanilla": vanilla.Discriminator, "dusty_v2": dusty_v2.Discriminator}


def _field(node, name):
    return node[name] if isinstance(node, dict) else getattr(node, name)


def _to_builtin(node):
    """OmegaConf / attribute-dict trees -> dicts and lists of plain values."""
    if hasattr(node, "items"):
        return {key: _to_builtin(val) for key, val in node.items()}
    if isinstance(node, (list, tuple)) or type(node).__name__ == "ListConfig":
        return [_to_builtin(val) for val in node]
    return node


def build_generator(cfg):
    arch = _field(cfg, "arch")
    if arch not in _GENERATORS:
        raise ValueError(arch)
    ctor, groups = _GENERATORS[arch]
    return ctor(**{group: _to_builtin(_field(cfg, group)) for group in groups})


def build_discriminator(cfg):
    arch = _field(cfg, "arch")
    if arch not in _DISCRIMINATORS:
        raise ValueError(arch)
    return _DISCRIMINATORS[arch](**_to_builtin(_field(cfg, "layer_kwargs")))
