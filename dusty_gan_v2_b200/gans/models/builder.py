"""Drop-in for gans/models/builder.py:4-32.  `cfg` may be an OmegaConf node, an attribute
dict or a plain dict."""
from . import dusty_v1, dusty_v2, vanilla


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def _plain(obj):
    if hasattr(obj, "items"):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)) or type(obj).__name__ == "ListConfig":
        return [_plain(v) for v in obj]
    return obj


def build_generator(cfg):
    arch = _get(cfg, "arch")
    if arch == "vanilla":
        return vanilla.Generator(synthesis_kwargs=_plain(_get(cfg, "synthesis_kwargs")))
    if arch == "dusty_v1":
        return dusty_v1.Generator(synthesis_kwargs=_plain(_get(cfg, "synthesis_kwargs")),
                                  measurement_kwargs=_plain(_get(cfg, "measurement_kwargs")))
    if arch == "dusty_v2":
        return dusty_v2.Generator(mapping_kwargs=_plain(_get(cfg, "mapping_kwargs")),
                                  synthesis_kwargs=_plain(_get(cfg, "synthesis_kwargs")),
                                  measurement_kwargs=_plain(_get(cfg, "measurement_kwargs")))
    raise ValueError(arch)


def build_discriminator(cfg):
    arch = _get(cfg, "arch")
    if arch == "vanilla":
        return vanilla.Discriminator(**_plain(_get(cfg, "layer_kwargs")))
    if arch == "dusty_v2":
        return dusty_v2.Discriminator(**_plain(_get(cfg, "layer_kwargs")))
    raise ValueError(arch)
