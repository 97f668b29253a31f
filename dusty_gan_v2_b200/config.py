"""Minimal attribute-dict config loader (the reference uses OmegaConf, which is not a
dependency here): `cfg.training.lazy.gp` style access over the reference's YAML files."""
import copy

import yaml


class AttrDict(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def load_config(path: str) -> AttrDict:
    with open(path) as f:
        return to_attr(yaml.safe_load(f))
