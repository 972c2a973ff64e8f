"""Import shim (test infrastructure only) for `omegaconf`: YAML -> attribute dict.
No arithmetic."""
import yaml


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ListConfig(list):
    pass


def _wrap(o):
    if isinstance(o, dict):
        return DictConfig({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return ListConfig(_wrap(v) for v in o)
    return o


def _unwrap(o):
    if isinstance(o, dict):
        return {k: _unwrap(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_unwrap(v) for v in o]
    return o


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path, "r") as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)

    @staticmethod
    def create(obj):
        return _wrap(obj)
