"""YAML config loader with the subset of the OmegaConf surface the reference uses
(`OmegaConf.load`, attribute + `.get` access, `OmegaConf.to_container`); see the
schema in reference src/model/sort/*/train_cf_*.yaml and base_model.py:69-106."""
from __future__ import annotations

import os

import yaml


class Config(dict):
    """dict with attribute access (cfg.train_hparams.lr), like a DictConfig."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(o):
    if isinstance(o, dict):
        return Config({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return [_wrap(v) for v in o]
    return o


def to_container(o):
    if isinstance(o, dict):
        return {k: to_container(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [to_container(v) for v in o]
    return o


def load_config(path_or_dict) -> Config:
    if isinstance(path_or_dict, dict):
        return _wrap(path_or_dict)
    if not os.path.exists(path_or_dict):
        raise FileNotFoundError(f"Config file not found: {path_or_dict}")  # base_model.py:71-72
    with open(path_or_dict, "r") as f:
        return _wrap(yaml.safe_load(f) or {})
