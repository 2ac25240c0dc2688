"""A small Hydra-compatible config composer (hydra / omegaconf are not available in this image).

Reads the same tree layout as `mava/configs` (`default/rec_magpo.yaml` composing `logger`, `arch`, `system`,
`network`, `env` groups, nested `defaults` such as `env/scenario`) and accepts Hydra-style CLI overrides:
`env=coordsum`, `env/scenario=5x20-80`, `arch.num_envs=4096`, `system.rollout_length=64`.
The result is an attribute-access dict with `OmegaConf.set_struct(cfg, False)` semantics (code may add keys).
"""
from __future__ import annotations

import os

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class Config(dict):
    """dict with attribute access, recursively."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(x):
        if isinstance(x, dict):
            return Config({k: Config.wrap(v) for k, v in x.items()})
        if isinstance(x, list):
            return [Config.wrap(v) for v in x]
        return x

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Config) else v) for k, v in self.items()}


def _load(path):
    with open(path) as f:
        return yaml.safe_load(f) or {}


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _compose_group(config_dir, group, name, choices):
    """Load `<config_dir>/<group>/<name>.yaml`, resolving its own nested defaults (e.g. env -> env/scenario)."""
    node = _load(os.path.join(config_dir, group, f"{name}.yaml"))
    defaults = node.pop("defaults", [])
    out = {}
    for d in defaults:
        if d == "_self_":
            _merge(out, node)
        elif isinstance(d, dict):
            (g, n), = d.items()
            sub = f"{group}/{g}"
            out[g] = _compose_group(config_dir, sub, choices.get(sub, n), choices)
    if "_self_" not in defaults:
        _merge(out, node)
    return out


def compose(config_name: str = "default/rec_magpo", overrides=(), config_dir: str = CONFIG_DIR) -> Config:
    root = _load(os.path.join(config_dir, f"{config_name}.yaml"))
    defaults = root.pop("defaults", [])
    root.pop("hydra", None)
    choices, dotted = {}, []
    for ov in overrides:
        k, _, v = ov.partition("=")
        k = k.lstrip("+")
        if "." in k:
            dotted.append((k, v))
        else:
            choices[k] = v  # group choice: env=coordsum, env/scenario=5x20-80
    cfg = {}
    for d in defaults:
        if d == "_self_":
            _merge(cfg, root)
        elif isinstance(d, dict):
            (g, n), = d.items()
            cfg[g] = _compose_group(config_dir, g, choices.get(g, n), choices)
    for k, v in dotted:
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(v)
    return Config.wrap(cfg)


def check_total_timesteps(config: Config, n_devices: int) -> Config:
    """mava/utils/config.py:47-81: derive num_updates / total_timesteps from one another."""
    sysc, arch = config.system, config.arch
    per_update = n_devices * sysc.rollout_length * sysc.update_batch_size * arch.num_envs
    if sysc.get("total_timesteps") is None:
        sysc.num_updates = int(sysc.num_updates)
        sysc.total_timesteps = int(per_update * sysc.num_updates)
    else:
        sysc.total_timesteps = int(sysc.total_timesteps)
        sysc.num_updates = int(sysc.total_timesteps // per_update)
    assert sysc.num_updates >= arch.num_evaluation, "num_updates must be >= num_evaluation"
    return config
