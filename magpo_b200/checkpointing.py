"""Checkpointer of the learner state — the surface of mava/utils/checkpointing.py:34-215 (`Checkpointer(model_name, metadata,
rel_dir, checkpoint_uid, save_interval_steps, max_to_keep, keep_period).save(timestep, unreplicated_learner_state, episode_return)`,
`.restore_params(input_params, timestep, restore_hstates, THiddenState)`, `.get_cfg()`), used at rec_magpo.py:732-741,779-786.

The reference stores an orbax PyTreeCheckpointer tree; orbax is not part of this build, so the documented alternative layout is
`<cwd>/<rel_dir>/<model_name>/<uid>/<step>/learner_state.npz` (one array per pytree leaf, keyed by its "/"-joined path, the same
paths orbax uses for its per-leaf directories) next to `metadata.json` (`checkpointer_version` + the config). Retention follows the
orbax options of the reference: `best_fn = episode_return`, `best_mode = max`, `max_to_keep`, `keep_period`,
`save_interval_steps`. Host-side plumbing only.
"""
from __future__ import annotations

import json
import os
import shutil
import time
from typing import Any, Dict, Optional, Tuple

import numpy as np

CHECKPOINTER_VERSION = 1.0


def flatten_pytree(tree: Any, prefix: str = "") -> Dict[str, np.ndarray]:
    """NamedTuple / dict / tensor pytree -> {"a/b/c": ndarray}; field names follow systems/gpo/types.py."""
    out: Dict[str, np.ndarray] = {}
    if tree is None:  # None is an empty subtree, as in a JAX pytree
        return out
    if hasattr(tree, "_fields"):
        items = [(f, getattr(tree, f)) for f in tree._fields]
    elif isinstance(tree, dict):
        items = list(tree.items())
    elif isinstance(tree, (list, tuple)):
        items = [(str(i), v) for i, v in enumerate(tree)]
    else:
        leaf = tree.detach().cpu().numpy() if hasattr(tree, "detach") else np.asarray(tree)
        return {prefix: leaf}
    for k, v in items:
        out.update(flatten_pytree(v, f"{prefix}/{k}" if prefix else str(k)))
    return out


def unreplicate_n_dims(tree: Any, n: int = 2) -> Any:
    """mava/utils/jax_utils.py `unreplicate_n_dims`: index the first `n` (device, update-batch) axes at 0 on every leaf."""
    if tree is None:
        return None
    if hasattr(tree, "_fields"):
        return type(tree)(*[unreplicate_n_dims(getattr(tree, f), n) for f in tree._fields])
    if isinstance(tree, dict):
        return {k: unreplicate_n_dims(v, n) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(unreplicate_n_dims(v, n) for v in tree)
    x = tree
    for _ in range(n):
        x = x[0]
    return x


class Checkpointer:
    def __init__(self, model_name: str, metadata: Optional[Dict] = None, rel_dir: str = "checkpoints",
                 checkpoint_uid: Optional[str] = None, save_interval_steps: int = 1, max_to_keep: Optional[int] = 1,
                 keep_period: Optional[int] = None):
        uid = checkpoint_uid if checkpoint_uid else time.strftime("%Y%m%d%H%M%S")
        self.directory = os.path.join(os.getcwd(), rel_dir, model_name, uid)
        os.makedirs(self.directory, exist_ok=True)
        self.save_interval_steps, self.max_to_keep, self.keep_period = max(1, int(save_interval_steps)), max_to_keep, keep_period
        meta_path = os.path.join(self.directory, "metadata.json")
        if metadata is not None or not os.path.exists(meta_path):
            md = metadata.to_dict() if hasattr(metadata, "to_dict") else dict(metadata or {})
            with open(meta_path, "w") as f:
                json.dump({"checkpointer_version": CHECKPOINTER_VERSION, **_json_ready(md)}, f, indent=1)
        self._index_path = os.path.join(self.directory, "index.json")
        self._index = json.load(open(self._index_path)) if os.path.exists(self._index_path) else {}
        self._last_saved: Optional[int] = None

    # ---------------------------------------------------------------- save
    def save(self, timestep: int, unreplicated_learner_state: Any, episode_return: float = 0.0) -> bool:
        timestep = int(timestep)
        if self._last_saved is not None and timestep - self._last_saved < self.save_interval_steps and self.save_interval_steps > 1:
            return False
        d = os.path.join(self.directory, str(timestep))
        os.makedirs(d, exist_ok=True)
        np.savez(os.path.join(d, "learner_state.npz"), **flatten_pytree(unreplicated_learner_state))
        self._index[str(timestep)] = {"episode_return": float(episode_return)}
        self._last_saved = timestep
        self._prune()
        with open(self._index_path, "w") as f:
            json.dump(self._index, f, indent=1)
        return True

    def _prune(self) -> None:
        if self.max_to_keep is None:
            return
        protected = {s for s in self._index if self.keep_period and int(s) % int(self.keep_period) == 0}
        cand = sorted((s for s in self._index if s not in protected), key=lambda s: (self._index[s]["episode_return"], int(s)))
        while len(cand) > self.max_to_keep:  # best_fn = episode_return, best_mode = max: drop the worst (oldest on ties)
            s = cand.pop(0)
            shutil.rmtree(os.path.join(self.directory, s), ignore_errors=True)
            del self._index[s]

    # ---------------------------------------------------------------- restore
    def latest_step(self) -> Optional[int]:
        return max((int(s) for s in self._index), default=None)

    def restore(self, timestep: Optional[int] = None) -> Dict[str, np.ndarray]:
        step = timestep if timestep else self.latest_step()
        if step is None:
            raise FileNotFoundError(f"no checkpoint under {self.directory}")
        with np.load(os.path.join(self.directory, str(step), "learner_state.npz")) as z:
            return {k: z[k] for k in z.files}

    def restore_params(self, input_params: Any, timestep: Optional[int] = None, restore_hstates: bool = False,
                       THiddenState: Optional[type] = None) -> Tuple[Any, Optional[Any]]:  # noqa: N803
        assert int(self.get_cfg()["checkpointer_version"]) == int(CHECKPOINTER_VERSION), \
            "Loaded checkpoint was created with a different major version of the checkpointer."
        flat = self.restore(timestep)

        def rebuild(template: Any, prefix: str) -> Any:
            if hasattr(template, "_fields"):
                return type(template)(*[rebuild(getattr(template, f), f"{prefix}/{f}") for f in template._fields])
            if isinstance(template, dict):
                return {k: rebuild(v, f"{prefix}/{k}") for k, v in template.items()}
            return flat[prefix]

        params = rebuild(input_params, "params")
        hstates = None
        if restore_hstates and THiddenState is not None:
            names = THiddenState._fields
            sub = {k[len("hstates/"):]: v for k, v in flat.items() if k.startswith("hstates/")}
            hstates = THiddenState(*[_nest({k[len(n) + 1:]: v for k, v in sub.items() if k.startswith(n + "/")} or sub.get(n))
                                     for n in names])
        return params, hstates

    def get_cfg(self) -> Dict:
        with open(os.path.join(self.directory, "metadata.json")) as f:
            return json.load(f)


def _nest(flat):
    if not isinstance(flat, dict):
        return flat
    out: Dict[str, Any] = {}
    for k, v in flat.items():
        node = out
        parts = k.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return out


def _json_ready(obj: Any) -> Any:
    if isinstance(obj, dict):
        return {str(k): _json_ready(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_json_ready(v) for v in obj]
    return obj if isinstance(obj, (bool, str, int, float, type(None))) else str(obj)
