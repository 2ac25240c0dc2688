"""Logger of the experiment loop — the `MavaLogger` surface of mava/utils/logger.py:91-154 with its console and JSON back-ends.

`MavaLogger(config).log(metrics, t, t_eval, event)` mirrors the reference: TRAIN metrics are reduced to their means, every other
event is `describe`d (mean / std / min / max, mava/utils/logger.py:397-413), `is_terminal_step` is dropped, nested dicts are
flattened with "/" (`log_dict`, :176-181). Back-ends: `ConsoleLogger` (one line per event, :360-394) and `JsonLogger`, which writes
the marl-eval layout `{env: {task: {algorithm: {"seed_<s>": {"step_<k>": {"step_count": t, metric: [v]}, "absolute_metrics":
{...}}}}}}` for the metrics marl-eval plots (`episode_return/mean`, `win_rate`, `steps_per_second`, :300-346). TensorBoard and
Neptune back-ends need packages that are not part of this build and are not provided. Host-side plumbing only (numpy).
"""
from __future__ import annotations

import json
import os
import time
from enum import Enum
from typing import Any, Callable, Dict, List

import numpy as np


class LogEvent(Enum):
    ACT = "actor"
    TRAIN = "trainer"
    EVAL = "evaluator"
    ABSOLUTE = "absolute"
    MISC = "misc"


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def describe(x) -> Dict[str, float] | float:
    """mava/utils/logger.py:397-413: scalars pass through, arrays become mean / std / min / max."""
    x = _np(x)
    if x.ndim == 0 or x.size <= 1:
        return float(x.reshape(-1)[0]) if x.size else float("nan")
    return {"mean": float(np.mean(x)), "std": float(np.std(x)), "min": float(np.min(x)), "max": float(np.max(x))}


def flatten_dict(d: Dict[str, Any], sep: str = "/", prefix: str = "") -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for k, v in d.items():
        key = f"{prefix}{sep}{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten_dict(v, sep, key))
        else:
            out[key] = v
    return out


def winrate_custom_metric(metrics: Dict[str, Any]) -> Dict[str, Any]:
    """mava/utils/logger.py:48-88."""
    if "won_episode" not in metrics:
        return metrics
    n_episodes = int(np.sum(_np(metrics.get("is_terminal_step", np.array([])))))
    if n_episodes == 0:
        return metrics
    metrics["win_rate"] = float(np.sum(_np(metrics["won_episode"]))) / n_episodes * 100
    metrics.pop("won_episode")
    return metrics


class BaseLogger:
    def log_stat(self, key: str, value: float, step: int, eval_step: int, event: LogEvent) -> None:
        raise NotImplementedError

    def log_config(self, config: Dict) -> None:
        return None

    def log_dict(self, data: Dict[str, Any], step: int, eval_step: int, event: LogEvent) -> None:
        for key, value in flatten_dict(data).items():
            self.log_stat(key, value, step, eval_step, event)

    def stop(self) -> None:
        return None


class MultiLogger(BaseLogger):
    def __init__(self, loggers: List[BaseLogger]):
        self.loggers = loggers

    def log_stat(self, key, value, step, eval_step, event):
        for lg in self.loggers:
            lg.log_stat(key, value, step, eval_step, event)

    def log_config(self, config):
        for lg in self.loggers:
            lg.log_config(config)

    def log_dict(self, data, step, eval_step, event):
        for lg in self.loggers:
            lg.log_dict(data, step, eval_step, event)

    def stop(self):
        for lg in self.loggers:
            lg.stop()


class ConsoleLogger(BaseLogger):
    """One line per event: `EVALUATOR - Episode return mean: 0.984 | ...` (mava/utils/logger.py:360-394)."""

    def __init__(self, sink: Callable[[str], None] = print):
        self.sink = sink

    def log_stat(self, key, value, step, eval_step, event):
        self.sink(f"{event.value.upper()} - {self._fmt(key, value)}")

    def log_dict(self, data, step, eval_step, event):
        # only the main metrics reach the console: plain keys and the means of described arrays
        keys = sorted(k for k in flatten_dict(data) if "/" not in k or k.endswith("/mean"))
        flat = flatten_dict(data)
        self.sink(f"{event.value.upper()} - " + " | ".join(self._fmt(k, flat[k]) for k in keys))

    @staticmethod
    def _fmt(key: str, value) -> str:
        name = key.replace("/mean", " mean").replace("_", " ").capitalize()
        v = float(_np(value))
        return f"{name}: {v:.3f}" if abs(v) < 1e6 else f"{name}: {v:.3e}"


class JsonLogger(BaseLogger):
    _METRICS_TO_LOG = ["episode_return/mean", "win_rate", "steps_per_second"]

    def __init__(self, base_exp_path: str, unique_token: str, system_name: str, path: str | None, task_name: str, env_name: str,
                 seed: int):
        logs = os.path.join(base_exp_path, "json", path) if path is not None else os.path.join(base_exp_path, system_name, "json",
                                                                                                 unique_token)
        os.makedirs(logs, exist_ok=True)
        self.file = os.path.join(logs, "metrics.json")
        self.run = (env_name, task_name, system_name, f"seed_{seed}")
        self.data: Dict[str, Any] = {}

    def log_stat(self, key, value, step, eval_step, event):
        if key not in self._METRICS_TO_LOG or event not in (LogEvent.ABSOLUTE, LogEvent.EVAL):
            return
        if "/" in key:  # <metric>/<agg> -> <agg>_<metric>
            key = "_".join(reversed(key.split("/")))
        node = self.data
        for k in self.run:
            node = node.setdefault(k, {})
        if event == LogEvent.ABSOLUTE:
            node.setdefault("absolute_metrics", {})[key] = [float(_np(value))]
        else:
            slot = node.setdefault(f"step_{eval_step}", {"step_count": int(step)})
            slot[key] = [float(_np(value))]
        with open(self.file, "w") as f:
            json.dump(self.data, f, indent=4)


class MavaLogger:
    def __init__(self, config, custom_metrics_fn: Callable[[Dict], Dict] = winrate_custom_metric, console_sink=print):
        lc = config.logger
        unique_token = time.strftime("%Y%m%d%H%M%S")
        system_name = lc.get("system_name", "rec_magpo")
        loggers: List[BaseLogger] = []
        if lc.get("use_console", True):
            loggers.append(ConsoleLogger(console_sink))
        if lc.get("use_json", False):
            kw = lc.get("kwargs", {}) or {}
            loggers.append(JsonLogger(lc.get("base_exp_path", "results"), unique_token, system_name, kw.get("json_path"),
                                      config.env.scenario.task_name, config.env.env_name, int(config.system.seed)))
        for unsupported in ("use_tb", "use_neptune"):
            if lc.get(unsupported, False):
                raise NotImplementedError(f"logger.{unsupported}: back-end not part of this build")
        self.logger = MultiLogger(loggers)
        self.cfg, self.custom_metrics_fn = config, custom_metrics_fn

    def log_config(self, config: Dict | None = None) -> None:
        self.logger.log_config(config if config is not None else self.cfg.to_dict())

    def log(self, metrics: Dict[str, Any], t: int, t_eval: int, event: LogEvent) -> None:
        metrics = self.custom_metrics_fn(dict(metrics))
        metrics.pop("is_terminal_step", None)
        if event == LogEvent.TRAIN:
            metrics = {k: float(np.mean(_np(v))) for k, v in metrics.items()}
        else:
            metrics = {k: describe(v) for k, v in metrics.items()}
        self.logger.log_dict(metrics, t, t_eval, event)

    def stop(self) -> None:
        self.logger.stop()
