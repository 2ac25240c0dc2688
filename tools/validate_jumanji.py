#!/usr/bin/env python
"""Validator for the RWARE / LBF dynamics against REAL Jumanji.

`oracle/lbf.py` and `oracle/rware.py` restate jumanji@9ced6b8 (v1.1.0, uv.lock:1217-1219) from memory, because the package is
neither vendored in the reference nor installable in the build image; the CUDA env kernels are bit-exact against that restatement,
and the committed traces `tests/golden/{lbf,rware}_trace.npz` are what both reproduce. This script closes the loop on a machine
that HAS jax + jumanji + the reference checkout: it builds the reference's own training env

    RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(<Lbf|Rware>Wrapper(jumanji.make(name, generator=RandomGenerator(**task_config), **env_kwargs)))))

exactly as `mava/utils/make_env.py:90-135` does, resets it with the trace's keys (`jax.vmap(env.reset)(keys)`), feeds the trace's
action stream through `jax.vmap(env.step)` and compares, step by step, the observation (`agents_view`, `action_mask`), reward,
step type and `extras["episode_metrics"]["episode_return"]` with the trace, then the final env PRNG key. It exits non-zero on the
first mismatch and prints which field, step and env differ. Nothing in tests/, smoke() or bench.py imports this file.

    python tools/validate_jumanji.py --mava /path/to/liyheng-MAGPO [--env lbf|rware|both]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# scenario -> (jumanji registration name, RandomGenerator task_config, env kwargs) as in mava/configs/env/{lbf,rware}.yaml and
# mava/configs/env/scenario/{2s-8x8-2p-2f-coop,tiny-4ag}.yaml; the RWARE trace was generated with time_limit=70 (make_golden.py)
CASES = {
    "lbf": dict(env_name="LevelBasedForaging", name="LevelBasedForaging-v0", time_limit=100,
                task_config=dict(grid_size=8, fov=2, num_agents=2, num_food=2, max_agent_level=2, force_coop=True)),
    "rware": dict(env_name="RobotWarehouse", name="RobotWarehouse-v0", time_limit=70,
                  task_config=dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1, request_queue_size=4)),
}


def fail(msg: str) -> None:
    print("MISMATCH:", msg)
    sys.exit(1)


def first_diff(a: np.ndarray, b: np.ndarray):
    idx = np.argwhere(np.asarray(a) != np.asarray(b))
    return tuple(int(i) for i in idx[0]) if len(idx) else None


def validate(env_key: str, mava_root: str) -> None:
    try:
        import jax
        import jax.numpy as jnp
        import jumanji  # noqa: F401
    except ImportError as e:
        print(f"jax / jumanji are required ({e}); this validator cannot run in the build image (no network)")
        sys.exit(2)
    sys.path.insert(0, mava_root)
    from omegaconf import OmegaConf

    from mava.utils import make_env as mk

    case = CASES[env_key]
    cfg = OmegaConf.create({"env": {"env_name": case["env_name"], "kwargs": {"time_limit": case["time_limit"]},
                                    "scenario": {"name": case["name"], "task_config": case["task_config"], "env_kwargs": {}}},
                            "system": {"add_agent_id": True}})
    env, _ = mk.make_jumanji_env(cfg)  # RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(Wrapper(jumanji.make(...)))))
    tr = np.load(os.path.join(GOLDEN, f"{env_key}_trace.npz"))
    keys = jnp.asarray(tr["keys"], dtype=jnp.uint32)
    state, ts = jax.vmap(env.reset)(keys)
    step = jax.jit(jax.vmap(env.step))

    def check(tag, t):
        view = np.asarray(ts.observation.agents_view, np.float32)
        d = first_diff(view, tr["agents_view"][t])
        if d is not None:
            fail(f"{env_key} {tag}: agents_view[env {d[0]}, agent {d[1]}, feature {d[2]}] = {view[d]} vs trace {tr['agents_view'][t][d]}")
        mask = np.asarray(ts.observation.action_mask, bool)
        d = first_diff(mask, tr["action_mask"][t].astype(bool))
        if d is not None:
            fail(f"{env_key} {tag}: action_mask[env {d[0]}, agent {d[1]}, action {d[2]}] = {mask[d]} vs trace {bool(tr['action_mask'][t][d])}")

    check("reset", 0)
    n_steps = tr["actions"].shape[0]
    for t in range(n_steps):
        state, ts = step(state, jnp.asarray(tr["actions"][t]))
        check(f"step {t}", t + 1)
        for name, got, ref in (("reward", np.asarray(ts.reward, np.float32), tr["rewards"][t]),
                               ("step_type", np.asarray(ts.step_type, np.int8), tr["step_type"][t].astype(np.int8)),
                               ("episode_return", np.asarray(ts.extras["episode_metrics"]["episode_return"], np.float32), tr["episode_return"][t])):
            d = first_diff(got, ref)
            if d is not None:
                fail(f"{env_key} step {t}: {name}{list(d)} = {got[d]} vs trace {ref[d]}")
    final_key = np.asarray(jax.random.key_data(state.env_state.key) if hasattr(jax.random, "key_data") else state.env_state.key, np.uint32)
    d = first_diff(final_key, tr["final_key"])
    if d is not None:
        fail(f"{env_key}: final env PRNG key differs at {d} (the generator's key-splitting order is one of the restatement's known unknowns)")
    print(f"{env_key}: {n_steps} steps x {keys.shape[0]} envs identical to tests/golden/{env_key}_trace.npz — "
          f"the restatement (and with it the CUDA kernel, which is bit-exact against it) matches Jumanji on this trace")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mava", default="/root/reference", help="checkout of liyheng/MAGPO (provides mava.utils.make_env)")
    ap.add_argument("--env", default="both", choices=["lbf", "rware", "both"])
    a = ap.parse_args()
    for e in (("lbf", "rware") if a.env == "both" else (a.env,)):
        validate(e, a.mava)
