"""CoordSum + the Mava training wrapper stack, vectorised over envs in NumPy — oracle only.

Restates, for a batch of B independent envs (the reference vmaps a scalar env):
  * `mava/coordsum/env.py:55-139`            CoordSum.reset / CoordSum.step
  * `mava/coordsum/__init__.py:6-45`         the four registered scenarios
  * `mava/wrappers/matrax.py:104-144`        CoordSumWrapper.modify_timestep
  * `mava/wrappers/observation.py:42-72`     AgentIDWrapper (int32 one-hot prefix)
  * `mava/wrappers/auto_reset_wrapper.py:52-101` AutoResetWrapper
  * `mava/wrappers/episode_metrics.py:60-112`    RecordEpisodeMetrics
in the order `mava/utils/make_env.py:90-104` stacks them:
RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(CoordSumWrapper(CoordSum)))).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import prng

SCENARIOS = {  # coordsum/__init__.py:6-45
    "5x20-80-v0": dict(num_agents=5, num_actions=20, time_limit=100, maxval=80),
    "3x30-50-v0": dict(num_agents=3, num_actions=30, time_limit=100, maxval=50),
    "3x10-30-v0": dict(num_agents=3, num_actions=10, time_limit=100, maxval=30),
    "8x15-100-v0": dict(num_agents=8, num_actions=15, time_limit=100, maxval=100),
}

STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2  # jumanji StepType (Appendix A12)


@dataclass
class CoordSumSpec:
    num_agents: int
    num_actions: int
    time_limit: int = 100
    maxval: int | None = None

    def __post_init__(self):
        if not self.maxval:  # coordsum/env.py:48-52
            self.maxval = self.num_actions

    @property
    def obs_dim(self):  # agent-id one-hot + the 1-wide target view
        return self.num_agents + 1

    @property
    def action_dim(self):
        return self.num_actions


def _base_reset(spec: CoordSumSpec, keys):
    """CoordSum.reset vmapped (coordsum/env.py:55-74). keys uint32[B,2]."""
    B = keys.shape[0]
    ks = prng.split_batched(keys, 2)
    key, target_key = ks[:, 0], ks[:, 1]
    target = prng.randint_batched(target_key, spec.time_limit + 1, 0, spec.maxval)
    state = dict(
        step_count=np.zeros(B, np.int32),
        target=target,
        record=-np.ones((B, spec.num_actions, spec.time_limit), np.int32),
        key=key.copy(),
    )
    return state


def _observation(spec: CoordSumSpec, target_val, step):
    """CoordSumWrapper.modify_timestep + AgentIDWrapper (matrax.py:117-134, observation.py:42-54)."""
    B, A = target_val.shape[0], spec.num_agents
    view = np.zeros((B, A, A + 1), np.int32)
    view[:, np.arange(A), np.arange(A)] = 1
    view[:, :, A] = target_val[:, None]
    return dict(
        agents_view=view,
        action_mask=np.ones((B, A, spec.num_actions), bool),
        step_count=np.repeat(step[:, None], A, axis=1).astype(np.int32),
    )


def reset(spec: CoordSumSpec, keys):
    """RecordEpisodeMetrics.reset over the whole stack (episode_metrics.py:60-77)."""
    keys = np.asarray(keys, np.uint32)
    B, A = keys.shape[0], spec.num_agents
    ks = prng.split_batched(keys, 2)
    key, reset_key = ks[:, 0], ks[:, 1]
    base = _base_reset(spec, reset_key)
    obs = _observation(spec, base["target"][:, 0], base["step_count"])
    state = dict(
        env_state=base,
        key=key.copy(),
        running_count_episode_return=np.zeros(B, np.float32),
        running_count_episode_length=np.zeros(B, np.int32),
        episode_return=np.zeros(B, np.float32),
        episode_length=np.zeros(B, np.int32),
    )
    timestep = dict(
        step_type=np.full(B, STEP_FIRST, np.int8),
        reward=np.zeros((B, A), np.float32),
        discount=np.ones((B, A), np.float32),
        observation=obs,
        extras=dict(
            real_next_obs={k: v.copy() for k, v in obs.items()},
            episode_metrics=dict(
                episode_return=np.zeros(B, np.float32),
                episode_length=np.zeros(B, np.int32),
                is_terminal_step=np.zeros(B, bool),
            ),
        ),
    )
    return state, timestep


def step(spec: CoordSumSpec, state, actions):
    """One vmapped env.step through the wrapper stack. actions int32[B,A]. Returns new (state, timestep)."""
    base = state["env_state"]
    B, A, a, TL = actions.shape[0], spec.num_agents, spec.num_actions, spec.time_limit
    ar = np.arange(B)
    sc = base["step_count"]
    # --- CoordSum.step (coordsum/env.py:76-139)
    target_t = base["target"][ar, sc]
    sum_match = actions.sum(axis=1) == target_t
    row_idx = np.minimum(target_t, a - 1)  # JAX gather clamps OOB reads (Appendix A6)
    row = base["record"][ar, row_idx]  # [B, TL]
    valid = row != -1
    counts = np.zeros((B, TL), np.float32)
    bb, tt = np.nonzero(valid)
    np.add.at(counts, (bb, row[bb, tt]), 1.0)  # bincount(length=time_limit) with 0/1 weights
    guess = counts.argmax(axis=1)  # first max; 0 if empty
    hit = guess == actions[:, 0]
    reward = np.where(sum_match, np.where(hit, 1.0, 2.0), 0.0).astype(np.float32)
    rewards = np.repeat(reward[:, None], A, axis=1)
    record = base["record"].copy()
    record[ar, row_idx, np.minimum(sc, TL - 1)] = actions[:, 0]  # dynamic_update_slice clamps the start
    steps = sc + 1
    done = steps >= TL
    next_target = base["target"][ar, np.minimum(steps, TL)]
    new_base = dict(step_count=steps.astype(np.int32), target=base["target"], record=record, key=base["key"])
    obs = _observation(spec, next_target, new_base["step_count"])
    real_next_obs = {k: v.copy() for k, v in obs.items()}
    step_type = np.where(done, STEP_LAST, STEP_MID).astype(np.int8)
    discount = np.where(done[:, None], 0.0, 1.0).astype(np.float32) * np.ones((1, A), np.float32)
    # --- AutoResetWrapper._auto_reset (auto_reset_wrapper.py:60-83): key, _ = split(state.key)
    if done.any():
        di = np.nonzero(done)[0]
        rkeys = prng.split_batched(new_base["key"][di], 2)[:, 0]
        rb = _base_reset(spec, rkeys)
        new_base = {k: v.copy() for k, v in new_base.items()}
        for k in ("step_count", "target", "record", "key"):
            new_base[k][di] = rb[k]
        robs = _observation(spec, rb["target"][:, 0], rb["step_count"])
        for k in obs:
            obs[k][di] = robs[k]
    # --- RecordEpisodeMetrics.step (episode_metrics.py:79-112)
    d = done.astype(np.int32)
    nd = 1 - d
    new_ret = (state["running_count_episode_return"] + rewards.mean(axis=1, dtype=np.float32)).astype(np.float32)
    new_len = state["running_count_episode_length"] + 1
    ep_ret = (state["episode_return"] * nd + new_ret * d).astype(np.float32)
    ep_len = (state["episode_length"] * nd + new_len * d).astype(np.int32)
    new_state = dict(
        env_state=new_base,
        key=state["key"],
        running_count_episode_return=(new_ret * nd).astype(np.float32),
        running_count_episode_length=(new_len * nd).astype(np.int32),
        episode_return=ep_ret,
        episode_length=ep_len,
    )
    timestep = dict(
        step_type=step_type,
        reward=rewards,
        discount=discount,
        observation=obs,
        extras=dict(
            real_next_obs=real_next_obs,
            episode_metrics=dict(episode_return=ep_ret, episode_length=ep_len, is_terminal_step=done.copy()),
        ),
    )
    return new_state, timestep
