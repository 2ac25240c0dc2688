"""TEST INFRASTRUCTURE ONLY — CPU restatement of the training wrapper stack around a Jumanji base env.

RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(<Lbf|Rware>Wrapper(env)))) — mava/utils/make_env.py:90-135:
  * JumanjiMarlWrapper.reset/step + modify_timestep      wrappers/jumanji.py:67-98,144-156,192-208
  * AgentIDWrapper._add_agent_ids                        wrappers/observation.py:42-54  (eye(A) first, then the view)
  * AutoResetWrapper.step/_auto_reset/_obs_in_extras     wrappers/auto_reset_wrapper.py:52-101 (key, _ = split(state.key))
  * RecordEpisodeMetrics.reset/step                      wrappers/episode_metrics.py:60-112

A base env is a module-like object with vectorised (leading axis B) functions
    base_reset(spec, keys[B,2]) -> base state dict (must hold `key` uint32[B,2] and `step_count` int32[B])
    base_step(spec, base, actions[B,A]) -> (new base, reward f32[B,A] after the Mava wrapper, terminate bool[B], truncate bool[B])
    observe(spec, base) -> (agents_view f32[B,A,d0], action_mask bool[B,A,a])
The state / timestep dictionaries have the same structure as oracle/coordsum.py so oracle/learner.py drives either.
"""
import numpy as np

from . import prng

STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2


def _observation(spec, envmod, base):
    view, mask = envmod.observe(spec, base)
    B, A = view.shape[0], view.shape[1]
    ids = np.broadcast_to(np.eye(A, dtype=np.float32)[None], (B, A, A))
    return dict(
        agents_view=np.concatenate([ids, view.astype(np.float32)], axis=-1),
        action_mask=mask.astype(bool),
        step_count=np.repeat(base["step_count"][:, None], A, axis=1).astype(np.int32),
    )


def reset(spec, envmod, keys):
    keys = np.asarray(keys, np.uint32)
    B, A = keys.shape[0], spec.num_agents
    ks = prng.split_batched(keys, 2)  # RecordEpisodeMetrics.reset: key, reset_key = split(key)
    key, reset_key = ks[:, 0], ks[:, 1]
    base = envmod.base_reset(spec, reset_key)
    obs = _observation(spec, envmod, base)
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


def step(spec, envmod, state, actions):
    actions = np.asarray(actions, np.int32)
    B, A = actions.shape
    new_base, rewards, terminate, truncate = envmod.base_step(spec, state["env_state"], actions)
    done = terminate | truncate
    obs = _observation(spec, envmod, new_base)
    real_next_obs = {k: v.copy() for k, v in obs.items()}
    step_type = np.where(done, STEP_LAST, STEP_MID).astype(np.int8)
    discount = np.repeat(np.where(terminate, 0.0, 1.0).astype(np.float32)[:, None], A, axis=1)  # truncation keeps discount 1
    if done.any():  # AutoResetWrapper._auto_reset
        di = np.nonzero(done)[0]
        rkeys = prng.split_batched(new_base["key"][di], 2)[:, 0]
        rb = envmod.base_reset(spec, rkeys)
        new_base = {k: v.copy() for k, v in new_base.items()}
        for k in new_base:
            new_base[k][di] = rb[k]
        robs = _observation(spec, envmod, rb)
        for k in obs:
            obs[k][di] = robs[k]
    # RecordEpisodeMetrics.step: jnp.mean(reward) = sum / A in fp32
    mean_r = rewards[:, 0].copy()
    for i in range(1, A):
        mean_r = (mean_r + rewards[:, i]).astype(np.float32)
    mean_r = (mean_r / np.float32(A)).astype(np.float32)
    d = done.astype(np.int32)
    nd = 1 - d
    new_ret = (state["running_count_episode_return"] + mean_r).astype(np.float32)
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
