"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference evaluator for the learner policy.

Follows mava/evaluator.py:82-163 (`get_eval_fn`: `episode_loops` x [`key, reset_key = split(key)`; reset `n` envs with
`split(reset_key, n)`; `time_limit + 1` steps of `key, act_key = split(key)`, act, `env.step`; metrics picked at the first
`last()`]) and :188-208 (`make_rec_eval_act_fn`: the recurrent actor applied with `done = timestep.last()`; `pi.mode()` if
`evaluation_greedy` else `pi.sample(seed=act_key)`).

Parity unpinned: `pi` is `IdentityTransformation(tfd.Categorical(logits))` (networks/heads.py:63); tensorflow-probability 0.25.0
is not available here, so its JAX-substrate sampler is restated from memory: `z = jax.random.gumbel(seed, logits_2d.shape +
(num_samples,))`, `argmax(logits_2d[..., None] + z, axis=-2)` with `logits_2d = logits.reshape(-1, a)` — i.e. noise element
(row, j) uses counter `row * a + j`, rows ordered (env, agent). The eval env has no AutoResetWrapper (utils/make_env.py:90-104);
up to and including the first terminal step the auto-resetting stack used here is identical to it.
"""
import numpy as np
import torch

from . import coordsum as ocs
from . import nets, prng
from .learner import env_module


def eval_act(ap_t, ncfg, timestep, act_key, hidden, greedy):
    """make_rec_eval_act_fn (evaluator.py:188-208)."""
    ob = timestep["observation"]
    A = ncfg.n_agents
    last = timestep["step_type"] == ocs.STEP_LAST
    last_done = np.repeat(last[:, None], A, axis=1)
    obs = torch.tensor(ob["agents_view"].astype(np.float32))[None]
    mask = torch.tensor(ob["action_mask"])[None]
    with torch.no_grad():
        h, logits = nets.actor_apply(ap_t, ncfg, torch.tensor(hidden), obs, torch.tensor(last_done)[None], mask)
    lg = logits[0].numpy()  # [E, A, a]
    if greedy:
        action = lg.argmax(-1)
    else:
        z = prng.gumbel(act_key, (lg.shape[0] * lg.shape[1] * lg.shape[2],)).reshape(lg.shape)
        action = (lg + z).argmax(-1)
    return action.astype(np.int32), h.numpy()


def eval_fn(spec, ncfg, actor_params, key, n_envs, episode_loops, greedy=False):
    """Returns dict(episode_return f32[episode_loops * n_envs], episode_length i32[...])."""
    ap_t = nets.to_torch(actor_params)
    rets, lens = [], []
    key = np.asarray(key, np.uint32)
    for _ in range(episode_loops):
        key, reset_key = prng.split(key)
        state, ts = env_module(spec).reset(spec, prng.split(reset_key, n_envs))
        hidden = np.zeros((n_envs, ncfg.n_agents, ncfg.hidden), np.float32)
        lasts, ers, els = [], [], []
        step_key = key  # mava/evaluator.py:139-146: the step scan's final key is discarded; `_episode` returns the post-reset key
        for _ in range(spec.time_limit + 1):
            ks = prng.split(step_key)
            step_key, act_key = ks[0], ks[1]
            action, hidden = eval_act(ap_t, ncfg, ts, act_key, hidden, greedy)
            state, ts = env_module(spec).step(spec, state, action)
            lasts.append(ts["step_type"] == ocs.STEP_LAST)
            ers.append(ts["extras"]["episode_metrics"]["episode_return"].copy())
            els.append(ts["extras"]["episode_metrics"]["episode_length"].copy())
        done_idx = np.stack(lasts).argmax(axis=0)  # first last() per env
        ar = np.arange(n_envs)
        rets.append(np.stack(ers)[done_idx, ar])
        lens.append(np.stack(els)[done_idx, ar])
    return dict(episode_return=np.concatenate(rets).astype(np.float32), episode_length=np.concatenate(lens).astype(np.int32))


def eval_fn_sable(spec, ncfg, guider_params, key, n_envs, episode_loops):
    """get_eval_fn with rec_sable's act function (rec_sable.py:497-516): the Sable network acts and carries its retention states."""
    gp_t = nets.to_torch(guider_params)
    rets, lens = [], []
    key = np.asarray(key, np.uint32)
    hs_shape = (n_envs, ncfg.n_head, ncfg.n_block, ncfg.head_size, ncfg.head_size)
    for _ in range(episode_loops):
        key, reset_key = prng.split(key)
        state, ts = env_module(spec).reset(spec, prng.split(reset_key, n_envs))
        hs = tuple(torch.zeros(hs_shape) for _ in range(3))
        lasts, ers, els = [], [], []
        step_key = key  # mava/evaluator.py:139-146: the step scan's final key is discarded; `_episode` returns the post-reset key
        for _ in range(spec.time_limit + 1):
            ks = prng.split(step_key)
            step_key, act_key = ks[0], ks[1]
            ob = ts["observation"]
            with torch.no_grad():
                action, _, _, hs = nets.sable_get_actions(gp_t, ncfg, torch.tensor(ob["agents_view"].astype(np.float32)),
                                                          torch.tensor(ob["action_mask"]), torch.tensor(ob["step_count"]), hs, act_key)
            state, ts = env_module(spec).step(spec, state, np.asarray(action, np.int32))
            lasts.append(ts["step_type"] == ocs.STEP_LAST)
            ers.append(ts["extras"]["episode_metrics"]["episode_return"].copy())
            els.append(ts["extras"]["episode_metrics"]["episode_length"].copy())
        done_idx = np.stack(lasts).argmax(axis=0)
        ar = np.arange(n_envs)
        rets.append(np.stack(ers)[done_idx, ar])
        lens.append(np.stack(els)[done_idx, ar])
    return dict(episode_return=np.concatenate(rets).astype(np.float32), episode_length=np.concatenate(lens).astype(np.int32))
