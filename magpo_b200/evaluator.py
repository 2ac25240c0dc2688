"""Evaluator of the learner policy on the GPU (mava/evaluator.py:66-163,188-208; rec_magpo.py:706-710,770-777).

`get_eval_fn(env, actor_network, config, absolute_metric)` returns `eval_fn(actor_params_flat, key) -> metrics` with one entry
per evaluation episode (`episode_return`, `episode_length`, plus `steps_per_second`), exactly the structure the reference's
`timed_eval_fn` returns for one device. Per episode loop: `key, reset_key = split(key)`; `n` envs reset from
`split(reset_key, n)`; `time_limit + 1` steps of `key, act_key = split(key)`; the recurrent actor applied with
`done = timestep.last()`; `pi.mode()` when `arch.evaluation_greedy` else a gumbel-max sample from `act_key` (noise element
(env, agent, j) uses counter `(env * A + agent) * a + j`, the tfp sampler layout); `env.step`; metrics at the first `last()`.
The eval env of the reference has no AutoResetWrapper; up to the first terminal step the training stack used here (the fused
env-step kernel) is the same env. This is evaluation plumbing, not the hot path: torch ops are used for the argmax / bookkeeping.
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np
import torch

from . import _lib as L
from . import init as minit
from .learner import alloc_timestep


def get_num_eval_envs(config, absolute_metric: bool, n_devices: int = 1) -> int:
    """evaluator.py:66-80."""
    n_parallel_envs = config.arch.num_envs * n_devices
    eval_episodes = config.arch.num_absolute_metric_eval_episodes if absolute_metric else config.arch.num_eval_episodes
    return math.ceil(eval_episodes / n_devices) if eval_episodes <= n_parallel_envs else int(config.arch.num_envs)


def get_eval_fn(env, actor_network, config, absolute_metric: bool, n_devices: int = 1, n_envs: int | None = None,
                episode_loops: int | None = None):
    lrn = actor_network.lrn
    dev, net = lrn.dev, lrn.net
    A, d, a = net.n_agents, net.obs_dim, net.action_dim
    eval_episodes = config.arch.num_absolute_metric_eval_episodes if absolute_metric else config.arch.num_eval_episodes
    n = n_envs or get_num_eval_envs(config, absolute_metric, n_devices)
    loops = episode_loops or math.ceil(eval_episodes / (n * n_devices))
    greedy = bool(config.arch.evaluation_greedy)
    state = env.alloc_state(n, dev)
    ts = alloc_timestep(n, A, d, a, dev)
    lib = L.lib()
    lib.magpo_update_workspace_bytes.restype = C.c_size_t
    lib.magpo_rollout_workspace_bytes.restype = C.c_size_t
    nbytes = max(int(lib.magpo_update_workspace_bytes(C.byref(lrn.c_net), 1, n)),
                 int(lib.magpo_rollout_workspace_bytes(C.byref(lrn.c_net), n, 0)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    mb = dict(agents_view=ts["agents_view"].view(1, n, A, d), action_mask=ts["action_mask"].view(1, n, A, a),
              step_count=ts["step_count"].view(1, n, A), done=z(1, n, dt=torch.uint8), action=z(1, n, A, dt=torch.int32),
              value=z(1, n, A), log_prob=z(1, n, A), advantages=z(1, n, A), targets=z(1, n, A), policy_h0=z(n, A, net.hidden))
    dummy_h = {k: z(1, *net.state_shape) for k in ("encoder", "decoder_self", "decoder_cross")}
    mbs = L.struct_of(L.Minibatch, **mb)
    mbs.sable_h0 = L.struct_of(L.SableHState, **dummy_h)
    mbs.T, mbs.N = 1, n
    logits = z(1, n, A, a)
    noise = z(n * A * a)
    action = z(n, A, dt=torch.int32)

    def eval_fn(actor_flat: torch.Tensor, key) -> dict:
        t0 = time.perf_counter()
        s = L.stream_ptr()
        key = np.asarray(key, np.uint32)
        rets, lens = [], []
        for _ in range(loops):
            key, reset_key = minit.split(key, 2, dev)
            reset_keys = torch.as_tensor(minit.split(reset_key, n, dev).view(np.int32)).to(dev)
            L.call(env.reset_fn, s, C.byref(env.cfg), n, L.ptr(reset_keys), env.state_struct(state),
                   L.struct_of(L.TimeStep, **ts))
            hidden = mb["policy_h0"]
            hidden.zero_()
            ts["step_type"].zero_()
            last_l, ret_l, len_l = [], [], []
            step_key = key  # evaluator.py:139-146: the key the step scan advances is discarded, `_episode` returns the post-reset key
            for _ in range(env.time_limit + 1):
                step_key, act_key = minit.split(step_key, 2, dev)
                mb["done"].copy_((ts["step_type"] == 2).to(torch.uint8).view(1, n))  # timestep.last()
                L.call("magpo_actor_forward", lrn.ctx, s, C.byref(lrn.c_net), L.ptr(actor_flat), mbs, L.ptr(logits), L.ptr(ws),
                       C.c_size_t(nbytes))
                L.call("magpo_actor_step", lrn.ctx, s, C.byref(lrn.c_net), n, L.ptr(actor_flat), L.ptr(mb["agents_view"]),
                       L.ptr(mb["done"]), L.ptr(hidden), L.ptr(ws), C.c_size_t(nbytes))
                if greedy:
                    action.copy_(logits[0].argmax(-1))
                else:
                    k = torch.as_tensor(np.asarray(act_key, np.uint32).view(np.int32)).to(dev)
                    L.call("magpo_prng_gumbel", s, L.ptr(k), C.c_int64(n * A * a), L.ptr(noise))
                    action.copy_((logits[0] + noise.view(n, A, a)).argmax(-1))
                L.call(env.step_fn, s, C.byref(env.cfg), n, L.ptr(action), env.state_struct(state),
                       L.struct_of(L.TimeStep, **ts))
                last_l.append(ts["step_type"] == 2)
                ret_l.append(ts["episode_return"].clone())
                len_l.append(ts["episode_length"].clone())
            done_idx = torch.stack(last_l).to(torch.int8).argmax(0)  # first last() per env
            ar = torch.arange(n, device=dev)
            rets.append(torch.stack(ret_l)[done_idx, ar])
            lens.append(torch.stack(len_l)[done_idx, ar])
        metrics = dict(episode_return=torch.cat(rets), episode_length=torch.cat(lens))
        torch.cuda.synchronize()
        metrics["steps_per_second"] = float(metrics["episode_length"].sum()) / (time.perf_counter() - t0)
        return metrics

    return eval_fn


def get_sable_eval_fn(env, lrn, config, absolute_metric: bool, n_devices: int = 1, n_envs: int | None = None,
                      episode_loops: int | None = None):
    """The same evaluator loop with rec_sable's act function (rec_sable.py:497-516, `make_rec_sable_act_fn`): the Sable network
    acts — `get_actions(params, observation, hidden_state, act_key)` — and carries its three retention states, which start at zero
    for every episode loop (`get_init_hidden_state`, rec_sable.py:545-547) and are neither reset nor masked inside an episode."""
    dev, net = lrn.dev, lrn.net
    A, d, a = net.n_agents, net.obs_dim, net.action_dim
    eval_episodes = config.arch.num_absolute_metric_eval_episodes if absolute_metric else config.arch.num_eval_episodes
    n = n_envs or get_num_eval_envs(config, absolute_metric, n_devices)
    loops = episode_loops or math.ceil(eval_episodes / (n * n_devices))
    state = env.alloc_state(n, dev)
    ts = alloc_timestep(n, A, d, a, dev)
    lib = L.lib()
    lib.magpo_rollout_workspace_bytes.restype = C.c_size_t
    nbytes = int(lib.magpo_rollout_workspace_bytes(C.byref(lrn.c_net), n, 1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    hs = {k: z(n, *net.state_shape) for k in ("encoder", "decoder_self", "decoder_cross")}
    action, log_prob, value = z(n, A, dt=torch.int32), z(n, A), z(n, A)

    def eval_fn(guider_flat: torch.Tensor, key) -> dict:
        t0 = time.perf_counter()
        s = L.stream_ptr()
        key = np.asarray(key, np.uint32)
        rets, lens = [], []
        for _ in range(loops):
            key, reset_key = minit.split(key, 2, dev)
            reset_keys = torch.as_tensor(minit.split(reset_key, n, dev).view(np.int32)).to(dev)
            L.call(env.reset_fn, s, C.byref(env.cfg), n, L.ptr(reset_keys), env.state_struct(state), L.struct_of(L.TimeStep, **ts))
            for h in hs.values():
                h.zero_()
            last_l, ret_l, len_l = [], [], []
            step_key = key  # evaluator.py:139-146: the key the step scan advances is discarded, `_episode` returns the post-reset key
            for _ in range(env.time_limit + 1):
                step_key, act_key = minit.split(step_key, 2, dev)
                sample_keys = np.zeros((A, 2), np.uint32)  # discrete_autoregressive_act: key, sample_key = split(key) per agent
                k = act_key
                for i in range(A):
                    k, sample_keys[i] = minit.split(k, 2, dev)
                sk = torch.as_tensor(sample_keys.view(np.int32)).to(dev)
                L.call("magpo_sable_get_actions", lrn.ctx, s, C.byref(lrn.c_net), n, n, L.ptr(guider_flat), L.ptr(ts["agents_view"]),
                       L.ptr(ts["action_mask"]), L.ptr(ts["step_count"]), None, L.ptr(sk), L.struct_of(L.SableHState, **hs),
                       L.ptr(action), L.ptr(log_prob), L.ptr(value), None, L.ptr(ws), C.c_size_t(nbytes))
                L.call(env.step_fn, s, C.byref(env.cfg), n, L.ptr(action), env.state_struct(state), L.struct_of(L.TimeStep, **ts))
                last_l.append(ts["step_type"] == 2)
                ret_l.append(ts["episode_return"].clone())
                len_l.append(ts["episode_length"].clone())
            done_idx = torch.stack(last_l).to(torch.int8).argmax(0)
            ar = torch.arange(n, device=dev)
            rets.append(torch.stack(ret_l)[done_idx, ar])
            lens.append(torch.stack(len_l)[done_idx, ar])
        metrics = dict(episode_return=torch.cat(rets), episode_length=torch.cat(lens))
        torch.cuda.synchronize()
        metrics["steps_per_second"] = float(metrics["episode_length"].sum()) / (time.perf_counter() - t0)
        return metrics

    return eval_fn
