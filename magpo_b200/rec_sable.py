"""System entry mirroring `mava/systems/sable/anakin/rec_sable.py`: the Sable network alone, trained with the
clipped PPO objective — the guider half of rec_magpo on the same kernels.

`learner_setup(env, (key, net_key), config) -> (learn, sable_execution_fn, init_learner_state)` (rec_sable.py:351-479),
`LearnerState(params, opt_states, key, env_state, timestep, hstates)` (mava/systems/sable/types.py), `learn` returning
`ExperimentOutput` with the four train metrics of rec_sable.py:251-256. Underneath it is `MagpoLearner` with `sable_only=True`: the
rollout skips the learner's GRU push, the update skips the learner's pass and feeds the guider's own log-probs where MAGPO's loss
reads the learner's — the guidance KL is then identically 0 and the double-clipped ratio is the PPO ratio (rec_sable.py:196-226;
checked against a direct restatement of that loss in oracle/learner.py `sable_ppo_loss`).

    python -m magpo_b200.rec_sable env=rware env/scenario=tiny-4ag arch.num_envs=1024 system.num_updates=8 arch.num_evaluation=2
"""
from __future__ import annotations

import sys
import time
from typing import Any, Dict, NamedTuple, Tuple

import numpy as np
import torch

from . import init as minit
from . import rec_magpo as rm
from .config import Config, check_total_timesteps, compose
from .learner import MagpoLearner


class LearnerState(NamedTuple):
    """mava/systems/sable/types.py LearnerState; every leaf with the leading [1, U, ...] of this device."""
    params: Dict[str, torch.Tensor]
    opt_states: tuple  # optax.chain(clip_by_global_norm, adam) state, see rec_magpo.optax_state
    key: torch.Tensor
    env_state: Dict[str, torch.Tensor]
    timestep: rm.TimeStep
    hstates: rm.SableHiddenStates


def _state_views(lrn: MagpoLearner) -> LearnerState:
    g = rm._state_views(lrn)
    return LearnerState(g.params.guider_params, g.opt_states.guider_opt_state, g.key, g.env_state, g.timestep,
                        g.hstates.sable_hidden_state)


def get_learner_fn(lrn: MagpoLearner, config: Config):
    """rec_sable.py:53-349: `learn` = num_updates_per_eval x `_update_step`."""
    n_upd = int(config.system.num_updates_per_eval)
    U, E = lrn.sys.update_batch_size, lrn.sys.num_envs

    def learn(learner_state: LearnerState) -> rm.ExperimentOutput:
        # adopt a foreign state (checkpoint restore, edited leaves); the state `learn` returned last costs nothing
        rm._adopt(lrn, rm.GPOLearnerState(rm.Params(learner_state.params, None), rm.OptStates(learner_state.opt_states, None),
                                          learner_state.key, learner_state.env_state, learner_state.timestep, None,
                                          rm.HiddenStates(learner_state.hstates, None)))
        ep = {k: [] for k in ("episode_return", "episode_length", "is_terminal_step")}
        tr = []
        for _ in range(n_upd):
            metrics, losses = lrn.update_step()
            for k in ep:
                ep[k].append(metrics[k].reshape(-1, U, E).permute(1, 0, 2).clone())
            tr.append(losses.clone())
        episode_metrics = {k: torch.stack(v)[None] for k, v in ep.items()}
        episode_metrics["is_terminal_step"] = episode_metrics["is_terminal_step"].bool()
        li = MagpoLearner.loss_info(torch.stack(tr), lrn.sys)
        # rec_sable.py:250-256: total = actor_loss - ent_coef * entropy + vf_coef * value_loss (the guider's total; the KL slot is 0)
        info = dict(total_loss=li["guider_loss"] + li["kl_loss"] - lrn.sys.ent_coef * li["entropy"] + lrn.sys.vf_coef * li["value_loss"],
                    value_loss=li["value_loss"], actor_loss=li["guider_loss"], entropy=li["entropy"])
        train_metrics = {k: v[None, :, None].expand(1, n_upd, U, *v.shape[1:]) for k, v in info.items()}
        return rm.ExperimentOutput(_state_views(lrn), episode_metrics, train_metrics)

    return learn


def learner_setup(env, keys: Tuple[Any, Any], config: Config, device=None, allreduce=None, rank: int = 0, world_size: int = 1, comm=None):
    """rec_sable.py:351-479."""
    key, net_key = keys
    sysc = rm._system_config(_with_magpo_defaults(config))
    sysc.sable_only = True
    lrn = MagpoLearner(env, sysc, device=device or "cuda:0", allreduce=allreduce, world_size=world_size,
                       net=rm._network_config(config, env))
    if comm is not None:
        comm.attach(lrn)
    lrn.set_params(minit.flax_init_guider(np.asarray(net_key, np.uint32), env.num_agents, env.obs_dim, env.action_dim, lrn.dev,
                                          lrn.net.embed_dim, lrn.net.n_head, lrn.net.n_block),
                   minit.init_actor(env.obs_dim, env.action_dim, 0))  # the learner buffers exist but are never touched
    U, E = sysc.update_batch_size, sysc.num_envs
    allk = minit.split(np.asarray(key, np.uint32), world_size * U * E + 1, lrn.dev)
    step_key = minit.split(allk[0], 2, lrn.dev)[1]
    lrn.reset(rm.shard_env_keys(allk[1:], world_size, rank, U, E), step_key)
    return get_learner_fn(lrn, config), lrn, _state_views(lrn)


def _with_magpo_defaults(config: Config) -> Config:
    for k, v in (("clip_gpo", 1.5), ("alpha", 1.0)):  # unused by rec_sable; the shared SystemConfig carries them
        if config.system.get(k) is None:
            setattr(config.system, k, v)
    return config


def run_experiment(config: Config, device=None, log=print) -> float:
    """rec_sable.py:481-625: `num_evaluation` x (`learn`, evaluation of the Sable policy with `make_rec_sable_act_fn`), logging
    MISC / ACT / TRAIN / EVAL events, optional checkpoints, the absolute metric with the best parameters at the end."""
    import os

    from . import evaluator as mev
    from .checkpointing import Checkpointer, unreplicate_n_dims
    from .logger import LogEvent, MavaLogger

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))  # one process per GPU (torchrun)
    comm = None
    if world > 1:
        from .comm import NcclComm
        comm = NcclComm.from_env(device)
    config = check_total_timesteps(config, world)
    config.system.num_updates_per_eval = config.system.num_updates // config.arch.num_evaluation
    env = rm.make_env(config)
    key, key_e, net_key = minit.split(minit.prng_key(int(config.system.seed)), 3, device or "cuda:0")
    learn, lrn, state = learner_setup(env, (key, net_key), config, device=device, comm=comm, rank=rank, world_size=world)
    evaluator = mev.get_sable_eval_fn(env, lrn, config, absolute_metric=False, n_devices=world)
    config.logger.system_name = "rec_sable"
    logger = MavaLogger(config, console_sink=log) if rank == 0 else None
    if logger:
        logger.log_config(config.to_dict())
    save_checkpoint = bool(config.logger.checkpointing.save_model) and rank == 0
    if save_checkpoint:
        checkpointer = Checkpointer(metadata=config, model_name=config.logger.system_name,
                                    **config.logger.checkpointing.save_args.to_dict())
    steps = world * config.system.num_updates_per_eval * config.system.rollout_length * config.system.update_batch_size * config.arch.num_envs

    def world_mean(x: torch.Tensor) -> float:
        m = x.float().mean().reshape(1).to(lrn.dev)
        if comm is not None:
            comm.allreduce_sum(m)
            m /= world
        return float(m)

    max_episode_return, best_params, eval_performance = float("-inf"), None, float("nan")
    for ev in range(int(config.arch.num_evaluation)):
        t0 = time.perf_counter()
        out = learn(state)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        state = out.learner_state
        t = int(steps * (ev + 1))
        term = out.episode_metrics["is_terminal_step"].bool()
        if logger:
            logger.log({"timestep": t}, t, ev, LogEvent.MISC)
            if bool(term.any()):
                act = {k: v[term] for k, v in out.episode_metrics.items() if k != "is_terminal_step"}
                act["steps_per_second"] = steps / dt
                logger.log(act, t, ev, LogEvent.ACT)
            logger.log(dict(out.train_metrics), t, ev, LogEvent.TRAIN)
        ks = minit.split(key_e, world + 1, device or "cuda:0")
        key_e, eval_key = ks[0], ks[1 + rank]
        eval_metrics = evaluator(lrn.guider, eval_key)
        eval_performance = world_mean(eval_metrics[config.env.eval_metric])
        if logger:
            logger.log(dict(eval_metrics), t, ev, LogEvent.EVAL)
        if save_checkpoint:
            checkpointer.save(timestep=t, unreplicated_learner_state=unreplicate_n_dims(state), episode_return=eval_performance)
        if config.arch.absolute_metric and max_episode_return <= eval_performance:
            best_params, max_episode_return = lrn.guider.clone(), eval_performance
    if config.arch.absolute_metric and best_params is not None:
        abs_evaluator = mev.get_sable_eval_fn(env, lrn, config, absolute_metric=True, n_devices=world)
        ks = minit.split(key_e, world + 1, device or "cuda:0")
        abs_metrics = abs_evaluator(best_params, ks[1 + rank])
        world_mean(abs_metrics[config.env.eval_metric])
        if logger:
            logger.log(dict(abs_metrics), int(steps * int(config.arch.num_evaluation)), int(config.arch.num_evaluation) - 1, LogEvent.ABSOLUTE)
    if logger:
        logger.stop()
    return eval_performance


def main(argv=None) -> float:
    return run_experiment(compose("default/rec_sable", list(sys.argv[1:] if argv is None else argv)))


if __name__ == "__main__":
    main()
