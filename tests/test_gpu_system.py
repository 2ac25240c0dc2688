"""-m gpu: the system entry (`learner_setup` / `learn`, rec_magpo.py:533-685,501-530) and the evaluator hook
(`actor_network.apply`) on the GPU: pytree shapes of Appendix B, metric shapes of scan∘vmap∘scan, state adoption."""
import os

import numpy as np
import pytest
import torch

from magpo_b200 import init as minit
from magpo_b200 import rec_magpo as rm
from magpo_b200.config import compose
from oracle import nets as onets
from oracle import learner as olr

from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _setup(dev, extra=()):
    cfg = compose("default/rec_magpo", ["arch.num_envs=6", "system.rollout_length=10", "system.ppo_epochs=2", "system.num_updates=4",
                                        "arch.num_evaluation=2", "system.total_timesteps=~", *extra])
    cfg.system.num_updates_per_eval = 2
    env = rm.make_env(cfg)
    key, _, ak, nk = minit.split(minit.prng_key(42), 4, dev)
    return cfg, env, rm.learner_setup(env, (key, ak, nk), cfg, device=dev)


def test_learn_shapes_and_state_contract(dev):
    cfg, env, (learn, actor_network, state) = _setup(dev)
    U, E, A, T, P, M = 2, 6, 3, 10, 2, 2
    assert state.key.shape == (1, U, 2) and state.dones.shape == (1, U, E, A)
    assert state.timestep.observation.agents_view.shape == (1, U, E, A, env.obs_dim)
    assert state.hstates.sable_hidden_state.encoder.shape == (1, U, E, 1, 1, 64, 64)
    assert state.hstates.policy_hidden_state.shape == (1, U, E, A, 128)
    assert state.env_state["record"].shape == (1, U, E, env.num_actions, env.time_limit)
    assert state.params.guider_params["encoder/encoder_block_0/retn/w_o"].shape == (1, U, 64, 64)
    p0 = state.params.actor_params["action_head/Dense_0/kernel"][0, 0].clone()
    out = learn(state)
    assert set(out.train_metrics) == {"total_loss", "value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy"}
    for v in out.train_metrics.values():
        assert v.shape == (1, 2, U, P, M)
    for k in ("episode_return", "episode_length", "is_terminal_step"):
        assert out.episode_metrics[k].shape == (1, 2, U, T, E)
    # optax.chain(clip_by_global_norm, adam) state tree: (EmptyState(), (ScaleByAdamState(count, mu, nu), EmptyState()))
    gos = out.learner_state.opt_states.guider_opt_state
    assert isinstance(gos[0], rm.EmptyState) and isinstance(gos[1][0], rm.ScaleByAdamState) and isinstance(gos[1][1], rm.EmptyState)
    assert int(rm.adam_of(gos).count[0, 0]) == 2 * P * M
    assert gos[1][0].mu["encoder/encoder_block_0/retn/w_o"].shape == (1, U, 64, 64)
    # CoordSum observations stay int32 through AgentIDWrapper (wrappers/observation.py:47-52); jumanji TimeStep field order + extras
    ts = out.learner_state.timestep
    assert ts._fields == ("step_type", "reward", "discount", "observation", "extras")
    assert ts.observation.agents_view.dtype == torch.int32 and ts.observation.action_mask.dtype == torch.bool
    assert set(ts.extras) == {"episode_metrics", "env_metrics", "real_next_obs"}
    assert ts.extras["episode_metrics"]["episode_return"].shape == (1, U, E)
    assert not torch.equal(out.learner_state.params.actor_params["action_head/Dense_0/kernel"][0, 0], p0)
    assert torch.isfinite(out.train_metrics["total_loss"]).all()
    # a foreign state (e.g. restored from a checkpoint) is adopted: zeroed parameters reach the device buffers
    foreign = out.learner_state._replace(params=rm.Params(
        {k: torch.zeros_like(v) for k, v in out.learner_state.params.guider_params.items()}, out.learner_state.params.actor_params))
    out2 = learn(foreign)
    # the update moved the parameters away from zero by at most a few Adam steps of size lr
    w = out2.learner_state.params.guider_params["encoder/encoder_block_0/retn/w_o"][0, 0]
    assert float(w.abs().max()) <= 2 * P * M * 2.5e-4 * 1.01


def test_actor_network_apply_matches_oracle(dev):
    cfg, env, (learn, actor_network, state) = _setup(dev)
    lrn = actor_network.lrn
    rng = np.random.default_rng(0)
    N, A, d, a = 7, env.num_agents, env.obs_dim, env.action_dim
    obs = rng.integers(0, 5, (1, N, A, d)).astype(np.float32)
    mask = rng.random((1, N, A, a)) < 0.8
    mask[..., 0] = True
    done = rng.random((1, N)) < 0.3
    h = (rng.standard_normal((N, A, 128)) * 0.3).astype(np.float32)
    t = lambda x: torch.as_tensor(x).to(dev)
    # RecurrentActor.__call__(hstate, (observation, done)) (networks/base.py:161-165); done per agent as the reference passes it
    carry, logits = actor_network.apply(lrn.actor, t(h), (rm.Observation(t(obs), t(mask), t(np.zeros((1, N, A), np.int32))),
                                                          t(np.repeat(done[..., None], A, -1))))
    _, ap = lrn.get_params()
    ncfg = onets.NetCfg(A, d, a)
    p = onets.to_torch({k: v.cpu().numpy() for k, v in ap.items()})
    h_ref, l_ref = onets.actor_apply(p, ncfg, torch.tensor(h), torch.tensor(obs), torch.tensor(np.repeat(done[..., None], A, -1)),
                                     torch.tensor(mask))
    got = logits.cpu().numpy()
    assert rel_err(got[mask], l_ref.numpy()[mask]) < 1e-4
    assert (got[~mask] == np.finfo(np.float32).min).all()
    assert rel_err(carry.cpu().numpy(), h_ref.numpy()) < 1e-4


def test_adopting_a_foreign_state_reproduces_the_run(dev):
    """`learn` honours every component of the state it is given (systems/gpo/types.py:62-71), not only params / optimiser / key:
    a deep copy of learner A's state (fresh tensors everywhere: env_state, timestep, dones, hstates) handed to learner B, which was
    set up from another seed, makes B continue exactly like A."""
    cfg, env, (learn_a, net_a, state_a) = _setup(dev)
    out_a = learn_a(state_a)
    cfg_b = compose("default/rec_magpo", ["arch.num_envs=6", "system.rollout_length=10", "system.ppo_epochs=2", "system.num_updates=4",
                                          "arch.num_evaluation=2", "system.total_timesteps=~", "system.seed=7"])
    cfg_b.system.num_updates_per_eval = 2
    key, _, ak, nk = minit.split(minit.prng_key(7), 4, dev)
    learn_b, net_b, state_b = rm.learner_setup(rm.make_env(cfg_b), (key, ak, nk), cfg_b, device=dev)
    learn_b(state_b)  # B has history of its own (first_rollout False, other params / env states)

    def clone(x):
        if isinstance(x, torch.Tensor):
            return x.clone()
        if isinstance(x, dict):
            return {k: clone(v) for k, v in x.items()}
        if isinstance(x, tuple):
            vals = [clone(v) for v in x]
            return type(x)(*vals) if hasattr(x, "_fields") else tuple(vals)
        return x

    foreign = clone(out_a.learner_state)
    ref = learn_a(out_a.learner_state)
    got = learn_b(foreign)
    for k in ("episode_return", "episode_length", "is_terminal_step"):
        assert torch.equal(ref.episode_metrics[k], got.episode_metrics[k]), k
    # the rollouts are bit-identical; the update's cross-CTA fp32 reductions are order-dependent in the last bits
    for k, v in ref.train_metrics.items():
        assert torch.allclose(v, got.train_metrics[k], rtol=1e-5, atol=1e-6), k
    for k, v in ref.learner_state.params.actor_params.items():
        w = got.learner_state.params.actor_params[k]
        assert float((v - w).abs().max()) <= 2e-6 * max(float(v.abs().max()), 1e-3), k
    assert torch.equal(ref.learner_state.key, got.learner_state.key)
    assert torch.equal(ref.learner_state.env_state["record"], got.learner_state.env_state["record"])


def test_decay_learning_rates_matches_oracle(dev):
    """system.decay_learning_rates (utils/training.py:30-64): the schedule evaluated on the device from the optimiser count."""
    from magpo_b200.learner import CoordSumVec, MagpoLearner, SystemConfig
    from oracle import coordsum as ocs, prng as oprng

    kw = ocs.SCENARIOS["3x10-30-v0"]
    spec = ocs.CoordSumSpec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    E, U, T, P, M, NU = 4, 1, 8, 2, 2, 3
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M, decay_learning_rates=True,
                      num_updates=NU)
    state = olr.learner_setup(spec, ncfg, osys, seed=5)
    lrn = MagpoLearner(CoordSumVec(**kw), SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M,
                                                       decay_learning_rates=True, num_updates=NU), device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(5), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    p_prev = {k: v.copy() for k, v in state["actor_params"].items()}
    steps = []
    for upd in range(NU):
        rec = {}
        olr.update_step(state, spec, ncfg, osys, record=rec)
        lrn.update_step()
        torch.cuda.synchronize()
        assert (lrn.traj["action"].cpu().numpy() == rec["traj"][0]["action"]).all(), upd
        _, ap = lrn.get_params()
        for name, r in state["actor_params"].items():
            assert np.abs(ap[name].cpu().numpy() - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-3), (upd, name)
        steps.append(max(float(np.abs(state["actor_params"][k] - p_prev[k]).max()) for k in p_prev))
        p_prev = {k: v.copy() for k, v in state["actor_params"].items()}
    # lr decays 1 -> 2/3 -> 1/3 of actor_lr over the three updates: the parameter steps shrink accordingly
    assert steps[0] > steps[1] > steps[2] > 0
    sv = rm._replicated_views(lrn)[1].guider_opt_state
    assert isinstance(sv[1][1], rm.ScaleByScheduleState) and int(sv[1][1].count[0, 0]) == NU * P * M


def test_network_config_is_read_from_the_config(dev):
    """config.network is honoured or refused, never ignored: decay_scaling_factor and the PE switch reach the kernels,
    shapes the kernels do not cover raise."""
    cfg, env, (learn, net, state) = _setup(dev, extra=["network.memory_config.decay_scaling_factor=0.5",
                                                        "network.memory_config.timestep_positional_encoding=False"])
    assert abs(net.lrn.net.decay_scaling_factor - 0.5) < 1e-9 and net.lrn.net.timestep_pe is False
    for bad in (["network.net_config.n_block=4"], ["network.net_config.embed_dim=96"], ["network.net_config.n_head=3"],
                ["network.net_config.embed_dim=32", "network.net_config.n_head=4", "network.net_config.n_block=1", "network.hidden_state_dim=64"],
                ["network.hidden_state_dim=64"], ["network.actor_network.pre_torso.layer_sizes=[64,64]"], ["system.add_agent_id=False"]):
        with pytest.raises(NotImplementedError):
            _setup(dev, extra=bad)


@pytest.mark.parametrize("pe,dsf", [(False, 0.8), (True, 0.5)])
def test_memory_config_variants_match_oracle(dev, pe, dsf):
    """memory_config.timestep_positional_encoding=False (retention.py:278,304: no PE added) and another decay_scaling_factor:
    one whole update step against the oracle built with the same NetCfg."""
    from magpo_b200.learner import CoordSumVec, MagpoLearner, NetworkConfig, SystemConfig
    from oracle import coordsum as ocs, prng as oprng

    kw = ocs.SCENARIOS["3x10-30-v0"]
    spec = ocs.CoordSumSpec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim, timestep_pe=pe, decay_scaling_factor=dsf)
    E, U, T, P, M = 6, 1, 12, 1, 2
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=11)
    vec = CoordSumVec(**kw)
    lrn = MagpoLearner(vec, SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M), device=dev,
                       net=NetworkConfig(vec.num_agents, vec.obs_dim, vec.action_dim, vec.time_limit, timestep_pe=pe, decay_scaling_factor=dsf))
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(11), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    rec = {}
    _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
    _, losses = lrn.update_step()
    torch.cuda.synchronize()
    assert (lrn.traj["action"].cpu().numpy() == rec["traj"][0]["action"]).all()
    assert rel_err(lrn.traj["value"].cpu().numpy(), rec["traj"][0]["value"]) < 1e-4
    li = rm.MagpoLearner.loss_info(losses.cpu(), lrn.sys)
    for m in range(M):
        for name in ("value_loss", "guider_loss", "entropy", "kl_loss"):
            assert abs(float(li[name][0, m]) - infos[m][name]) <= 2e-4 * max(1.0, abs(infos[m][name])), (m, name)
    gp, _ = lrn.get_params()
    for name, r in state["guider_params"].items():
        assert np.abs(gp[name].cpu().numpy() - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-3), name


@pytest.mark.parametrize("greedy", [True, False])
def test_evaluator_matches_oracle(dev, greedy):
    """evaluator.py:82-163,188-208 on the GPU against the CPU restatement: same keys -> same per-episode returns / lengths."""
    from magpo_b200 import evaluator as mev
    from oracle import coordsum as ocs
    from oracle import evaluator as oev
    from oracle import prng as oprng

    cfg, env, (learn, actor_network, state) = _setup(dev, extra=[f"arch.evaluation_greedy={greedy}"])
    learn(state)  # move the policy off its initialisation
    lrn = actor_network.lrn
    n, loops = 5, 2
    key = oprng.split(oprng.prng_key(3))[1]
    eval_fn = mev.get_eval_fn(env, actor_network, cfg, absolute_metric=False, n_envs=n, episode_loops=loops)
    got = eval_fn(lrn.actor, key)
    _, ap = lrn.get_params()
    spec = ocs.CoordSumSpec(**rm.COORDSUM_REGISTRY[cfg.env.scenario.task_name])
    ncfg = onets.NetCfg(env.num_agents, env.obs_dim, env.action_dim)
    ref = oev.eval_fn(spec, ncfg, {k: v.cpu().numpy() for k, v in ap.items()}, key, n, loops, greedy=greedy)
    assert got["episode_length"].shape == (n * loops,)
    assert (got["episode_length"].cpu().numpy() == ref["episode_length"]).all()
    assert np.allclose(got["episode_return"].cpu().numpy(), ref["episode_return"], rtol=0, atol=1e-5)
    assert got["steps_per_second"] > 0


@pytest.mark.parametrize("env_args,env_dims", [([], (3, 10)), (["env=lbf"], (2, 6)),
                                               (["env=rware", "env/scenario=tiny-4ag", "env.kwargs.time_limit=30"], (4, 5))])
def test_run_experiment_trains_and_evaluates(dev, env_args, env_dims, tmp_path_factory, monkeypatch):
    """rec_magpo.py:688-815 end to end (`python -m magpo_b200.rec_magpo env=...`): learn, evaluate the learner policy after every
    `learn`, absolute metric with the best parameters at the end — on all three env families."""
    lines = []
    cfg = compose("default/rec_magpo", [*env_args, "arch.num_envs=8", "system.rollout_length=8", "system.num_updates=2",
                                        "arch.num_evaluation=2", "system.total_timesteps=~", "arch.num_eval_episodes=4",
                                        "arch.num_absolute_metric_eval_episodes=8"])
    tmp_path = tmp_path_factory.mktemp("run")
    monkeypatch.chdir(tmp_path)
    cfg.logger.checkpointing.save_model = True
    cfg.logger.use_json = True
    cfg.logger.base_exp_path = str(tmp_path / "results")
    perf = rm.run_experiment(cfg, device=dev, log=lines.append)
    kinds = [ln.split(" - ")[0] for ln in lines]
    assert kinds.count("EVALUATOR") == 2 and kinds.count("TRAINER") == 2 and kinds.count("ABSOLUTE") == 1 and kinds.count("MISC") == 2
    assert any("Episode return mean" in ln for ln in lines if ln.startswith("EVALUATOR"))
    assert np.isfinite(perf)
    # checkpoint of the unreplicated learner state (rec_magpo.py:779-786) and the marl-eval JSON
    from magpo_b200.checkpointing import Checkpointer
    uid = os.listdir(tmp_path / "checkpoints" / "rec_magpo")[0]
    ck = Checkpointer("rec_magpo", checkpoint_uid=uid)
    flat = ck.restore()
    assert flat["params/actor_params/action_head/Dense_0/kernel"].shape == (128, env_dims[1])
    assert flat["hstates/policy_hidden_state"].shape == (8, env_dims[0], 128) and flat["key"].shape == (2,)
    jf = list((tmp_path / "results").rglob("metrics.json"))
    assert len(jf) == 1 and "absolute_metrics" in open(jf[0]).read()
