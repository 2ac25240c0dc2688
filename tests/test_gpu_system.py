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
    assert int(out.learner_state.opt_states.guider_opt_state.count[0, 0]) == 2 * P * M
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
    carry, logits = actor_network.apply(lrn.actor, t(h), rm.Observation(t(obs), t(mask), t(np.zeros((1, N, A), np.int32))), t(done))
    _, ap = lrn.get_params()
    ncfg = onets.NetCfg(A, d, a)
    p = onets.to_torch({k: v.cpu().numpy() for k, v in ap.items()})
    h_ref, l_ref = onets.actor_apply(p, ncfg, torch.tensor(h), torch.tensor(obs), torch.tensor(np.repeat(done[..., None], A, -1)),
                                     torch.tensor(mask))
    got = logits.cpu().numpy()
    assert rel_err(got[mask], l_ref.numpy()[mask]) < 1e-4
    assert (got[~mask] == np.finfo(np.float32).min).all()
    assert rel_err(carry.cpu().numpy(), h_ref.numpy()) < 1e-4


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
