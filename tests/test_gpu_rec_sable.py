"""-m gpu: the rec_sable sibling (mava/systems/sable/anakin/rec_sable.py) on the rec_magpo kernels: `sable_only` update steps against
a direct restatement of rec_sable's loss (oracle/learner.py `sable_ppo_loss`), and the system entry's pytree / metric shapes."""
import numpy as np
import pytest
import torch

from magpo_b200 import init as minit
from magpo_b200 import rec_sable as rs
from magpo_b200.config import compose
from magpo_b200.learner import CoordSumVec, LbfVec, MagpoLearner, SystemConfig
from oracle import coordsum as ocs
from oracle import lbf as olbf
from magpo_b200 import rec_magpo as rm
from oracle import learner as olr
from oracle import nets as onets
from oracle import prng as oprng

from gpu_util import as_u32, rel_err, sync

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_name", ["lbf", "coordsum"])
def test_sable_only_update_matches_rec_sable_loss(dev, env_name):
    if env_name == "lbf":
        kw = olbf.SCENARIOS["2s-8x8-2p-2f-coop"]
        spec, vec = olbf.LbfSpec(**kw), LbfVec(**kw)
    else:
        kw = ocs.SCENARIOS["3x10-30-v0"]
        spec, vec = ocs.CoordSumSpec(**kw), CoordSumVec(**kw)
    E, U, T, P, M = 6, 2, 20, 2, 2
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M, sable_only=True)
    state = olr.learner_setup(spec, ncfg, osys, seed=11)
    a0 = {k: v.copy() for k, v in state["actor_params"].items()}
    lrn = MagpoLearner(vec, SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M,
                                         sable_only=True), device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(11), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    for it in range(2):
        rec = {}
        _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
        _, losses = lrn.update_step()
        sync()
        for u in range(U):
            sl = slice(u * E, (u + 1) * E)
            assert (lrn.traj["action"].cpu().numpy()[:, sl] == rec["traj"][u]["action"]).all(), (it, "actions")
            assert (lrn.traj["reward"].cpu().numpy()[:, sl] == rec["traj"][u]["reward"]).all()
            assert rel_err(lrn.traj["value"].cpu().numpy()[:, sl], rec["traj"][u]["value"]) < 1e-4
        li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
        k = 0
        for p in range(P):
            for m in range(M):
                for ours, theirs in (("value_loss", "value_loss"), ("guider_loss", "actor_loss"), ("entropy", "entropy")):
                    ref, got = infos[k][theirs], float(li[ours][p, m])
                    assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (it, p, m, ours, got, ref)
                assert float(li["kl_loss"][p, m]) == 0.0  # KL(pi || pi) is exactly zero
                k += 1
        gp, ap = lrn.get_params()
        for name, r in state["guider_params"].items():
            assert np.abs(gp[name].cpu().numpy() - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-3), (it, name)
        for name, r in a0.items():  # no learner in rec_sable: its buffers never move
            assert (ap[name].cpu().numpy() == r).all(), name
        assert (as_u32(lrn.key) == state["slots"][0]["key"]).all()


def test_rec_sable_system_entry(dev):
    """learner_setup / learn of rec_sable.py:351-479,319-349: LearnerState pytree and the four train metrics."""
    cfg = compose("default/rec_sable", ["env=lbf", "arch.num_envs=6", "system.rollout_length=10", "system.ppo_epochs=2",
                                        "system.num_updates=4", "arch.num_evaluation=2"])
    cfg.system.num_updates_per_eval = 2
    env = rs.rm.make_env(cfg)
    key, _, nk = minit.split(minit.prng_key(42), 3, dev)
    learn, _, state = rs.learner_setup(env, (key, nk), cfg, device=dev)
    U, E, A, T, P, M = 2, 6, 2, 10, 2, 2
    assert state._fields == ("params", "opt_states", "key", "env_state", "timestep", "hstates")
    assert state.hstates.encoder.shape == (1, U, E, 1, 1, 64, 64) and state.key.shape == (1, U, 2)
    w0 = state.params["decoder/head/layers_3/kernel"][0, 0].clone()
    out = learn(state)
    assert set(out.train_metrics) == {"total_loss", "value_loss", "actor_loss", "entropy"}
    for v in out.train_metrics.values():
        assert v.shape == (1, 2, U, P, M) and torch.isfinite(v).all()
    assert out.episode_metrics["episode_return"].shape == (1, 2, U, T, E)
    assert int(rm.adam_of(out.learner_state.opt_states).count[0, 0]) == 2 * P * M
    assert not torch.equal(out.learner_state.params["decoder/head/layers_3/kernel"][0, 0], w0)
    lines = []
    perf = rs.run_experiment(compose("default/rec_sable", ["env=lbf", "arch.num_envs=8", "system.rollout_length=8", "system.num_updates=2",
                                                           "arch.num_evaluation=2", "arch.num_eval_episodes=4",
                                                           "arch.num_absolute_metric_eval_episodes=8"]), device=dev, log=lines.append)
    kinds = [ln.split(" - ")[0] for ln in lines]
    assert kinds.count("TRAINER") == 2 and kinds.count("EVALUATOR") == 2 and kinds.count("ABSOLUTE") == 1 and np.isfinite(perf)


def test_sable_evaluator_matches_oracle(dev):
    """rec_sable's evaluation (rec_sable.py:497-516 + evaluator.py:82-163): same keys -> same per-episode returns / lengths."""
    from magpo_b200 import evaluator as mev
    from oracle import evaluator as oev

    cfg = compose("default/rec_sable", ["env=lbf", "arch.num_envs=6", "system.rollout_length=10", "system.num_updates=4",
                                        "arch.num_evaluation=2"])
    cfg.system.num_updates_per_eval = 2
    env = rs.rm.make_env(cfg)
    key, _, nk = minit.split(minit.prng_key(42), 3, dev)
    learn, lrn, state = rs.learner_setup(env, (key, nk), cfg, device=dev)
    learn(state)  # move the policy off its initialisation
    n, loops = 5, 2
    ekey = oprng.split(oprng.prng_key(3))[1]
    got = mev.get_sable_eval_fn(env, lrn, cfg, absolute_metric=False, n_envs=n, episode_loops=loops)(lrn.guider, ekey)
    gp, _ = lrn.get_params()
    spec = olbf.LbfSpec(**olbf.SCENARIOS["2s-8x8-2p-2f-coop"])
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    ref = oev.eval_fn_sable(spec, ncfg, {k: v.cpu().numpy() for k, v in gp.items()}, ekey, n, loops)
    assert (got["episode_length"].cpu().numpy() == ref["episode_length"]).all()
    assert np.allclose(got["episode_return"].cpu().numpy(), ref["episode_return"], rtol=0, atol=1e-5)
