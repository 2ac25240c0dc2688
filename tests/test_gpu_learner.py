"""-m gpu: the whole `_update_step` (rollout -> GAE -> P epochs x M minibatches -> clip+Adam) of the CUDA path
against the CPU oracle on the same seeds (config c1 shape, shortened). Env transitions / rewards / dones /
sampled actions bit-exact; values, log-probs, advantages, losses, updated parameters within fp32 rtol 1e-4."""
import numpy as np
import pytest
import torch

from magpo_b200.learner import CoordSumVec, MagpoLearner, SystemConfig, param_views
from oracle import coordsum as ocs
from oracle import learner as olr
from oracle import nets as onets
from oracle import prng as oprng

from gpu_util import as_u32, rel_err, sync

pytestmark = pytest.mark.gpu


def build(dev, E=8, U=2, T=16, P=2, M=2, scenario="3x10-30-v0", seed=42, chunk=0):
    kw = ocs.SCENARIOS[scenario]
    spec = ocs.CoordSumSpec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=seed)
    sysc = SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M, chunk_envs=chunk)
    lrn = MagpoLearner(CoordSumVec(**kw), sysc, device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    # the same key derivation as learner_setup (rec_magpo.py:642-673)
    ks = oprng.split(oprng.prng_key(seed), 4)
    allk = oprng.split(ks[0], U * E + 1)
    step_key = oprng.split(allk[0])[1]
    lrn.reset(allk[1:], step_key)
    return spec, ncfg, osys, state, lrn


def test_rollout_and_gae_match_oracle(dev):
    spec, ncfg, osys, state, lrn = build(dev)
    E, U, T, A = osys.num_envs, osys.update_batch_size, osys.rollout_length, ncfg.n_agents
    lrn.rollout()
    lrn.gae()
    sync()
    tr = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in lrn.traj.items() if k != "sable_h0"}
    adv_c, tgt_c = lrn.adv.cpu().numpy(), lrn.targets.cpu().numpy()
    for u, slot in enumerate(state["slots"]):
        sl = slice(u * E, (u + 1) * E)
        traj, met = olr.rollout(spec, ncfg, osys, state["guider_params"], state["actor_params"], slot)
        assert (tr["action"][:, sl] == traj["action"]).all(), "sampled actions differ"
        assert (tr["reward"][:, sl] == traj["reward"]).all()
        assert (tr["done"][:T, sl] == traj["done"][:, :, 0]).all()
        assert (tr["agents_view"][:T, sl] == traj["obs"].astype(np.float32)).all()
        assert (tr["step_count"][:T, sl] == traj["step_count"]).all()
        assert (tr["episode_return"][:, sl] == met["episode_return"]).all()
        assert (tr["is_terminal_step"][:, sl].astype(bool) == met["is_terminal_step"]).all()
        assert rel_err(tr["value"][:, sl], traj["value"]) < 1e-4
        assert rel_err(tr["log_prob"][:, sl], traj["log_prob"]) < 1e-4
        assert rel_err(lrn.policy_h.cpu().numpy()[sl], slot["hstates"]["policy"]) < 1e-4
        hs = lrn.sable_hidden_state()
        for name, ref in zip(("encoder", "decoder_self", "decoder_cross"), slot["hstates"]["sable"]):
            assert rel_err(hs[name].cpu().numpy()[sl], ref) < 1e-4, name  # [E, n_head = 1, n_block = 1, 64, 64]
        last_val = olr.bootstrap_value(ncfg, state["guider_params"], slot)
        assert rel_err(tr["last_value"][sl], last_val) < 1e-4
        adv, tgt = olr.gae(traj["done"], traj["value"], traj["reward"], last_val, slot["dones"], osys.gamma, osys.gae_lambda)
        assert rel_err(adv_c[:, sl], adv) < 1e-4 and rel_err(tgt_c[:, sl], tgt) < 1e-4
        assert (as_u32(lrn.key) == slot["key"]).all()
        # env state after the rollout
        assert (lrn.env_state["record"].cpu().numpy()[sl] == slot["env_state"]["env_state"]["record"]).all()
        assert (as_u32(lrn.env_state["key"])[sl] == slot["env_state"]["env_state"]["key"]).all()


@pytest.mark.parametrize("chunk", [0, 3])
def test_update_step_matches_oracle(dev, chunk):
    spec, ncfg, osys, state, lrn = build(dev, E=8, U=2, T=16, P=2, M=2, chunk=chunk)
    g0 = {k: v.copy() for k, v in state["guider_params"].items()}
    a0 = {k: v.copy() for k, v in state["actor_params"].items()}
    rec = {}
    mets, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
    metrics, losses = lrn.update_step()
    sync()
    # trajectories must agree exactly for the update comparison to be meaningful
    for u in range(osys.update_batch_size):
        sl = slice(u * osys.num_envs, (u + 1) * osys.num_envs)
        assert (lrn.traj["action"].cpu().numpy()[:, sl] == rec["traj"][u]["action"]).all()
    li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
    k = 0
    for p in range(osys.ppo_epochs):
        for m in range(osys.num_minibatches):
            for name in ("value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy", "total_loss"):
                ref = infos[k][name]
                got = float(li[name][p, m])
                assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (p, m, name, got, ref)
            k += 1
    gp, ap = lrn.get_params()
    worst = 0.0
    for new, ref, old in ((gp, state["guider_params"], g0), (ap, state["actor_params"], a0)):
        for name, r in ref.items():
            got = new[name].cpu().numpy()
            assert np.abs(got - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-3), name
            step = np.abs(r - old[name]).max()
            if step > 0:
                worst = max(worst, np.abs(got - r).max() / step)
    print("worst |p_cuda - p_oracle| / |update| =", worst)
    assert worst < 0.05
    assert int(lrn.g_count.item()) == osys.ppo_epochs * osys.num_minibatches
    assert (as_u32(lrn.key) == state["slots"][0]["key"]).all()


def test_second_rollout_carries_state(dev):
    """Three consecutive update steps (eager rollout, CUDA-graph capture + replay, replay): the carried timestep / hidden
    states / env state keep matching the oracle."""
    spec, ncfg, osys, state, lrn = build(dev, E=4, U=1, T=110, P=1, M=1)  # T > time_limit: crosses an auto-reset
    for it in range(3):
        rec = {}
        olr.update_step(state, spec, ncfg, osys, record=rec)
        lrn.update_step()
        sync()
        assert (lrn.traj["action"].cpu().numpy() == rec["traj"][0]["action"]).all(), it
        assert (lrn.traj["reward"].cpu().numpy() == rec["traj"][0]["reward"]).all(), it
        assert rel_err(lrn.traj["value"].cpu().numpy(), rec["traj"][0]["value"]) < 2e-4, it


@pytest.mark.parametrize("scenario,E,T", [("5x20-80-v0", 4, 9), ("8x15-100-v0", 3, 6), ("3x30-50-v0", 5, 12)])
def test_other_scenarios_match_oracle(dev, scenario, E, T):
    """The other registered CoordSum scenarios (coordsum/__init__.py:6-45): A = 5 and 8 take the per-layer rollout path
    (the fused step kernel covers A <= 4), a = 20 / 30 the general GEMM path instead of the thin action-head kernels."""
    spec, ncfg, osys, state, lrn = build(dev, E=E, U=1, T=T, P=1, M=1, scenario=scenario)
    rec = {}
    _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
    _, losses = lrn.update_step()
    sync()
    assert (lrn.traj["action"].cpu().numpy() == rec["traj"][0]["action"]).all(), "sampled actions differ"
    assert (lrn.traj["reward"].cpu().numpy() == rec["traj"][0]["reward"]).all()
    assert rel_err(lrn.traj["value"].cpu().numpy(), rec["traj"][0]["value"]) < 1e-4
    assert rel_err(lrn.traj["log_prob"].cpu().numpy(), rec["traj"][0]["log_prob"]) < 1e-4
    li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
    for name in ("value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy"):
        ref, got = infos[0][name], float(li[name][0, 0])
        assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (name, got, ref)
    gp, ap = lrn.get_params()
    for new, ref in ((gp, state["guider_params"]), (ap, state["actor_params"])):
        for name, r in ref.items():
            assert np.abs(new[name].cpu().numpy() - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-3), name
