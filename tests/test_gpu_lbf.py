"""-m gpu: LevelBasedForaging (config c2 of SURVEY.md §8) — the env-step kernel bit-exact against oracle/lbf.py, and the whole
`_update_step` on it against the CPU oracle. The LBF dynamics are a restatement of the un-vendored jumanji 1.1.0 package (parity
unpinned, see oracle/lbf.py); what this file proves is CUDA == restatement, bit for bit."""
import ctypes as C

import numpy as np
import pytest
import torch

from magpo_b200 import _lib as L
from magpo_b200.learner import LbfVec, MagpoLearner, SystemConfig, alloc_timestep
from oracle import lbf as olbf
from oracle import learner as olr
from oracle import nets as onets
from oracle import prng as oprng

from gpu_util import as_u32, dt, rel_err, sync, u32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scenario,rows", [("2s-8x8-2p-2f-coop", True), ("2s-10x10-3p-3f", True), ("15x15-4p-5f", True),
                                            ("8x8-2p-2f-coop", False), ("15x15-3p-5f", False)])
def test_lbf_bit_exact(dev, scenario, rows, monkeypatch):
    monkeypatch.setattr(olbf, "AGENT_MASK_CLEARS_ROWS", rows)
    kw = olbf.SCENARIOS[scenario]
    spec = olbf.LbfSpec(**kw)
    env = LbfVec(**kw, agent_mask_rows=rows)
    B, A, a, d = 96, spec.num_agents, spec.action_dim, spec.obs_dim
    keys = oprng.split(oprng.prng_key(11), B)
    ostate, ots = olbf.reset(spec, keys)
    st = env.alloc_state(B, dev)
    ts = alloc_timestep(B, A, d, a, dev)
    s = L.stream_ptr()
    kd = u32(keys, dev)
    L.call("magpo_lbf_reset", s, C.byref(env.cfg), B, L.ptr(kd), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
    rng = np.random.default_rng(0)

    def check(tag):
        sync()
        b = ostate["env_state"]
        for k in ("agent_pos", "agent_level", "food_pos", "food_level", "step_count"):
            assert (st[k].cpu().numpy() == b[k]).all(), (tag, k)
        for k in ("agent_loading", "food_eaten"):
            assert (st[k].cpu().numpy().astype(bool) == b[k]).all(), (tag, k)
        assert (as_u32(st["key"]) == b["key"]).all(), tag
        assert (as_u32(st["metrics_key"]) == ostate["key"]).all(), tag
        for k1, k2 in (("running_return", "running_count_episode_return"), ("running_length", "running_count_episode_length"),
                       ("episode_return", "episode_return"), ("episode_length", "episode_length")):
            assert (st[k1].cpu().numpy() == ostate[k2]).all(), (tag, k1)
        ob = ots["observation"]
        assert (ts["agents_view"].cpu().numpy() == ob["agents_view"]).all(), tag
        assert (ts["action_mask"].cpu().numpy().astype(bool) == ob["action_mask"]).all(), tag
        assert (ts["step_count"].cpu().numpy() == ob["step_count"]).all(), tag
        assert (ts["step_type"].cpu().numpy() == ots["step_type"]).all(), tag
        assert (ts["reward"].cpu().numpy() == ots["reward"]).all(), tag
        assert (ts["discount"].cpu().numpy() == ots["discount"]).all(), tag
        ex = ots["extras"]
        assert (ts["next_agents_view"].cpu().numpy() == ex["real_next_obs"]["agents_view"]).all(), tag
        assert (ts["next_step_count"].cpu().numpy() == ex["real_next_obs"]["step_count"]).all(), tag
        em = ex["episode_metrics"]
        assert (ts["episode_return"].cpu().numpy() == em["episode_return"]).all(), tag
        assert (ts["episode_length"].cpu().numpy() == em["episode_length"]).all(), tag
        assert (ts["is_terminal_step"].cpu().numpy().astype(bool) == em["is_terminal_step"]).all(), tag

    check("reset")
    n_term = n_rew = 0
    for step in range(240):  # crosses the time limit twice; early terminations (all food eaten) auto-reset in between
        m = ots["observation"]["action_mask"]
        # mostly legal actions, LOAD whenever legal half of the time (so food gets eaten), some illegal moves (must be refused)
        act = rng.integers(0, a, size=(B, A)).astype(np.int32)
        legal = np.take_along_axis(m, act[..., None].astype(np.int64), -1)[..., 0]
        act = np.where(legal | (rng.random((B, A)) < 0.1), act, 0).astype(np.int32)
        act = np.where(m[..., 5] & (rng.random((B, A)) < 0.6), 5, act).astype(np.int32)
        ostate, ots = olbf.step(spec, ostate, act)
        n_term += int(((ots["step_type"] == 2) & (ots["discount"][:, 0] == 0)).sum())
        n_rew += int((ots["reward"][:, 0] > 0).sum())
        ad = dt(act, dev)
        L.call("magpo_lbf_step", s, C.byref(env.cfg), B, L.ptr(ad), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
        check(f"step {step}")
    assert n_rew > 0, "the action bias never produced a reward: the test would not cover eat_food / get_reward"
    if spec.num_food <= 3:
        assert n_term > 0, "no early termination was covered"


def build(dev, E=8, U=2, T=16, P=2, M=2, scenario="2s-8x8-2p-2f-coop", seed=42, chunk=0):
    kw = olbf.SCENARIOS[scenario]
    spec = olbf.LbfSpec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=seed)
    sysc = SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M, chunk_envs=chunk)
    lrn = MagpoLearner(LbfVec(**kw), sysc, device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(seed), 4)
    allk = oprng.split(ks[0], U * E + 1)
    step_key = oprng.split(allk[0])[1]
    lrn.reset(allk[1:], step_key)
    return spec, ncfg, osys, state, lrn


@pytest.mark.parametrize("scenario,E,T", [("2s-8x8-2p-2f-coop", 8, 24), ("2s-10x10-3p-3f", 4, 12), ("15x15-4p-5f", 4, 8)])
def test_lbf_update_step_matches_oracle(dev, scenario, E, T):
    """Masked action heads (LBF's action mask is not constant), float observations with negative entries, d = 14 (fused rollout step
    kernel), 21 and 31 (per-layer rollout path)."""
    spec, ncfg, osys, state, lrn = build(dev, E=E, U=2, T=T, P=2, M=2, scenario=scenario)
    rec = {}
    _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
    _, losses = lrn.update_step()
    sync()
    for u in range(2):
        sl = slice(u * E, (u + 1) * E)
        assert (lrn.traj["action"].cpu().numpy()[:, sl] == rec["traj"][u]["action"]).all(), "sampled actions differ"
        assert (lrn.traj["reward"].cpu().numpy()[:, sl] == rec["traj"][u]["reward"]).all()
        assert (lrn.traj["agents_view"].cpu().numpy()[:T, sl] == rec["traj"][u]["obs"].astype(np.float32)).all()
        assert rel_err(lrn.traj["value"].cpu().numpy()[:, sl], rec["traj"][u]["value"]) < 1e-4
        assert rel_err(lrn.traj["log_prob"].cpu().numpy()[:, sl], rec["traj"][u]["log_prob"]) < 1e-4
    li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
    k = 0
    for p in range(2):
        for m in range(2):
            for name in ("value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy", "total_loss"):
                ref, got = infos[k][name], float(li[name][p, m])
                assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (p, m, name, got, ref)
            k += 1
    gp, ap = lrn.get_params()
    for new, ref in ((gp, state["guider_params"]), (ap, state["actor_params"])):
        for name, r in ref.items():
            assert np.abs(new[name].cpu().numpy() - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-3), name


def test_lbf_second_rollout_carries_state(dev):
    spec, ncfg, osys, state, lrn = build(dev, E=4, U=1, T=110, P=1, M=1)  # T > time_limit: crosses an auto-reset
    for it in range(3):
        rec = {}
        olr.update_step(state, spec, ncfg, osys, record=rec)
        lrn.update_step()
        sync()
        assert (lrn.traj["action"].cpu().numpy() == rec["traj"][0]["action"]).all(), it
        assert (lrn.traj["reward"].cpu().numpy() == rec["traj"][0]["reward"]).all(), it
        assert rel_err(lrn.traj["value"].cpu().numpy(), rec["traj"][0]["value"]) < 2e-4, it
