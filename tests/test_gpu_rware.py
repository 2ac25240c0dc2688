"""-m gpu: RobotWarehouse (configs c3/c4 of SURVEY.md §8) — the env-step kernel bit-exact against oracle/rware.py, and the whole
`_update_step` on it (obs_dim 75: the general, non-thin observation-embedding path) against the CPU oracle. The dynamics are a
restatement of the un-vendored jumanji 1.1.0 package (parity unpinned, see oracle/rware.py): this proves CUDA == restatement."""
import ctypes as C

import numpy as np
import pytest
import torch

from magpo_b200 import _lib as L
from magpo_b200.learner import MagpoLearner, RwareVec, SystemConfig, alloc_timestep
from oracle import learner as olr
from oracle import nets as onets
from oracle import prng as oprng
from oracle import rware as orw

from gpu_util import as_u32, dt, rel_err, sync, u32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scenario,time_limit", [("tiny-4ag", 60), ("tiny-2ag", 500), ("small-4ag", 45), ("medium-6ag", 40)])
def test_rware_bit_exact(dev, scenario, time_limit):
    kw = dict(orw.SCENARIOS[scenario], time_limit=time_limit)
    spec = orw.RwareSpec(**kw)
    env = RwareVec(**kw)
    assert env.num_shelves == len(spec.shelf_cells)
    B, A, a, d = 64, spec.num_agents, spec.action_dim, spec.obs_dim
    keys = oprng.split(oprng.prng_key(5), B)
    ostate, ots = orw.reset(spec, keys)
    st = env.alloc_state(B, dev)
    ts = alloc_timestep(B, A, d, a, dev)
    s = L.stream_ptr()
    kd = u32(keys, dev)
    L.call("magpo_rware_reset", s, C.byref(env.cfg), B, L.ptr(kd), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
    rng = np.random.default_rng(0)

    def check(tag):
        sync()
        b = ostate["env_state"]
        for k in ("grid", "agent_pos", "agent_dir", "shelf_pos", "request_queue", "step_count"):
            assert (st[k].cpu().numpy() == b[k]).all(), (tag, k)
        for k in ("agent_carry", "shelf_req", "action_mask"):
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
    # plant deliveries: a carrying agent one cell above a goal, facing down, holding a requested shelf (state edited on both sides)
    b = ostate["env_state"]
    H, W = spec.grid_size
    for e in range(0, B, 4):
        sid = int(b["request_queue"][e, 0])
        ox, oy = b["shelf_pos"][e, sid]
        gx, gy = H - 2, W // 2 - 1 + (e // 4) % 2
        if b["grid"][e, 1, gx, gy] or b["grid"][e, 1, gx + 1, gy]:
            continue
        ax, ay = b["agent_pos"][e, 0]
        b["grid"][e, 1, ax, ay] = 0
        b["grid"][e, 1, gx, gy] = 1
        b["grid"][e, 0, ox, oy] = 0
        b["grid"][e, 0, gx, gy] = sid + 1
        b["shelf_pos"][e, sid] = (gx, gy)
        b["agent_pos"][e, 0] = (gx, gy)
        b["agent_dir"][e, 0] = 2
        b["agent_carry"][e, 0] = True
        b["action_mask"][e] = orw._action_mask(spec, b["grid"][e], b["agent_pos"][e], b["agent_dir"][e], b["agent_carry"][e])
    for k in ("grid", "agent_pos", "agent_dir", "shelf_pos"):
        st[k].copy_(dt(b[k], dev))
    for k in ("agent_carry", "action_mask"):
        st[k].copy_(dt(b[k].astype(np.uint8), dev))
    n_rew = n_coll = 0
    for step in range(130):
        m = ots["observation"]["action_mask"] if step else b["action_mask"]
        act = rng.integers(0, a, size=(B, A)).astype(np.int32)  # includes masked FORWARDs (must become NOOP)
        act = np.where(rng.random((B, A)) < 0.5, 1, act).astype(np.int32)  # move a lot: collisions, shelf transport
        if step == 0:
            act[::4, 0] = 1  # the planted agents step onto the goal
        ostate, ots = orw.step(spec, ostate, act)
        n_rew += int((ots["reward"][:, 0] > 0).sum())
        n_coll += int(((ots["step_type"] == 2) & (ots["extras"]["real_next_obs"]["step_count"][:, 0] < time_limit)).sum())
        ad = dt(act, dev)
        L.call("magpo_rware_step", s, C.byref(env.cfg), B, L.ptr(ad), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
        check(f"step {step}")
    assert n_rew > 0, "no delivery was covered"
    if A >= 4:
        assert n_coll > 0, "no collision termination was covered"


def build(dev, E, U, T, P, M, scenario, time_limit=500, seed=42):
    kw = dict(orw.SCENARIOS[scenario], time_limit=time_limit)
    spec = orw.RwareSpec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=seed)
    sysc = SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    lrn = MagpoLearner(RwareVec(**kw), sysc, device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(seed), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    return spec, ncfg, osys, state, lrn


@pytest.mark.parametrize("scenario,E,T", [("tiny-4ag", 6, 16), ("small-4ag", 4, 10), ("tiny-2ag", 4, 12)])
def test_rware_update_step_matches_oracle(dev, scenario, E, T):
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


def test_rware_rollouts_carry_state_across_resets(dev):
    spec, ncfg, osys, state, lrn = build(dev, E=4, U=1, T=40, P=1, M=1, scenario="tiny-4ag", time_limit=25)
    for it in range(3):
        rec = {}
        olr.update_step(state, spec, ncfg, osys, record=rec)
        lrn.update_step()
        sync()
        assert (lrn.traj["action"].cpu().numpy() == rec["traj"][0]["action"]).all(), it
        assert (lrn.traj["reward"].cpu().numpy() == rec["traj"][0]["reward"]).all(), it
        assert rel_err(lrn.traj["value"].cpu().numpy(), rec["traj"][0]["value"]) < 2e-4, it
