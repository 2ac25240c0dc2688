"""-m gpu: edge cases of the env kernels and the learner — maximum supported sizes, single env / single step, ragged batch sizes —
against the oracle; plus (CPU) the C ABI's argument validation, which returns before anything touches the device."""
import ctypes as C

import numpy as np
import pytest
import torch

from magpo_b200 import _lib as L
from oracle import lbf as olbf
from oracle import prng as oprng
from oracle import rware as orw

from gpu_util import dt, sync, u32


def _null_ts():
    return L.TimeStep()


def test_abi_rejects_bad_arguments_without_a_gpu():
    lib = L.lib()
    s = C.c_void_p(0)
    lbf = L.LbfCfg(8, 2, 2, 2, 2, 1, 100, 1)
    # null pointers / negative sizes
    assert lib.magpo_lbf_reset(s, None, 4, None, L.LbfState(), _null_ts()) == L.ERR_ARG
    assert lib.magpo_lbf_step(s, C.byref(lbf), -1, None, L.LbfState(), _null_ts()) == L.ERR_ARG
    assert lib.magpo_rware_step(s, None, 4, None, L.RwareState(), _null_ts()) == L.ERR_ARG
    assert lib.magpo_coordsum_step(s, None, 4, None, L.CoordSumState(), _null_ts()) == L.ERR_ARG
    # empty batches are a no-op
    keys = (C.c_uint32 * 2)()
    assert lib.magpo_lbf_reset(s, C.byref(lbf), 0, keys, L.LbfState(), _null_ts()) == L.OK
    # configurations beyond the kernels' limits
    big = L.LbfCfg(17, 2, 2, 2, 2, 1, 100, 1)
    assert lib.magpo_lbf_reset(s, C.byref(big), 4, keys, L.LbfState(), _null_ts()) == L.ERR_UNSUPPORTED
    many = L.LbfCfg(8, 2, 9, 2, 2, 1, 100, 1)
    assert lib.magpo_lbf_reset(s, C.byref(many), 4, keys, L.LbfState(), _null_ts()) == L.ERR_UNSUPPORTED
    assert lib.magpo_rware_num_shelves(C.byref(L.RwareCfg(8, 1, 3, 4, 1, 4, 500))) == 32
    assert lib.magpo_rware_num_shelves(C.byref(L.RwareCfg(8, 2, 3, 4, 1, 4, 500))) == 80
    assert lib.magpo_rware_num_shelves(C.byref(L.RwareCfg(8, 9, 9, 4, 1, 4, 500))) == L.ERR_UNSUPPORTED  # 83 x 28 cells
    assert lib.magpo_rware_num_shelves(C.byref(L.RwareCfg(8, 1, 3, 9, 1, 4, 500))) == L.ERR_UNSUPPORTED  # 9 agents
    assert lib.magpo_rware_num_shelves(C.byref(L.RwareCfg(0, 1, 3, 4, 1, 4, 500))) == L.ERR_ARG
    bad_net = L.NetCfg(4, 200, 5, 64, 1, 1, 128, 1, 0.8, 500)  # obs_dim beyond 128
    assert lib.magpo_param_count(C.byref(bad_net), 0) < 0


def _run_env(dev, mod, spec, env, B, steps, seed, act_fn):
    from magpo_b200.learner import alloc_timestep

    A, a, d = spec.num_agents, spec.action_dim, spec.obs_dim
    keys = oprng.split(oprng.prng_key(seed), B)
    ostate, ots = mod.reset(spec, keys)
    st, ts = env.alloc_state(B, dev), alloc_timestep(B, A, d, a, dev)
    s = L.stream_ptr()
    kd = u32(keys, dev)
    L.call(env.reset_fn, s, C.byref(env.cfg), B, L.ptr(kd), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
    rng = np.random.default_rng(seed)
    for t in range(steps):
        act = act_fn(rng, ots, B, A, a)
        ostate, ots = mod.step(spec, ostate, act)
        ad = dt(act, dev)
        L.call(env.step_fn, s, C.byref(env.cfg), B, L.ptr(ad), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
        sync()
        assert (ts["agents_view"].cpu().numpy() == ots["observation"]["agents_view"]).all(), t
        assert (ts["action_mask"].cpu().numpy().astype(bool) == ots["observation"]["action_mask"]).all(), t
        assert (ts["reward"].cpu().numpy() == ots["reward"]).all(), t
        assert (ts["step_type"].cpu().numpy() == ots["step_type"]).all(), t
        assert (ts["episode_return"].cpu().numpy() == ots["extras"]["episode_metrics"]["episode_return"]).all(), t
    assert (st["key"].cpu().numpy().view(np.uint32) == ostate["env_state"]["key"]).all()


def _legal_mostly(rng, ots, B, A, a):
    m = ots["observation"]["action_mask"]
    act = rng.integers(0, a, size=(B, A)).astype(np.int32)
    legal = np.take_along_axis(m, act[..., None].astype(np.int64), -1)[..., 0]
    act = np.where(legal | (rng.random((B, A)) < 0.1), act, 0).astype(np.int32)
    if a == 6:
        act = np.where(m[..., 5] & (rng.random((B, A)) < 0.6), 5, act).astype(np.int32)
    return act


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 5, 33])
def test_lbf_maximum_size_and_ragged_batches(dev, B):
    """16 x 16 grid (256 cells: 8 per lane in the generator's scans), 8 agents, 8 food items, levels up to 3, no forced co-operation."""
    from magpo_b200.learner import LbfVec

    kw = dict(grid_size=16, fov=3, num_agents=8, num_food=8, max_agent_level=3, force_coop=False, time_limit=40)
    _run_env(dev, olbf, olbf.LbfSpec(**kw), LbfVec(**kw), B, 90, 3, _legal_mostly)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 7])
def test_rware_maximum_size(dev, B):
    """29 x 16 grid (464 cells, 224 shelves), 8 agents, 16 requests, sensor range 2 (183-wide views: env kernel only)."""
    from magpo_b200.learner import RwareVec

    kw = dict(column_height=8, shelf_rows=3, shelf_columns=5, num_agents=8, sensor_range=2, request_queue_size=16, time_limit=30)
    spec = orw.RwareSpec(**kw)
    assert spec.grid_size == (29, 16)
    _run_env(dev, orw, spec, RwareVec(**kw), B, 70, 4,
             lambda rng, ots, B_, A, a: np.where(rng.random((B_, A)) < 0.5, 1, rng.integers(0, a, size=(B_, A))).astype(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("E,U,T,M", [(1, 1, 1, 1), (1, 1, 5, 1), (2, 1, 3, 2), (3, 3, 4, 1)])
def test_learner_smallest_shapes(dev, E, U, T, M):
    """One env, one step, one minibatch; odd slot counts: the update still matches the oracle."""
    from magpo_b200.learner import LbfVec, MagpoLearner, SystemConfig
    from oracle import learner as olr
    from oracle import nets as onets

    from gpu_util import rel_err

    kw = olbf.SCENARIOS["2s-8x8-2p-2f-coop"]
    spec = olbf.LbfSpec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=2, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=7)
    lrn = MagpoLearner(LbfVec(**kw), SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=2, num_minibatches=M),
                       device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(7), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    for it in range(2):
        rec = {}
        _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
        _, losses = lrn.update_step()
        sync()
        for u in range(U):
            assert (lrn.traj["action"].cpu().numpy()[:, u * E:(u + 1) * E] == rec["traj"][u]["action"]).all(), (it, u)
            assert rel_err(lrn.traj["value"].cpu().numpy()[:, u * E:(u + 1) * E], rec["traj"][u]["value"]) < 2e-4
        li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
        for name in ("value_loss", "guider_loss", "entropy"):
            ref, got = infos[0][name], float(li[name][0, 0])
            assert abs(got - ref) <= 3e-4 * max(1.0, abs(ref)), (it, name, got, ref)
