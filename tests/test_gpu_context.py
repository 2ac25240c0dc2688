"""-m gpu: the per-device MagpoContext (include/magpo_b200.h) — no process-global state behind the ABI. Two learners of different
shapes (different workspace plans, different tensor-core weight images, their own side streams), interleaved call by call in one
process, must reproduce their solo runs (rollouts bit for bit, updates to the fp32 reduction order); a context refuses a call made with another device current; a single-rank
communicator is a no-op."""
import ctypes as C

import pytest
import torch

from magpo_b200 import _lib as L
from magpo_b200 import init as minit
from magpo_b200.learner import CoordSumVec, LbfVec, MagpoLearner, SystemConfig

pytestmark = pytest.mark.gpu


def _make(kind, dev):
    if kind == "a":
        env, sysc = CoordSumVec(3, 10, 100, 30), SystemConfig(num_envs=24, update_batch_size=2, rollout_length=16, ppo_epochs=2, num_minibatches=2)
    else:
        env, sysc = LbfVec(), SystemConfig(num_envs=80, update_batch_size=1, rollout_length=24, ppo_epochs=2, num_minibatches=2)
    lrn = MagpoLearner(env, sysc, device=dev)
    lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 3), minit.init_actor(env.obs_dim, env.action_dim, 4))
    env_keys, step_key, _ = minit.setup_keys(11, 1, sysc.update_batch_size, sysc.num_envs, dev)
    lrn.reset(env_keys[0], step_key)
    return lrn


def _snapshot(lrn):
    return dict(guider=lrn.guider.clone(), actor=lrn.actor.clone(), action=lrn.traj["action"].clone(), key=lrn.key.clone(),
                g_mu=lrn.g_mu.clone())


def test_two_interleaved_learners_match_their_solo_runs(dev):
    solo = {}
    for kind in ("a", "b"):
        lrn = _make(kind, dev)
        for _ in range(2):
            lrn.update_step()
        torch.cuda.synchronize()
        solo[kind] = _snapshot(lrn)
        del lrn
    la, lb = _make("a", dev), _make("b", dev)
    assert la.ctx.value != lb.ctx.value
    for _ in range(2):
        la.rollout(); lb.rollout(); lb.gae(); la.gae()
        for p in range(2):
            la.epoch_indices(p == 0); lb.epoch_indices(p == 0)
            for m in range(2):
                la.minibatch_grads(m); lb.minibatch_grads(m)   # B's weight images are registered while A's gradients are pending
                lb.apply_grads(); la.apply_grads()
    torch.cuda.synchronize()
    # integer / rollout products are bit-identical; the update's gradients are reduced across CTAs with fp32 atomics / TMA reduce-adds
    # whose order differs from run to run (two SOLO runs differ by the same last bits), so parameters are compared at 2e-6 of
    # their scale — a stale or foreign weight image, the failure this test is after, shows up at 1e-2
    for kind, lrn in (("a", la), ("b", lb)):
        got = _snapshot(lrn)
        for k in ("action", "key"):
            assert torch.equal(solo[kind][k], got[k]), (kind, k)
        for k in ("guider", "actor", "g_mu"):
            v, w = solo[kind][k], got[k]
            assert float((v - w).abs().max()) <= 2e-6 * max(float(v.abs().max()), 1e-3), (kind, k, float((v - w).abs().max()))


def test_context_is_bound_to_its_device_and_required(dev):
    lib = L.lib()
    lrn = _make("a", dev)
    # a NULL context is refused by every stateful entry point
    rc = lib.magpo_actor_step(None, L.stream_ptr(), C.byref(lrn.c_net), 1, L.ptr(lrn.actor), L.ptr(lrn.traj["agents_view"][0]),
                              L.ptr(lrn.traj["done"][0]), L.ptr(lrn.policy_h), L.ptr(lrn.workspace), C.c_size_t(lrn.ws_bytes))
    assert rc == L.ERR_ARG
    assert lib.magpo_context_create(10_000, C.byref(L.vp())) == L.ERR_ARG


def test_single_rank_communicator_is_a_no_op(dev):
    from magpo_b200.comm import NcclComm

    comm = NcclComm(0, 1, dev)
    t = torch.arange(8, dtype=torch.float32, device=dev)
    comm.allreduce_sum(t)
    comm.barrier()
    assert torch.equal(t.cpu(), torch.arange(8, dtype=torch.float32))
    lrn = _make("a", dev)
    ref = _make("a", dev)
    comm.attach(lrn)  # reduce_grads is passed, the single-rank communicator changes nothing
    lrn.update_step(); ref.update_step()
    torch.cuda.synchronize()
    assert torch.equal(lrn.traj["action"], ref.traj["action"])
    for a_, b_ in ((lrn.guider, ref.guider), (lrn.actor, ref.actor)):  # fp32 reduction order only (see above)
        assert float((a_ - b_).abs().max()) <= 2e-6 * float(b_.abs().max())
    comm.close()
