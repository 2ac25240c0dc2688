"""-m gpu: the bench workload at BASELINE.json's full size (configs[1]: LBF 2s-8x8-2p-2f-coop, 4096 envs x 2 slots, T=128) checked through
size-independent properties — the oracle cannot run this size in reasonable time:
  * determinism: two learners from the same keys produce bit-identical trajectories and, after an update, bit-identical parameters;
  * eager rollout == CUDA-graph replay of the rollout, bit for bit;
  * env invariants over 1 M env-steps: legal masks, rewards and returns in [0, 1], step counts below the time limit, episode
    lengths consistent with the terminal flags, observations inside the value ranges of the VectorObserver;
  * GAE identities: targets - advantages == values exactly; the advantage of the last step equals its TD error;
  * sampled actions are always legal under the action mask;
  * differentiating the minibatch in env chunks gives the same update as one pass (fp32 summation order only)."""
import numpy as np
import pytest
import torch

from magpo_b200 import init as minit
from magpo_b200.learner import LbfVec, MagpoLearner, SystemConfig

pytestmark = pytest.mark.gpu

E, U, T = 4096, 2, 128


def _make(dev, chunk=4096, graph=True, seed=42):
    env = LbfVec()
    lrn = MagpoLearner(env, SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, chunk_envs=chunk), device=dev, graph_rollout=graph)
    lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 0), minit.init_actor(env.obs_dim, env.action_dim, 1))
    env_keys, step_key, _ = minit.setup_keys(seed, 1, U, E, dev)
    lrn.reset(env_keys[0], step_key)
    return lrn


def test_full_size_determinism_graph_replay_and_invariants(dev):
    a, b = _make(dev, chunk=1024), _make(dev, chunk=1024, graph=False)  # the update workspace is 11 GB per learner at 1024-env chunks
    for it in range(3):  # a: eager, capture, replay; b: always eager
        a.rollout(); b.rollout()
        a.gae(); b.gae()
        torch.cuda.synchronize()
        for k in ("action", "reward", "value", "log_prob", "agents_view", "action_mask", "done", "step_count", "episode_return",
                  "episode_length", "is_terminal_step", "last_value"):
            assert torch.equal(a.traj[k], b.traj[k]), (it, k)
        assert torch.equal(a.adv, b.adv) and torch.equal(a.policy_h, b.policy_h)
    tr = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in a.traj.items()}
    adv, targets = a.adv.clone(), a.targets.clone()
    del a, b
    torch.cuda.empty_cache()
    A, d = 2, 14
    # sampled actions are legal
    legal = torch.gather(tr["action_mask"][:T], -1, tr["action"].long()[..., None])[..., 0]
    assert bool(legal.all())
    # rewards / returns of the normalised game; step counts; NOOP always legal
    assert float(tr["reward"].min()) >= 0.0 and float(tr["reward"].max()) <= 1.0 + 1e-6
    assert float(tr["episode_return"].max()) <= 1.0 + 1e-5 and float(tr["episode_return"].min()) >= 0.0
    assert int(tr["step_count"].max()) < 100 and int(tr["step_count"].min()) >= 0
    assert bool(tr["action_mask"][..., 0].all())
    # published episode lengths are within the time limit and at least 1 wherever an episode ended
    term = tr["is_terminal_step"].bool()
    assert int(term.sum()) > 2 * E  # 384 steps of 100-step episodes: every env finished at least twice
    assert int(tr["episode_length"][term].min()) >= 1 and int(tr["episode_length"][term].max()) <= 100
    # done[t+1] is the terminal flag of step t; a terminal step resets the step count of the next observation
    assert torch.equal(tr["done"][1:].bool(), term)
    assert int(tr["step_count"][1:][term].max()) == 0
    # observation ranges: one-hot agent id, then (row', col', level) triples in [-1, 2 fov] x [-1, 2 fov] x [0, 4]
    view = tr["agents_view"]
    assert torch.equal(view[..., :A].sum(-1), torch.ones_like(view[..., 0]))
    trip = view[..., A:].reshape(*view.shape[:-1], -1, 3)
    assert float(trip[..., :2].min()) >= -1 and float(trip[..., :2].max()) <= 4 and float(trip[..., 2].min()) >= 0 and float(trip[..., 2].max()) <= 4
    # GAE identities
    assert torch.equal(targets - adv, tr["value"]) or float((targets - adv - tr["value"]).abs().max()) <= 1e-6
    nd = 1.0 - tr["done"][T].float()[:, None]
    delta_last = tr["reward"][T - 1] + 0.99 * tr["last_value"] * nd - tr["value"][T - 1]
    assert float((adv[T - 1] - delta_last).abs().max()) <= 1e-6
    # a full update from identical states is bit-reproducible
    del tr, adv, targets
    torch.cuda.empty_cache()
    a2, b2 = _make(dev, chunk=1024), _make(dev, chunk=1024)
    a2.update_step(); b2.update_step()
    torch.cuda.synchronize()
    # cross-CTA fp32 reductions (atomics, TMA reduce-adds) add in a run-dependent order: last bits only
    assert torch.equal(a2.guider, b2.guider) or float((a2.guider - b2.guider).abs().max()) <= 1e-6
    assert float((a2.actor - b2.actor).abs().max()) <= 1e-6
    assert torch.equal(a2.key, b2.key)
    del a2, b2
    torch.cuda.empty_cache()


def test_full_size_chunked_update_equals_single_pass(dev):
    one, chunked = _make(dev, chunk=4096), _make(dev, chunk=1024)
    _, l1 = one.update_step()
    _, l2 = chunked.update_step()
    torch.cuda.synchronize()
    assert torch.equal(one.traj["action"], chunked.traj["action"])
    assert float((l1 - l2).abs().max()) <= 1e-5 * max(1.0, float(l1.abs().max()))
    # parameters move by ~lr per Adam step; the two differ by summation order only
    dg = float((one.guider - chunked.guider).abs().max())  # noqa: E501
    da = float((one.actor - chunked.actor).abs().max())
    assert dg <= 2e-5 and da <= 2e-5, (dg, da)
    del one, chunked
    torch.cuda.empty_cache()
