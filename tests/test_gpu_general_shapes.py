"""-m gpu: network.net_config beyond the default (64, 1, 1) — embed_dim 32 / 64 / 128, n_head 1 / 2 / 4, n_block 1..3 (the reference loops
over blocks and heads, sable_network.py:111-119,286-294, retention.py:229-259,281-287; its tuned runs use these shapes,
experiment_data/params.csv:61-103). The general path (csrc/generic*.cu) against the oracle, which implements the same loops: guider forward
and the full minibatch gradient with every parameter moved off its init value (non-zero SwiGLU), then whole update steps — rollout through
the layer-by-layer get_actions with bit-identical sampled actions — on the tuned shapes of tiny-4ag (128, 2, 3) and LBF 2s-8x8 (32, 4, 2)."""
import ctypes as C

import numpy as np
import pytest
import torch

from magpo_b200 import _lib as L
from magpo_b200.learner import CoordSumVec, LbfVec, MagpoLearner, NetworkConfig, RwareVec, SystemConfig
from oracle import coordsum as ocs, lbf as olbf, learner as olr, nets as onets, prng as oprng, rware as orw

import test_gpu_networks as tgn
from gpu_util import from_time_major, rel_err, sync

pytestmark = pytest.mark.gpu

SHAPES = [(128, 2, 3), (32, 4, 2), (64, 2, 2), (64, 1, 2), (32, 1, 1), (128, 4, 1)]


# n_head = 4 leaves GroupNorm groups of hs / n_head = 2 (embed_dim 32) or 8 (embed_dim 128) features: the normalisation of a 2-element
# group amplifies rounding by up to 1 / sqrt(eps) = 1000, and the fp32 ORACLE itself then sits 2-4e-4 away from its float64 run
# (tools/diag_general.py: (32, 4, 1) o32-vs-o64 1.8e-4 / 3.0e-4, CUDA-vs-o64 3.7e-4 / 6.2e-4; every other shape 3-5e-7). Those shapes
# are therefore compared against the float64 oracle with the fp32 oracle's own deviation as the yardstick.
def _ill_conditioned(shape):
    return shape[1] == 4


@pytest.mark.parametrize("shape", SHAPES)
def test_general_guider_forward(dev, shape):
    A, d, a, T, N = 3, 9, 7, 10, 5
    cfg, net, gp, ap, (gt, ng, gflat), _ = tgn.setup_nets(A, d, a, dev, shape=shape)
    mb = tgn.make_case(1, 1, N, T, A, d, a, shape=shape)
    mbs, keep = tgn.device_minibatch(mb, T, A, dev)
    ws, nbytes = tgn.workspace(net, T, N, dev)
    value, logits = torch.zeros(T, N, A, device=dev), torch.zeros(T, N, A, a, device=dev)
    L.call("magpo_guider_forward", L.context(), L.stream_ptr(), C.byref(net.c_struct()), L.ptr(gflat), mbs, L.ptr(value), L.ptr(logits),
           L.ptr(ws), C.c_size_t(nbytes))
    ref = {}
    for dt_ in (torch.float32, torch.float64):
        v, _, _, l = onets.sable_apply(onets.to_torch(gp, dt_), cfg, torch.tensor(mb["obs"], dtype=dt_), torch.tensor(mb["action_mask"]),
                                       torch.tensor(mb["step_count"]), torch.tensor(mb["action"]),
                                       tuple(torch.tensor(h, dtype=dt_) for h in mb["prev_hstates"]), torch.tensor(mb["done"]), T)
        ref[dt_] = (v.numpy(), l.numpy())
    sync()
    lg, legal = from_time_major(logits.cpu().numpy()), mb["action_mask"]
    v64, l64 = ref[torch.float64]
    ev, el = rel_err(from_time_major(value.cpu().numpy()), v64), rel_err(lg[legal], l64[legal])
    ov, ol = rel_err(ref[torch.float32][0], v64), rel_err(ref[torch.float32][1][legal], l64[legal])
    print(shape, "cuda vs o64", ev, el, "o32 vs o64", ov, ol)
    assert ev < max(1e-4, 4 * ov) and el < max(1e-4, 4 * ol), (ev, el, ov, ol)
    if not _ill_conditioned(shape):
        assert ev < 1e-4 and el < 1e-4
    assert (lg[~legal] == np.finfo(np.float32).min).all()


@pytest.mark.parametrize("shape,Ns", [((128, 2, 3), 4), ((32, 4, 2), 6), ((64, 2, 2), 5), ((64, 1, 2), 40), ((128, 4, 1), 40)])
def test_general_minibatch_grads(dev, shape, Ns):
    """every gradient tensor (per block, per head) against the fp64 oracle; Ns = 40: >= 256 token rows, the tensor-core GEMMs engage"""
    tgn.test_minibatch_grads(dev, 3, 9, 7, 8, Ns, 2, 5e-3 if _ill_conditioned(shape) else 2e-4, shape=shape, yardstick=_ill_conditioned(shape))


def _build(dev, kind, shape, E, U, T, P, M, seed=42):
    if kind == "rware":
        kw = orw.SCENARIOS["tiny-4ag"]
        spec, vec = orw.RwareSpec(**kw), RwareVec(**kw)
    elif kind == "lbf":
        kw = olbf.SCENARIOS["2s-8x8-2p-2f-coop"]
        spec, vec = olbf.LbfSpec(**kw), LbfVec(**kw)
    else:
        kw = ocs.SCENARIOS["5x20-80-v0"]
        spec, vec = ocs.CoordSumSpec(**kw), CoordSumVec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim, embed_dim=shape[0], n_head=shape[1], n_block=shape[2])
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=seed)
    net = NetworkConfig(vec.num_agents, vec.obs_dim, vec.action_dim, vec.time_limit, embed_dim=shape[0], n_head=shape[1], n_block=shape[2])
    lrn = MagpoLearner(vec, SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M), device=dev, net=net)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(seed), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    return spec, ncfg, osys, state, lrn


# the reference's tuned shapes: rware tiny-4ag (n_embd 128, n_head 2, n_block 3), lbf 2s-8x8-2p-2f-coop (32, 4, 2), coordsum 5x20 (64, 2, 2)
@pytest.mark.parametrize("kind,shape,E,T", [("rware", (128, 2, 3), 4, 10), ("lbf", (32, 2, 2), 8, 16), ("coordsum", (64, 2, 2), 4, 12)])
def test_general_update_steps_match_oracle(dev, kind, shape, E, T):
    """(the LBF row of params.csv is (32, 4, 2); its 2-feature GroupNorm groups make sampled actions sensitive to the last bits of ANY
    fp32 implementation, see _ill_conditioned — the bit-exact trajectory comparison runs on (32, 2, 2), the tuned shape itself in
    test_tuned_lbf_shape_rollout_statistics below)"""
    spec, ncfg, osys, state, lrn = _build(dev, kind, shape, E, 2, T, 2, 2)
    assert lrn.hs["encoder"].shape == (2 * E, shape[1], shape[2], shape[0] // shape[1], shape[0] // shape[1])
    for it in range(2):  # the second update rolls out from carried multi-block / multi-head states (and replays the CUDA graph)
        rec = {}
        _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
        _, losses = lrn.update_step()
        sync()
        for u in range(2):
            sl = slice(u * E, (u + 1) * E)
            assert (lrn.traj["action"].cpu().numpy()[:, sl] == rec["traj"][u]["action"]).all(), (it, "sampled actions differ")
            assert (lrn.traj["reward"].cpu().numpy()[:, sl] == rec["traj"][u]["reward"]).all(), it
            assert rel_err(lrn.traj["value"].cpu().numpy()[:, sl], rec["traj"][u]["value"]) < 1e-4
            assert rel_err(lrn.traj["log_prob"].cpu().numpy()[:, sl], rec["traj"][u]["log_prob"]) < 1e-4
        li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
        k = 0
        for p in range(2):
            for m in range(2):
                for name in ("value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy"):
                    ref, got = infos[k][name], float(li[name][p, m])
                    assert abs(got - ref) <= 2e-4 * max(1.0, abs(ref)), (it, p, m, name, got, ref)
                k += 1
        gp, ap = lrn.get_params()
        for new, ref in ((gp, state["guider_params"]), (ap, state["actor_params"])):
            for name, r in ref.items():  # |dp| <= 1e-4 |p| + 0.1 lr per update (the 2nd update carries the 1st's deviation)
                d_ = np.abs(new[name].cpu().numpy() - r)
                assert (d_ <= (it + 1) * (1e-4 * np.abs(r) + 0.1 * osys.actor_lr)).all(), (it, name, float(d_.max()))
        hs = lrn.sable_hidden_state()
        for name, ref in zip(("encoder", "decoder_self", "decoder_cross"), state["slots"][0]["hstates"]["sable"]):
            assert rel_err(hs[name].cpu().numpy()[:E], ref) < 1e-4, (it, name)


def test_tuned_lbf_shape_rollout_statistics(dev):
    """The reference's tuned LBF shape (32, 4, 2): rollout + update run, and the rollout agrees with the oracle to what its conditioning
    allows — the first env step (identical states) is compared exactly on everything but the sampled actions of near-tied logits, the whole
    rollout on the fraction of identical actions."""
    spec, ncfg, osys, state, lrn = _build(dev, "lbf", (32, 4, 2), 8, 2, 16, 2, 2)
    rec = {}
    olr.update_step(state, spec, ncfg, osys, record=rec)
    _, losses = lrn.update_step()
    sync()
    same, total = 0, 0
    for u in range(2):
        sl = slice(u * 8, (u + 1) * 8)
        act = lrn.traj["action"].cpu().numpy()[:, sl]
        same += int((act == rec["traj"][u]["action"]).sum())
        total += act.size
        assert rel_err(lrn.traj["value"].cpu().numpy()[0, sl], rec["traj"][u]["value"][0]) < 5e-3
    print("identical sampled actions:", same, "/", total)
    assert same >= 0.9 * total
    assert torch.isfinite(losses).all()


def test_system_entry_with_a_tuned_network_shape(dev):
    """`python -m magpo_b200.rec_magpo network.net_config.n_block=2 ...`: the config reaches the kernels, the state pytree has the
    [.., n_head, n_block, hs, hs] Sable states and `learn` trains."""
    from magpo_b200 import init as minit
    from magpo_b200 import rec_magpo as rm
    from magpo_b200.config import compose

    cfg = compose("default/rec_magpo", ["arch.num_envs=6", "system.rollout_length=8", "system.ppo_epochs=2", "system.num_updates=2",
                                        "arch.num_evaluation=1", "system.total_timesteps=~", "network.net_config.n_block=2",
                                        "network.net_config.embed_dim=32", "network.net_config.n_head=2"])
    cfg.system.num_updates_per_eval = 2
    env = rm.make_env(cfg)
    key, _, ak, nk = minit.split(minit.prng_key(42), 4, dev)
    learn, net, state = rm.learner_setup(env, (key, ak, nk), cfg, device=dev)
    assert state.hstates.sable_hidden_state.encoder.shape == (1, 2, 6, 2, 2, 16, 16)
    assert state.params.guider_params["decoder/decoder_block_1/retn2/retention_heads_1/w_v"].shape == (1, 2, 32, 16)
    out = learn(state)
    assert torch.isfinite(out.train_metrics["total_loss"]).all()
