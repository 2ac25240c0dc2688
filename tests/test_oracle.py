"""CPU tests (-m "not gpu"): pin the oracle against published known-answer vectors, hand-computed cases and
internal identities (the reference ships no tests / golden vectors — SURVEY.md §4), and check the C-ABI
library loads and exports every symbol include/magpo_b200.h declares."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import coordsum as ocs
from oracle import learner as olr
from oracle import nets as onets
from oracle import prng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- PRNG
def test_threefry_random123_kats():
    # Random123 / JAX test-suite known answers
    for key, ctr, exp in (((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
                          ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
                          ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))):
        o0, o1 = prng.threefry2x32(np.array(key, np.uint32), np.uint32(ctr[0]), np.uint32(ctr[1]))
        assert (int(o0), int(o1)) == exp


def test_split_kats_both_schemes():
    k0 = prng.prng_key(0)
    assert prng.split(k0).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]  # partitionable (jax>=0.5)
    prng.PARTITIONABLE = False
    try:
        assert prng.split(k0).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]  # legacy
    finally:
        prng.PARTITIONABLE = True
    assert prng.prng_key(42).tolist() == [0, 42]


def test_uniform_gumbel_randint_permutation_properties():
    key = prng.prng_key(3)
    u = prng.uniform(key, (10000,))
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    g = prng.gumbel(key, (10000,))
    assert np.isfinite(g).all() and abs(g.mean() - 0.5772) < 0.05
    r = prng.randint(key, (5000,), 0, 30)
    assert r.min() == 0 and r.max() == 29
    for n in (1, 3, 16, 4096):
        p = prng.permutation(key, n)
        assert sorted(p.tolist()) == list(range(n))
    assert prng.permutation_rounds(16) == 1 and prng.permutation_rounds(4096) == 2
    # batched helpers == scalar loop
    keys = prng.split(key, 5)
    assert (prng.randint_batched(keys, 11, 0, 30) == np.stack([prng.randint(k, (11,), 0, 30) for k in keys])).all()


def test_golden_prng_fixture():
    with open(os.path.join(ROOT, "tests", "golden", "prng.json")) as f:
        gold = json.load(f)
    key = prng.prng_key(gold["seed"])
    assert prng.split(key, 4).tolist() == gold["split4"]
    assert prng.random_bits(key, (8,)).tolist() == gold["bits8"]
    assert prng.randint(key, (8,), 0, 30).tolist() == gold["randint8_0_30"]
    assert prng.permutation(key, 16).tolist() == gold["perm16"]


# ----------------------------------------------------------------------------- CoordSum
def _mk(B=2, **kw):
    spec = ocs.CoordSumSpec(**ocs.SCENARIOS["3x10-30-v0"])
    keys = prng.split(prng.prng_key(1), B)
    return spec, *ocs.reset(spec, keys)


def test_coordsum_hand_cases():
    spec, st, ts = _mk()
    B, A = 2, spec.num_agents
    assert ts["observation"]["agents_view"].dtype == np.int32  # AgentIDWrapper keeps int32 for CoordSum
    assert (ts["observation"]["agents_view"][:, :, :A] == np.eye(A, dtype=np.int32)).all()
    assert (ts["observation"]["agents_view"][:, :, A] == st["env_state"]["target"][:, :1]).all()
    # reward 0 when the sum misses; 2 when it matches but the modal first-action differs; 1 when it agrees
    tgt = st["env_state"]["target"][:, 0].copy()
    st["env_state"]["target"][:, 1] = tgt  # same target twice, so that the record row is consulted
    act = np.zeros((B, A), np.int32)
    act[:, 0] = np.minimum(tgt, 9)
    act[:, 1] = np.clip(tgt - act[:, 0], 0, 9)
    act[:, 2] = tgt - act[:, 0] - act[:, 1]
    assert (act.sum(1) == tgt).all()
    st, ts = ocs.step(spec, st, act)
    exp = np.where(act[:, 0] == 0, 1.0, 2.0)  # empty record row -> argmax of zeros = 0
    assert (ts["reward"][:, 0] == exp).all() and (ts["reward"] == ts["reward"][:, :1]).all()
    # the row index clamps to num_actions-1 (targets go up to 29, only 10 rows)
    row = np.minimum(tgt, spec.num_actions - 1)
    assert (st["env_state"]["record"][np.arange(B), row, 0] == act[:, 0]).all()
    st, ts = ocs.step(spec, st, act)
    assert (ts["reward"][:, 0] == 1.0).all()  # now the modal past first-action equals actions[0]
    miss = act.copy()
    miss[:, 2] += 1
    st, ts = ocs.step(spec, st, miss)
    assert (ts["reward"] == 0).all()


def test_coordsum_autoreset_and_metrics():
    spec, st, ts = _mk(B=3)
    rng = np.random.default_rng(0)
    inner_key = st["env_state"]["key"].copy()
    total = np.zeros(3, np.float32)
    for t in range(100):
        st, ts = ocs.step(spec, st, rng.integers(0, 10, (3, 3)).astype(np.int32))
        total += ts["reward"][:, 0]
    assert (ts["step_type"] == ocs.STEP_LAST).all() and (ts["discount"] == 0).all()
    assert (ts["extras"]["episode_metrics"]["is_terminal_step"]).all()
    assert (ts["extras"]["episode_metrics"]["episode_length"] == 100).all()
    assert np.allclose(ts["extras"]["episode_metrics"]["episode_return"], total)
    # auto-reset: observation is the reset one, real_next_obs the terminal one; key = split(state.key)[0] then reset
    assert (ts["observation"]["step_count"] == 0).all() and (ts["extras"]["real_next_obs"]["step_count"] == 100).all()
    exp_key = np.stack([prng.split(prng.split(k)[0])[0] for k in inner_key])
    assert (st["env_state"]["key"] == exp_key).all()
    assert (st["env_state"]["record"] == -1).all() and (st["running_count_episode_length"] == 0).all()


# ----------------------------------------------------------------------------- networks
def _nets(A=3, d=4, a=10, ffn_zero=True):
    cfg = onets.NetCfg(A, d, a)
    return cfg, onets.init_guider_params(cfg, 0, ffn_zero=ffn_zero), onets.init_actor_params(cfg, 1)


@pytest.mark.parametrize("ffn_zero", [True, False])
def test_recurrent_equals_chunkwise(ffn_zero):
    """SableNetwork.get_actions stepped over t (decay, reset on done, token-by-token decoder) reproduces the
    chunkwise __call__ on the same sequence: pins D, xi, the done semantics and the shifted actions (Appendix G)."""
    cfg, gp, _ = _nets(ffn_zero=ffn_zero)
    p = onets.to_torch(gp, torch.float64)
    rng = np.random.default_rng(0)
    N, T, A, a, d = 3, 9, cfg.n_agents, cfg.action_dim, cfg.obs_dim
    obs = torch.tensor(rng.standard_normal((N, T, A, d)))
    mask = torch.tensor(rng.random((N, T, A, a)) < 0.8)
    mask[..., 0] = True
    steps = torch.tensor(rng.integers(0, 50, (N, T, 1)).repeat(A, axis=2))
    acts = rng.integers(0, a, (N, T, A))
    done = rng.random((N, T)) < 0.25
    done[0, 0] = True
    h0 = tuple(torch.tensor(rng.standard_normal((N, 1, 1, 64, 64)) * 0.1) for _ in range(3))
    hs = h0
    vals, logits = [], []
    for t in range(T):
        dz = torch.tensor(done[:, t])[:, None, None, None, None]
        hs = tuple(torch.where(dz, torch.zeros_like(h), h) for h in hs)  # reset happened after the previous env step
        _, _, v, hs, lg = onets.sable_get_actions(p, cfg, obs[:, t], mask[:, t], steps[:, t], hs, prng.prng_key(0),
                                                  forced_actions=acts[:, t], return_logits=True)
        vals.append(v)
        logits.append(lg)
    v_rec = torch.stack(vals, 1).reshape(N, T * A)
    l_rec = torch.stack(logits, 1).reshape(N, T * A, a)
    dones_tok = torch.tensor(np.repeat(done, A, axis=1))
    v_chk, _, _, l_chk = onets.sable_apply(p, cfg, obs.reshape(N, T * A, d), mask.reshape(N, T * A, a), steps.reshape(N, T * A),
                                           torch.tensor(acts.reshape(N, T * A)), h0, dones_tok, T)
    assert torch.allclose(v_rec, v_chk, rtol=1e-9, atol=1e-11)
    legal = mask.reshape(N, T * A, a)
    assert torch.allclose(l_rec[legal], l_chk[legal], rtol=1e-9, atol=1e-11)


def test_timestep_chunking_is_exact():
    cfg, gp, _ = _nets(ffn_zero=False)
    p = onets.to_torch(gp, torch.float64)
    rng = np.random.default_rng(1)
    N, T, A, a, d = 2, 8, cfg.n_agents, cfg.action_dim, cfg.obs_dim
    args = (torch.tensor(rng.standard_normal((N, T * A, d))), torch.ones(N, T * A, a, dtype=torch.bool),
            torch.tensor(rng.integers(0, 50, (N, T, 1)).repeat(A, axis=2).reshape(N, T * A)),
            torch.tensor(rng.integers(0, a, (N, T * A))),
            tuple(torch.tensor(rng.standard_normal((N, 1, 1, 64, 64)) * 0.1) for _ in range(3)),
            torch.tensor(np.repeat(rng.random((N, T)) < 0.2, A, axis=1)), T)
    v1, lp1, _, _ = onets.sable_apply(p, cfg, *args)
    cfg2 = onets.NetCfg(A, d, a, timestep_chunk_size=2)
    v2, lp2, _, _ = onets.sable_apply(p, cfg2, *args)
    assert torch.allclose(v1, v2, rtol=1e-9, atol=1e-12) and torch.allclose(lp1, lp2, rtol=1e-9, atol=1e-12)


def test_gru_matches_manual_cell():
    cfg, _, ap = _nets()
    p = onets.to_torch(ap, torch.float64)
    rng = np.random.default_rng(2)
    E, A, T = 2, cfg.n_agents, 4
    h = torch.tensor(rng.standard_normal((E, A, 128)))
    obs = torch.tensor(rng.standard_normal((T, E, A, cfg.obs_dim)))
    done = torch.tensor(rng.random((T, E, A)) < 0.3)
    hT, logits = onets.actor_apply(p, cfg, h, obs, done, torch.ones(T, E, A, cfg.action_dim, dtype=torch.bool))
    g = "ScannedRNN_0/GRUCell_0"
    hh = h.clone()
    for t in range(T):
        hh = torch.where(done[t][..., None], torch.zeros_like(hh), hh)
        x = torch.relu(obs[t] @ p["pre_torso/Dense_0/kernel"] + p["pre_torso/Dense_0/bias"])
        r = torch.sigmoid(x @ p[f"{g}/ir/kernel"] + p[f"{g}/ir/bias"] + hh @ p[f"{g}/hr/kernel"])
        z = torch.sigmoid(x @ p[f"{g}/iz/kernel"] + p[f"{g}/iz/bias"] + hh @ p[f"{g}/hz/kernel"])
        n = torch.tanh(x @ p[f"{g}/in/kernel"] + p[f"{g}/in/bias"] + r * (hh @ p[f"{g}/hn/kernel"] + p[f"{g}/hn/bias"]))
        hh = (1 - z) * n + z * hh
    assert torch.allclose(hT, hh)
    assert logits.shape == (T, E, A, cfg.action_dim)


# ----------------------------------------------------------------------------- GAE / optimiser / losses
def test_gae_direct_sum():
    rng = np.random.default_rng(0)
    T, E, A = 20, 3, 2
    r, v = rng.standard_normal((T, E, A)).astype(np.float32), rng.standard_normal((T, E, A)).astype(np.float32)
    d = rng.random((T, E, A)) < 0.2
    lv, ld = rng.standard_normal((E, A)).astype(np.float32), rng.random((E, A)) < 0.2
    adv, tgt = olr.gae(d, v, r, lv, ld, 0.99, 0.95)
    nd = np.concatenate([d[1:], ld[None]]).astype(np.float64)
    nv = np.concatenate([v[1:], lv[None]]).astype(np.float64)
    delta = r + 0.99 * nv * (1 - nd) - v
    ref = np.zeros_like(delta)
    for t in range(T):
        w = np.ones((E, A))
        for k in range(t, T):
            ref[t] += w * delta[k]
            w = w * 0.99 * 0.95 * (1 - nd[k])
    assert np.allclose(adv, ref, atol=1e-4) and np.allclose(tgt, ref + v, atol=1e-4)


def test_clip_adam_closed_form():
    p = {"w": np.array([1.0, -2.0, 3.0], np.float32)}
    g = {"w": np.array([0.3, -0.4, 1.2], np.float32)}  # norm 1.3 > 0.5 -> clipped
    opt = olr.init_opt(p)
    gn = olr.clip_adam_step(p, g, opt, lr=1e-2, max_norm=0.5)
    assert abs(gn - 1.3) < 1e-6
    gc = g["w"] / 1.3 * 0.5
    # first Adam step: mu_hat = g, nu_hat = g^2  ->  update = -lr * g / (|g| + eps)
    exp = np.array([1.0, -2.0, 3.0]) - 1e-2 * gc / (np.abs(gc) + 1e-5)
    assert np.allclose(p["w"], exp, atol=1e-6) and opt["count"] == 1
    p2 = {"w": np.zeros(2, np.float32)}
    olr.clip_adam_step(p2, {"w": np.array([0.1, 0.2], np.float32)}, olr.init_opt(p2), 1e-2, 0.5)  # below the clip
    assert np.allclose(p2["w"], -1e-2 * np.array([0.1, 0.2]) / (np.array([0.1, 0.2]) + 1e-5), atol=1e-6)


def test_loss_gradients_finite_difference():
    """fp64 central differences of both MAGPO losses w.r.t. a few parameters against autograd."""
    cfg, gp, ap = _nets(ffn_zero=False)
    sysc = olr.SysCfg(num_envs=2, update_batch_size=1, rollout_length=4, num_minibatches=1)
    rng = np.random.default_rng(0)
    N, T, A, a, d = 2, 4, cfg.n_agents, cfg.action_dim, cfg.obs_dim
    C_ = T * A
    mb = dict(obs=rng.standard_normal((N, C_, d)), action_mask=np.ones((N, C_, a), bool),
              step_count=rng.integers(0, 9, (N, T, 1)).repeat(A, 2).reshape(N, C_), action=rng.integers(0, a, (N, C_)),
              done=np.repeat(rng.random((N, T)) < 0.3, A, 1), value=rng.standard_normal((N, C_)),
              log_prob=-np.abs(rng.standard_normal((N, C_))) - 1.5, adv=rng.standard_normal((N, C_)),
              targets=rng.standard_normal((N, C_)), policy_h0=rng.standard_normal((N, A, 128)) * 0.1,
              prev_hstates=tuple(rng.standard_normal((N, 1, 1, 64, 64)) * 0.1 for _ in range(3)))
    gp = {k: v.astype(np.float64) + 0.05 * rng.standard_normal(v.shape) for k, v in gp.items()}
    ap = {k: v.astype(np.float64) + 0.05 * rng.standard_normal(v.shape) for k, v in ap.items()}
    gg, ga, info, _ = olr.minibatch_losses_and_grads(gp, ap, mb, cfg, sysc, dtype=torch.float64)

    def tot(gp_, ap_, which):
        return olr.minibatch_losses_and_grads(gp_, ap_, mb, cfg, sysc, dtype=torch.float64)[2][which]

    eps = 1e-6
    for name, idx in (("decoder/head/layers_3/kernel", (3, 2)), ("encoder/encoder_block_0/retn/retention_heads_0/w_k", (5, 7)),
                      ("decoder/decoder_block_0/retn2/w_g", (1, 1)), ("encoder/obs_encoder/layers_1/kernel", (2, 9))):
        plus = {k: v.copy() for k, v in gp.items()}
        minus = {k: v.copy() for k, v in gp.items()}
        plus[name][idx] += eps
        minus[name][idx] -= eps
        fd = (tot(plus, ap, "total_guider") - tot(minus, ap, "total_guider")) / (2 * eps)
        assert abs(fd - gg[name][idx]) < 1e-5 * max(1.0, abs(fd)), name
    for name, idx in (("ScannedRNN_0/GRUCell_0/hn/kernel", (4, 4)), ("pre_torso/Dense_0/kernel", (1, 3)),
                      ("action_head/Dense_0/bias", (2,))):
        plus = {k: v.copy() for k, v in ap.items()}
        minus = {k: v.copy() for k, v in ap.items()}
        plus[name][idx] += eps
        minus[name][idx] -= eps
        fd = (tot(gp, plus, "total_actor") - tot(gp, minus, "total_actor")) / (2 * eps)
        assert abs(fd - ga[name][idx]) < 1e-5 * max(1.0, abs(fd)), name


def test_update_step_runs_and_keeps_hstate_permutation_quirk():
    spec = ocs.CoordSumSpec(**ocs.SCENARIOS["3x10-30-v0"])
    cfg = onets.NetCfg(3, 4, 10)
    sysc = olr.SysCfg(num_envs=4, update_batch_size=2, rollout_length=6, ppo_epochs=2, num_minibatches=2)
    state = olr.learner_setup(spec, cfg, sysc, seed=42)
    rec = {}
    mets, infos = olr.update_step(state, spec, cfg, sysc, record=rec)
    assert len(infos) == 4 and all(np.isfinite(list(i.values())).all() for i in infos)
    assert state["guider_opt"]["count"] == 4 and state["actor_opt"]["count"] == 4
    assert (state["slots"][0]["key"] == state["slots"][1]["key"]).all()  # identical PRNG stream in every slot


# ----------------------------------------------------------------------------- the C-ABI library
def test_library_exports_every_declared_symbol():
    from magpo_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build the library first (make / __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "magpo_b200.h")).read()
    names = sorted(set(re.findall(r"\b(magpo_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    lib.magpo_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.magpo_version()


def test_param_table_matches_oracle_param_names():
    from magpo_b200.learner import NetworkConfig, param_table

    net = NetworkConfig(3, 4, 10, 100)
    cfg = onets.NetCfg(3, 4, 10)
    for which, ref in ((0, onets.init_guider_params(cfg)), (1, onets.init_actor_params(cfg))):
        table, total = param_table(net, which)
        assert {t[0] for t in table} == set(ref)
        covered = np.zeros(total, bool)
        for name, off, d0, d1, ld in table:
            shape = ref[name].shape
            assert (d0,) == shape if d1 == 0 else (d0, d1) == shape, name
            for r in range(d0 if d1 else 1):
                sl = slice(off + r * ld, off + r * ld + (d1 if d1 else d0))
                assert not covered[sl].any(), name  # no overlap between tensors
                covered[sl] = True
        assert covered.sum() == sum(v.size for v in ref.values())


def test_fold_in_and_normal_restatements():
    """fold_in(key, i) == split(key, n)[i] under the partitionable threefry scheme (both are threefry(key, (0, i))), the library's host
    fold_in agrees with the oracle's, and the normal draws have the right moments / tails (erf_inv against scipy in float64)."""
    import ctypes as C

    from scipy.special import erfinv

    from magpo_b200 import _lib as L

    key = prng.prng_key(7)
    ks = prng.split(key, 6)
    for i in range(6):
        assert (prng.fold_in(key, i) == ks[i]).all()
    out = (C.c_uint32 * 2)()
    kk = (C.c_uint32 * 2)(int(key[0]), int(key[1]))
    assert L.lib().magpo_prng_fold_in_host(kk, C.c_uint32(0xDEADBEEF), out) == 0
    assert (np.array([out[0], out[1]], np.uint32) == prng.fold_in(key, 0xDEADBEEF)).all()
    x = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    assert np.abs(prng.erf_inv_f32(x) - erfinv(x.astype(np.float64))).max() < 3e-6 * 3.5
    z = prng.normal(key, (200_000,))
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01 and np.isfinite(z).all()
    t = prng.truncated_normal(key, -2.0, 2.0, (100_000,))
    assert float(t.min()) > -2.0 and float(t.max()) < 2.0 and abs(float(t.std()) - 0.8796) < 0.01
