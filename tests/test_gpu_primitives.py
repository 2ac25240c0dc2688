"""-m gpu: the integer / building-block kernels through the C ABI against the CPU oracle.
Bit-exact: threefry bits, split, randint, permutation, every CoordSum state/timestep field, GAE.
ulp-level: gumbel noise (logf). rtol: GEMM / retention (different summation order than the oracle)."""
import ctypes as C

import numpy as np
import pytest
import torch

from magpo_b200 import _lib as L
from oracle import coordsum as ocs
from oracle import learner as olr
from oracle import nets as onets
from oracle import prng as oprng

from gpu_util import as_u32, dt, rel_err, sync, u32

pytestmark = pytest.mark.gpu


def test_prng_matches_oracle(dev):
    s = L.stream_ptr()
    for seed in (0, 42, 2**33 + 7):
        key = oprng.prng_key(seed)
        k = u32(key, dev)
        out = torch.zeros(7, 2, dtype=torch.int32, device=dev)
        L.call("magpo_prng_split", s, L.ptr(k), 7, L.ptr(out))
        assert (as_u32(out) == oprng.split(key, 7)).all()
        n = 1000
        bits = torch.zeros(n, dtype=torch.int32, device=dev)
        L.call("magpo_prng_random_bits", s, L.ptr(k), C.c_int64(n), L.ptr(bits))
        assert (as_u32(bits) == oprng.random_bits(key, (n,))).all()
        ri = torch.zeros(n, dtype=torch.int32, device=dev)
        L.call("magpo_prng_randint", s, L.ptr(k), C.c_int64(n), 0, 30, L.ptr(ri))
        assert (ri.cpu().numpy() == oprng.randint(key, (n,), 0, 30)).all()
        g = torch.zeros(n, dtype=torch.float32, device=dev)
        L.call("magpo_prng_gumbel", s, L.ptr(k), C.c_int64(n), L.ptr(g))
        ref = oprng.gumbel(key, (n,))
        assert np.allclose(g.cpu().numpy(), ref, rtol=4e-6, atol=1e-6)  # <= a few ulp of logf
        for m in (1, 3, 16, 1000, 4096):
            perm = torch.zeros(m, dtype=torch.int32, device=dev)
            scr = torch.zeros(2 * m + 8, dtype=torch.int32, device=dev)
            L.call("magpo_prng_permutation", s, L.ptr(k), m, L.ptr(perm), L.ptr(scr))
            assert (perm.cpu().numpy() == oprng.permutation(key, m)).all(), m


@pytest.mark.parametrize("scenario", ["3x10-30-v0", "5x20-80-v0", "8x15-100-v0"])
def test_coordsum_bit_exact(dev, scenario):
    from magpo_b200.learner import CoordSumVec, alloc_timestep

    kw = ocs.SCENARIOS[scenario]
    spec = ocs.CoordSumSpec(**kw)
    env = CoordSumVec(**kw)
    B, A, a, d = 48, spec.num_agents, spec.num_actions, spec.obs_dim
    keys = oprng.split(oprng.prng_key(7), B)
    ostate, ots = ocs.reset(spec, keys)
    st = env.alloc_state(B, dev)
    ts = alloc_timestep(B, A, d, a, dev)
    s = L.stream_ptr()
    kd = u32(keys, dev)
    L.call("magpo_coordsum_reset", s, C.byref(env.cfg), B, L.ptr(kd), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
    rng = np.random.default_rng(0)

    def check(tag):
        sync()
        b = ostate["env_state"]
        assert (st["step_count"].cpu().numpy() == b["step_count"]).all(), tag
        assert (st["target"].cpu().numpy() == b["target"]).all(), tag
        assert (st["record"].cpu().numpy() == b["record"]).all(), tag
        assert (as_u32(st["key"]) == b["key"]).all(), tag
        assert (as_u32(st["metrics_key"]) == ostate["key"]).all(), tag
        for k1, k2 in (("running_return", "running_count_episode_return"), ("running_length", "running_count_episode_length"),
                       ("episode_return", "episode_return"), ("episode_length", "episode_length")):
            assert (st[k1].cpu().numpy() == ostate[k2]).all(), (tag, k1)
        ob = ots["observation"]
        assert (ts["agents_view"].cpu().numpy() == ob["agents_view"].astype(np.float32)).all(), tag
        assert (ts["action_mask"].cpu().numpy().astype(bool) == ob["action_mask"]).all(), tag
        assert (ts["step_count"].cpu().numpy() == ob["step_count"]).all(), tag
        assert (ts["step_type"].cpu().numpy() == ots["step_type"]).all(), tag
        assert (ts["reward"].cpu().numpy() == ots["reward"]).all(), tag
        assert (ts["discount"].cpu().numpy() == ots["discount"]).all(), tag
        ex = ots["extras"]
        assert (ts["next_agents_view"].cpu().numpy() == ex["real_next_obs"]["agents_view"].astype(np.float32)).all(), tag
        assert (ts["next_step_count"].cpu().numpy() == ex["real_next_obs"]["step_count"]).all(), tag
        em = ex["episode_metrics"]
        assert (ts["episode_return"].cpu().numpy() == em["episode_return"]).all(), tag
        assert (ts["episode_length"].cpu().numpy() == em["episode_length"]).all(), tag
        assert (ts["is_terminal_step"].cpu().numpy().astype(bool) == em["is_terminal_step"]).all(), tag

    check("reset")
    for step in range(230):  # crosses two auto-resets
        # bias the actions so that sum matches / modal hits actually occur
        tgt = ostate["env_state"]["target"][np.arange(B), ostate["env_state"]["step_count"]]
        act = rng.integers(0, a, size=(B, A)).astype(np.int32)
        fix = rng.random(B) < 0.5
        rest = act[:, 1:].sum(1)
        want = np.clip(tgt - rest, 0, a - 1)
        act[fix, 0] = want[fix]
        ostate, ots = ocs.step(spec, ostate, act)
        ad = dt(act, dev)
        L.call("magpo_coordsum_step", s, C.byref(env.cfg), B, L.ptr(ad), env.state_struct(st), L.struct_of(L.TimeStep, **ts))
        if step % 7 == 0 or step in (98, 99, 100, 101, 199, 200):
            check(f"step {step}")
    check("end")


@pytest.mark.parametrize("T,B,A", [(128, 32, 3), (16, 1, 1), (33, 1000, 4), (1, 5, 2), (300, 50, 2), (129, 7, 3), (512, 11, 4)])
def test_gae_bit_exact(dev, T, B, A):
    rng = np.random.default_rng(1)
    reward = rng.standard_normal((T, B, A)).astype(np.float32)
    value = rng.standard_normal((T, B, A)).astype(np.float32)
    done_env = rng.random((T, B)) < 0.1
    last_value = rng.standard_normal((B, A)).astype(np.float32)
    last_done = rng.random(B) < 0.2
    adv_ref, tgt_ref = olr.gae(np.repeat(done_env[..., None], A, -1), value, reward, last_value,
                               np.repeat(last_done[:, None], A, -1), 0.99, 0.95)
    adv = torch.zeros(T, B, A, device=dev)
    tgt = torch.zeros(T, B, A, device=dev)
    keep = [dt(reward, dev), dt(value, dev), dt(done_env.astype(np.uint8), dev), dt(last_value, dev),
            dt(last_done.astype(np.uint8), dev)]  # device tensors must outlive the call
    L.call("magpo_gae", L.stream_ptr(), T, B, A, *[L.ptr(k) for k in keep], C.c_double(0.99), C.c_double(0.95), L.ptr(adv),
           L.ptr(tgt))
    assert (adv.cpu().numpy() == adv_ref).all()
    assert (tgt.cpu().numpy() == tgt_ref).all()
    # O(T^2) direct sum (size-independent property)
    g, lam = 0.99, 0.95
    nd = np.concatenate([done_env[1:], last_done[None]], 0).astype(np.float64)[..., None]
    nv = np.concatenate([value[1:], last_value[None]], 0).astype(np.float64)
    delta = reward + g * nv * (1 - nd) - value
    direct = np.zeros_like(delta)
    for t in range(T):
        w = np.ones((B, 1))
        for k in range(t, T):
            direct[t] += w * delta[k]
            w = w * g * lam * (1 - nd[k])
    assert np.allclose(adv.cpu().numpy(), direct, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (257, 256, 64), (4096, 384, 128), (300, 10, 128), (513, 64, 4), (129, 1, 64),
                                   (640, 128, 75)])
def test_gemm_blocks(dev, M, N, K):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((M, K)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    Y = torch.zeros(M, N, device=dev)
    s = L.stream_ptr()
    Xd, Wd, bd = dt(X, dev), dt(W, dev), dt(b, dev)
    L.call("magpo_test_gemm", s, 0, C.c_int64(M), N, K, L.ptr(Xd), L.ptr(Wd), L.ptr(bd), L.ptr(Y), 0)
    ref = X.astype(np.float64) @ W.astype(np.float64) + b
    assert rel_err(Y.cpu().numpy(), ref) < 2e-6
    L.call("magpo_test_gemm", s, 0, C.c_int64(M), N, K, L.ptr(Xd), L.ptr(Wd), L.ptr(bd), L.ptr(Y), 3)
    assert rel_err(Y.cpu().numpy(), np.maximum(2 * ref, 0)) < 2e-6  # accumulate + relu
    dY = rng.standard_normal((M, N)).astype(np.float32)
    dW = torch.zeros(K, N, device=dev)
    dYd = dt(dY, dev)
    L.call("magpo_test_gemm", s, 1, C.c_int64(M), N, K, L.ptr(Xd), L.ptr(dYd), None, L.ptr(dW), 0)
    assert rel_err(dW.cpu().numpy(), X.astype(np.float64).T @ dY.astype(np.float64)) < 1e-5
    db = torch.zeros(N, device=dev)
    L.call("magpo_test_gemm", s, 2, C.c_int64(M), N, K, L.ptr(dYd), None, None, L.ptr(db), 0)
    assert rel_err(db.cpu().numpy(), dY.astype(np.float64).sum(0)) < 1e-5
    WT = torch.zeros(N, K, device=dev)
    L.call("magpo_test_gemm", s, 3, C.c_int64(K), N, K, L.ptr(Wd), None, None, L.ptr(WT), 0)
    assert (WT.cpu().numpy() == W.T).all()


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("T,N,A", [(12, 5, 3), (40, 3, 4), (1, 7, 1), (9, 2, 8), (9, 100, 3), (16, 300, 2)])
def test_retention_scan_equals_chunkwise_reference(dev, causal, T, N, A):
    """The recurrent-form CUDA scan (fwd + bwd) against the oracle's chunkwise D-matrix/xi form with autograd."""
    rng = np.random.default_rng(3)
    kappa = float(onets.NetCfg(A, 4, 4).kappas()[0])
    C_ = T * A
    q, k, v = (torch.tensor(rng.standard_normal((N, C_, 64)) * 0.5, dtype=torch.float64, requires_grad=True) for _ in range(3))
    H0 = torch.tensor(rng.standard_normal((N, 64, 64)) * 0.3, dtype=torch.float64)
    ts_done = rng.random((N, T)) < 0.15
    ts_done[0, 0] = True
    dones = torch.tensor(np.repeat(ts_done, A, axis=1))
    eye = {"h/w_q": torch.eye(64, dtype=torch.float64), "h/w_k": torch.eye(64, dtype=torch.float64), "h/w_v": torch.eye(64, dtype=torch.float64)}
    ret_ref, nh_ref = onets.simple_retention_chunk(eye, "h", k, q, v, H0, dones, kappa, A, causal)
    dret = torch.tensor(rng.standard_normal((N, C_, 64)), dtype=torch.float64)
    gq, gk, gv = torch.autograd.grad((ret_ref * dret).sum(), [q, k, v])

    def tm(x):  # [N, C, 64] -> [T, N, A, 64]
        return np.ascontiguousarray(np.swapaxes(x.detach().numpy().reshape(N, T, A, 64), 0, 1)).astype(np.float32)

    packed = np.zeros((T, N, A, 256), np.float32)
    packed[..., 0:64], packed[..., 64:128], packed[..., 128:192] = tm(q), tm(k), tm(v)
    pk = dt(packed, dev)
    ret = torch.zeros(T, N, A, 64, device=dev)
    Hs = torch.zeros(T, N, 64, 64, device=dev)
    Hout = torch.zeros(N, 64, 64, device=dev)
    done_d = dt(ts_done.T.astype(np.uint8), dev)
    H0d = dt(H0.numpy().astype(np.float32), dev)
    s = L.stream_ptr()
    base = pk.data_ptr()
    L.call("magpo_test_retention", s, 0, T, N, A, C.c_float(kappa), int(causal), C.c_void_p(base), C.c_void_p(base + 256),
           C.c_void_p(base + 512), 256, L.ptr(H0d), L.ptr(done_d), L.ptr(ret), L.ptr(Hs), L.ptr(Hout), None, None, None, None, 0)
    assert rel_err(ret.cpu().numpy(), tm(ret_ref)) < 2e-5
    if not ts_done.any(axis=1).all():
        pass
    # the chunk's next_hstate (retention.py:88-92) equals the scan's final state
    assert rel_err(Hout.cpu().numpy(), nh_ref.detach().numpy()) < 2e-5
    dpk = torch.zeros(T, N, A, 256, device=dev)
    dretd = dt(tm(dret), dev)
    db = dpk.data_ptr()
    L.call("magpo_test_retention", s, 1, T, N, A, C.c_float(kappa), int(causal), C.c_void_p(base), C.c_void_p(base + 256),
           C.c_void_p(base + 512), 256, L.ptr(H0d), L.ptr(done_d), None, L.ptr(Hs), None, L.ptr(dretd),
           C.c_void_p(db), C.c_void_p(db + 256), C.c_void_p(db + 512), 256)
    out = dpk.cpu().numpy()
    assert rel_err(out[..., 0:64], tm(gq)) < 5e-5
    assert rel_err(out[..., 64:128], tm(gk)) < 5e-5
    assert rel_err(out[..., 128:192], tm(gv)) < 5e-5


@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (4096, 256, 64), (777, 128, 64), (512, 192, 64), (2048, 384, 128), (3000, 128, 128),
                                   (1024, 64, 256), (1500, 64, 192), (2000, 128, 384), (300, 64, 64), (40000, 256, 64)])
def test_gemm_tensor_core_3xtf32(dev, M, N, K):
    """tcgen05 + TMA GEMM (3xTF32 split) against fp64: fp32-faithful accuracy, bias / relu / accumulate epilogues,
    sub-matrix views (leading dimensions larger than the logical width)."""
    rng = np.random.default_rng(5)
    ldx, ldy = K + 64, N + 32
    Xf = rng.standard_normal((M, ldx)).astype(np.float32)
    W = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    Xd, WTd, bd = dt(Xf, dev), dt(np.ascontiguousarray(W.T), dev), dt(b, dev)
    scratch = torch.zeros(2 * N * K, device=dev)
    Y = torch.full((M, ldy), 7.0, device=dev)
    s = L.stream_ptr()
    ref = Xf[:, :K].astype(np.float64) @ W.astype(np.float64)
    L.call("magpo_test_gemm_tc", s, C.c_int64(M), N, K, L.ptr(Xd), ldx, L.ptr(WTd), L.ptr(scratch), L.ptr(bd), L.ptr(Y), ldy, 0)
    out = Y.cpu().numpy()
    tol = 2e-6 * max(1.0, (K / 64) ** 0.5)  # error grows like sqrt(K) as in any fp32 dot product
    assert rel_err(out[:, :N], ref + b) < tol
    assert (out[:, N:] == 7.0).all()  # nothing written outside the logical width
    L.call("magpo_test_gemm_tc", s, C.c_int64(M), N, K, L.ptr(Xd), ldx, L.ptr(WTd), L.ptr(scratch), L.ptr(bd), L.ptr(Y), ldy, 2)
    assert rel_err(Y.cpu().numpy()[:, :N], np.maximum(ref + b, 0)) < tol
    L.call("magpo_test_gemm_tc", s, C.c_int64(M), N, K, L.ptr(Xd), ldx, L.ptr(WTd), L.ptr(scratch), None, L.ptr(Y), ldy, 1)
    assert rel_err(Y.cpu().numpy()[:, :N], np.maximum(ref + b, 0) + ref) < tol  # TMA reduce-add


@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (4096, 256, 64), (777, 128, 64), (50000, 64, 64), (2048, 384, 128), (3000, 128, 128),
                                   (300, 192, 64), (33000, 384, 128)])
def test_gemm_tensor_core_tn_weight_grad(dev, M, N, K):
    """dW += X^T dY on tcgen05 with MN-major operands (both split 3xTF32), TMA reduce-add across CTAs; accumulates
    into existing contents and respects leading dimensions."""
    rng = np.random.default_rng(6)
    ldx, ldy, ldw = K + 64, N + 32, N + 64
    Xf = rng.standard_normal((M, ldx)).astype(np.float32)
    dYf = rng.standard_normal((M, ldy)).astype(np.float32)
    init = rng.standard_normal((K, ldw)).astype(np.float32)
    Xd, dYd, dW = dt(Xf, dev), dt(dYf, dev), dt(init, dev)
    L.call("magpo_test_gemm_tc_tn", L.stream_ptr(), C.c_int64(M), N, K, L.ptr(Xd), ldx, L.ptr(dYd), ldy, L.ptr(dW), ldw)
    ref = Xf[:, :K].astype(np.float64).T @ dYf[:, :N].astype(np.float64) + init[:, :N]
    out = dW.cpu().numpy()
    assert rel_err(out[:, :N], ref) < 3e-6 * max(1.0, (M / 1000) ** 0.5)
    assert (out[:, N:] == init[:, N:]).all()


@pytest.mark.parametrize("R,K,N", [(1000, 4, 128), (777, 14, 64), (5000, 16, 128), (33, 1, 64), (3000, 10, 128)])
def test_thin_k_layers(dev, R, K, N):
    """thin.cu: Y = relu(X W + b) and (dW, db) = (X^T dY, colsum dY) for a thin input side, against fp64 NumPy."""
    rng = np.random.default_rng(11)
    X = rng.standard_normal((R, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) * 0.5).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((R, N)).astype(np.float32)
    Xd, Wd, bd, dYd = dt(X, dev), dt(W, dev), dt(b, dev), dt(dY, dev)
    Y = torch.zeros(R, N, device=dev)
    s = L.stream_ptr()
    L.call("magpo_test_thin", s, 0, C.c_int64(R), K, N, L.ptr(Xd), L.ptr(Wd), L.ptr(bd), None, L.ptr(Y), None, None, 2)
    ref = np.maximum(X.astype(np.float64) @ W + b, 0)
    assert rel_err(Y.cpu().numpy(), ref) < 2e-6
    dW = torch.full((K, N), 0.5, device=dev)
    db = torch.full((N,), -0.25, device=dev)
    L.call("magpo_test_thin", s, 1, C.c_int64(R), K, N, L.ptr(Xd), None, None, L.ptr(dYd), L.ptr(dW), L.ptr(db), None, 0)
    assert rel_err(dW.cpu().numpy(), X.astype(np.float64).T @ dY + 0.5) < 3e-6 * max(1.0, (R / 1000) ** 0.5)
    assert rel_err(db.cpu().numpy(), dY.astype(np.float64).sum(0) - 0.25) < 3e-6 * max(1.0, (R / 1000) ** 0.5)
    # with the relu mask of the layer's own output (Y from above) applied to dY on the fly
    dW.fill_(0.0)
    db.fill_(0.0)
    L.call("magpo_test_thin", s, 1, C.c_int64(R), K, N, L.ptr(Xd), L.ptr(Y), None, L.ptr(dYd), L.ptr(dW), L.ptr(db), None, 2)
    dYm = dY.astype(np.float64) * (ref > 0)
    assert rel_err(dW.cpu().numpy(), X.astype(np.float64).T @ dYm) < 3e-6 * max(1.0, (R / 1000) ** 0.5)
    assert rel_err(db.cpu().numpy(), dYm.sum(0)) < 3e-6 * max(1.0, (R / 1000) ** 0.5)


@pytest.mark.parametrize("R,N", [(1000, 10), (777, 5), (5000, 16), (33, 1), (2500, 6)])
def test_thin_n_layers(dev, R, N):
    """thin.cu: the action head Y = X W + b (K = 128) and its fused backward (dX masked by X > 0, dW, db)."""
    K = 128
    rng = np.random.default_rng(12)
    X = rng.standard_normal((R, K)).astype(np.float32)
    X[X < -0.5] = 0.0  # relu-like activations with exact zeros
    W = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    dY = rng.standard_normal((R, N)).astype(np.float32)
    Xd, Wd, bd, dYd = dt(X, dev), dt(W, dev), dt(b, dev), dt(dY, dev)
    Y = torch.zeros(R, N, device=dev)
    s = L.stream_ptr()
    L.call("magpo_test_thin", s, 2, C.c_int64(R), K, N, L.ptr(Xd), L.ptr(Wd), L.ptr(bd), None, L.ptr(Y), None, None, 0)
    assert rel_err(Y.cpu().numpy(), X.astype(np.float64) @ W + b) < 2e-6
    dX = torch.zeros(R, K, device=dev)
    dW = torch.full((K, N), 0.5, device=dev)
    db = torch.full((N,), -0.25, device=dev)
    dbx = torch.full((K,), 0.125, device=dev)
    L.call("magpo_test_thin", s, 3, C.c_int64(R), K, N, L.ptr(Xd), L.ptr(Wd), L.ptr(dbx), L.ptr(dYd), L.ptr(dX), L.ptr(dW), L.ptr(db), 2 | 4)
    assert rel_err(dX.cpu().numpy(), (dY.astype(np.float64) @ W.T) * (X > 0)) < 2e-6
    assert rel_err(dbx.cpu().numpy(), ((dY.astype(np.float64) @ W.T) * (X > 0)).sum(0) + 0.125) < 3e-6 * max(1.0, (R / 1000) ** 0.5)
    assert rel_err(dW.cpu().numpy(), X.astype(np.float64).T @ dY + 0.5) < 3e-6 * max(1.0, (R / 1000) ** 0.5)
    assert rel_err(db.cpu().numpy(), dY.astype(np.float64).sum(0) - 0.25) < 3e-6 * max(1.0, (R / 1000) ** 0.5)


def test_prng_normal_and_flax_shaped_init(dev):
    """jax.random.normal / truncated_normal kernels against the NumPy restatement (same threefry bits, same erf_inv expansion; libm
    log1p may differ in the last bits), and the flax-shaped initialisers built on them: orthogonal kernels are orthogonal with the
    right gain, every parameter gets its own key, the result is a pure function of the net key."""
    from magpo_b200 import init as minit
    from oracle import prng as oprng

    key = oprng.split(oprng.prng_key(42), 4)[3]
    got = minit._draw(key, 5000, dev)
    ref = oprng.normal(key, (5000,))
    assert np.abs(got - ref).max() <= 4e-6 * max(1.0, float(np.abs(ref).max()))
    got_t = minit._draw(key, 5000, dev, truncated=True)
    ref_t = oprng.truncated_normal(key, -2.0, 2.0, (5000,))
    assert np.abs(got_t - ref_t).max() <= 4e-6 and float(np.abs(got_t).max()) < 2.0
    gp = minit.flax_init_guider(key, 3, 7, 5, dev)
    gp2 = minit.flax_init_guider(key, 3, 7, 5, dev)
    other = minit.flax_init_guider(oprng.split(key)[0], 3, 7, 5, dev)
    for name, v in gp.items():
        assert v.dtype == np.float32 and (v == gp2[name]).all(), name
    w = gp["encoder/head/layers_0/kernel"]
    assert np.abs(w.T @ w - 2.0 * np.eye(64)).max() < 1e-4  # orthogonal(sqrt 2)
    w = gp["encoder/obs_encoder/layers_1/kernel"]  # rows < cols: orthonormal rows
    assert np.abs(w @ w.T - 2.0 * np.eye(7)).max() < 1e-4
    w = gp["decoder/head/layers_3/kernel"]
    assert np.abs(w.T @ w - 1e-4 * np.eye(5)).max() < 1e-8
    wq, wk = gp["encoder/encoder_block_0/retn/retention_heads_0/w_q"], gp["encoder/encoder_block_0/retn/retention_heads_0/w_k"]
    assert abs(float(wq.std()) - 1 / 64) < 2e-3 and not np.allclose(wq, wk)
    assert not np.allclose(wq, gp["decoder/decoder_block_0/retn1/retention_heads_0/w_q"])
    assert not np.allclose(wq, other["encoder/encoder_block_0/retn/retention_heads_0/w_q"])
    assert (gp["encoder/encoder_block_0/ffn/W_gate"] == 0).all() and (gp["encoder/ln/scale"] == 1).all()
    ap = minit.flax_init_actor(key, 7, 5, dev)
    wh = ap["ScannedRNN_0/GRUCell_0/hr/kernel"]
    assert np.abs(wh.T @ wh - np.eye(128)).max() < 1e-4
    wi = ap["ScannedRNN_0/GRUCell_0/ir/kernel"]
    assert abs(float(wi.std()) - 1 / np.sqrt(128)) < 5e-3 and float(np.abs(wi).max()) <= 2.0 / np.sqrt(128) / 0.8796 + 1e-6
