"""Parameter initialisation and key derivation of `learner_setup` (rec_magpo.py:533-685).

Initialisers follow the flax ones named in the reference (orthogonal(sqrt(2)) torsos / Sable encoders,
orthogonal(0.01) output layers, normal(1/embed_dim) retention weights, zeros SwiGLU, lecun_normal / orthogonal
GRU kernels — networks/torsos.py:40,89-95, sable_network.py:97-107,263-282, retention.py:50-64,237-246, flax
GRUCell). The *distributions* match; the bits cannot (flax draws them from jax.random through a QR that is not
reproducible without JAX), so benchmarks and tests pass parameters explicitly when bit-level agreement matters.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L


def _orthogonal(rng, rows, cols, gain):
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q[:rows, :cols]).astype(np.float32)


def _normal(rng, rows, cols, std):
    return (rng.standard_normal((rows, cols)) * std).astype(np.float32)


def init_guider(n_agents: int, obs_dim: int, action_dim: int, seed: int = 0, D: int = 64) -> dict:
    rng = np.random.default_rng(seed)
    ones, zeros = (lambda n: np.ones(n, np.float32)), (lambda *s: np.zeros(s, np.float32))
    p = {"encoder/obs_encoder/layers_0/scale": ones(obs_dim),
         "encoder/obs_encoder/layers_1/kernel": _orthogonal(rng, obs_dim, D, math.sqrt(2)),
         "encoder/ln/scale": ones(D)}

    def retn(pre):
        p[f"{pre}/w_g"], p[f"{pre}/w_o"] = _normal(rng, D, D, 1 / D), _normal(rng, D, D, 1 / D)
        p[f"{pre}/group_norm/scale"], p[f"{pre}/group_norm/bias"] = ones(D), zeros(D)
        for w in ("w_q", "w_k", "w_v"):
            p[f"{pre}/retention_heads_0/{w}"] = _normal(rng, D, D, 1 / D)

    def ffn(pre):
        for w in ("W_linear", "W_gate", "W_output"):
            p[f"{pre}/{w}"] = zeros(D, D)

    def head(pre, nout):
        p[f"{pre}/layers_0/kernel"], p[f"{pre}/layers_0/bias"] = _orthogonal(rng, D, D, math.sqrt(2)), zeros(D)
        p[f"{pre}/layers_2/scale"] = ones(D)
        p[f"{pre}/layers_3/kernel"], p[f"{pre}/layers_3/bias"] = _orthogonal(rng, D, nout, 0.01), zeros(nout)

    eb = "encoder/encoder_block_0"
    p[f"{eb}/ln1/scale"], p[f"{eb}/ln2/scale"] = ones(D), ones(D)
    retn(f"{eb}/retn")
    ffn(f"{eb}/ffn")
    head("encoder/head", 1)
    p["decoder/action_encoder/layers_0/kernel"] = _orthogonal(rng, action_dim + 1, D, math.sqrt(2))
    p["decoder/ln/scale"] = ones(D)
    db = "decoder/decoder_block_0"
    for ln in ("ln1", "ln2", "ln3"):
        p[f"{db}/{ln}/scale"] = ones(D)
    retn(f"{db}/retn1")
    retn(f"{db}/retn2")
    ffn(f"{db}/ffn")
    head("decoder/head", action_dim)
    return p


def init_actor(obs_dim: int, action_dim: int, seed: int = 1, H: int = 128) -> dict:
    rng = np.random.default_rng(seed)
    zeros = lambda *s: np.zeros(s, np.float32)
    g = "ScannedRNN_0/GRUCell_0"
    p = {"pre_torso/Dense_0/kernel": _orthogonal(rng, obs_dim, H, math.sqrt(2)), "pre_torso/Dense_0/bias": zeros(H)}
    for n in ("ir", "iz", "in"):
        p[f"{g}/{n}/kernel"], p[f"{g}/{n}/bias"] = _normal(rng, H, H, 1 / math.sqrt(H)), zeros(H)
    for n in ("hr", "hz", "hn"):
        p[f"{g}/{n}/kernel"] = _orthogonal(rng, H, H, 1.0)
    p[f"{g}/hn/bias"] = zeros(H)
    p["post_torso/Dense_0/kernel"], p["post_torso/Dense_0/bias"] = _orthogonal(rng, H, H, math.sqrt(2)), zeros(H)
    p["action_head/Dense_0/kernel"], p["action_head/Dense_0/bias"] = _orthogonal(rng, H, action_dim, 0.01), zeros(action_dim)
    return p


def prng_key(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) as raw uint32[2]."""
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key: np.ndarray, num: int = 2, device="cuda:0") -> np.ndarray:
    """jax.random.split through the library's threefry kernel (jax_threefry_partitionable scheme)."""
    k = torch.as_tensor(np.asarray(key, np.uint32).view(np.int32)).to(device)
    out = torch.zeros(num, 2, dtype=torch.int32, device=device)
    L.call("magpo_prng_split", L.stream_ptr(), L.ptr(k), num, L.ptr(out))
    return out.cpu().numpy().view(np.uint32)


def setup_keys(seed: int, n_devices: int, U: int, E: int, device="cuda:0"):
    """Key plumbing of run_experiment / learner_setup (rec_magpo.py:699-701, 642-673):
    returns (env reset keys [n_devices, U*E, 2], the step key shared by every device and slot, net keys)."""
    key, key_e, actor_net_key, net_key = split(prng_key(seed), 4, device)
    allk = split(key, n_devices * U * E + 1, device)
    key, env_keys = allk[0], allk[1:].reshape(n_devices, U * E, 2)
    step_key = split(key, 2, device)[1]
    return env_keys, step_key, (actor_net_key, net_key)
