"""jax.random (threefry2x32) restated in NumPy — oracle, test infrastructure only.

The source lives in jax 0.6.0 (`jax/_src/prng.py`, `jax/_src/random.py`), which is a
pinned dependency of the reference (`uv.lock:993-995`) and is NOT under
/root/reference; the algorithms are restated from SURVEY.md Appendix A1-A5.
Reference call sites: `rec_magpo.py:135,202,373,439,443,450,642,660,699`,
`coordsum/env.py:56-57`, `wrappers/auto_reset_wrapper.py:74`,
`wrappers/episode_metrics.py:62`, `networks/utils/sable/decode.py:140-142`.

Keys are raw uint32[2] arrays. `PARTITIONABLE` mirrors
`jax_threefry_partitionable` (default True since jax 0.5.0).
"""
from __future__ import annotations

import math

import numpy as np

PARTITIONABLE = True

_R0 = (13, 15, 26, 6)
_R1 = (17, 29, 16, 24)
_U32 = np.uint32


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds (Appendix A1). key: uint32[2]; x0,x1: uint32 arrays."""
    key = np.asarray(key, dtype=_U32)
    k0, k1 = key[..., 0], key[..., 1]  # scalars, or arrays broadcasting against x0/x1
    ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
    shape = np.broadcast_shapes(np.shape(x0), np.shape(x1), np.shape(k0))
    x0 = np.broadcast_to(np.asarray(x0, dtype=_U32), shape).copy()
    x1 = np.broadcast_to(np.asarray(x1, dtype=_U32), shape).copy()
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            rots = _R0 if i % 2 == 0 else _R1
            for r in rots:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def prng_key(seed: int):
    """PRNGKey(seed) = [seed >> 32, seed & 0xffffffff] (Appendix A2)."""
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key, num: int = 2):
    """jax.random.split(key, num) -> uint32[num, 2] (Appendix A2)."""
    if PARTITIONABLE:
        idx = np.arange(num, dtype=np.uint64)
        hi = (idx >> np.uint64(32)).astype(_U32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(_U32)
        o0, o1 = threefry2x32(key, hi, lo)
        return np.stack([o0, o1], axis=-1)
    counts = np.arange(2 * num, dtype=_U32)
    o0, o1 = threefry2x32(key, counts[:num], counts[num:])
    return np.concatenate([o0, o1]).reshape(num, 2)


def random_bits(key, shape):
    """32-bit random bits of `shape` (Appendix A2)."""
    size = int(np.prod(shape)) if len(shape) else 1
    if PARTITIONABLE:
        idx = np.arange(size, dtype=np.uint64)
        hi = (idx >> np.uint64(32)).astype(_U32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(_U32)
        o0, o1 = threefry2x32(key, hi, lo)
        return (o0 ^ o1).reshape(shape)
    n = size + (size % 2)
    counts = np.arange(n, dtype=_U32)
    o0, o1 = threefry2x32(key, counts[: n // 2], counts[n // 2 :])
    return np.concatenate([o0, o1])[:size].reshape(shape)


def uniform(key, shape, minval=0.0, maxval=1.0):
    """jax.random.uniform float32 (Appendix A3)."""
    bits = random_bits(key, shape)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, f * (maxval - minval) + minval).astype(np.float32)


def gumbel(key, shape):
    """jax.random.gumbel float32: -log(-log(uniform(tiny, 1))) (Appendix A3)."""
    u = uniform(key, shape, minval=np.finfo(np.float32).tiny, maxval=1.0)
    return (-np.log(-np.log(u))).astype(np.float32)


def randint(key, shape, minval: int, maxval: int):
    """jax.random.randint int32 (Appendix A4)."""
    k1, k2 = split(key)
    hi = random_bits(k1, shape)
    lo = random_bits(k2, shape)
    span = _U32(max(1, maxval - minval)) if maxval > minval else _U32(1)
    mult = _U32(((65536 % int(span)) ** 2) % int(span))
    with np.errstate(over="ignore"):
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int32(minval) + off.astype(np.int32)).astype(np.int32)


def permutation_rounds(n: int) -> int:
    return int(math.ceil(3 * math.log(max(1, n)) / math.log(2**32 - 1)))


def permutation(key, n: int):
    """jax.random.permutation(key, n) via repeated stable sort by random bits (Appendix A5)."""
    x = np.arange(n, dtype=np.int32)
    for _ in range(permutation_rounds(n)):
        ks = split(key)
        key, sub = ks[0], ks[1]
        sort_keys = random_bits(sub, (n,))
        x = x[np.argsort(sort_keys, kind="stable")]
    return x


def categorical_distrax(key, logits):
    """distrax.Categorical(logits).sample_and_log_prob(seed=key) (Appendix A3).

    logits float32[..., a] (already masked). Noise tensor has shape (1,)+logits.shape,
    i.e. linear index = row*a + j. Returns (int32 sample, float32 log_prob, normalized logits).
    """
    logits = np.asarray(logits, dtype=np.float32)
    m = logits.max(axis=-1, keepdims=True)
    lse = (np.log(np.exp(logits - m).sum(axis=-1, keepdims=True, dtype=np.float32)) + m).astype(np.float32)
    norm = (logits - lse).astype(np.float32)
    g = gumbel(key, logits.shape)
    sample = np.argmax(g + norm, axis=-1).astype(np.int32)
    logp = np.take_along_axis(norm, sample[..., None].astype(np.int64), axis=-1)[..., 0]
    return sample, logp.astype(np.float32), norm


# ---- batched helpers (a vmap over keys), used by the vectorised env oracle ----
def split_batched(keys, num: int = 2):
    """vmap(split)(keys): keys uint32[B,2] -> uint32[B,num,2]."""
    assert PARTITIONABLE, "batched helpers implement the partitionable scheme only"
    keys = np.asarray(keys, dtype=_U32)
    lo = np.arange(num, dtype=_U32)[None, :]
    o0, o1 = threefry2x32(keys[:, None, :], np.zeros_like(lo), lo)
    return np.stack([o0, o1], axis=-1)


def random_bits_batched(keys, n: int):
    """vmap(lambda k: random_bits(k, (n,)))(keys) -> uint32[B,n]."""
    assert PARTITIONABLE
    keys = np.asarray(keys, dtype=_U32)
    lo = np.arange(n, dtype=_U32)[None, :]
    o0, o1 = threefry2x32(keys[:, None, :], np.zeros_like(lo), lo)
    return o0 ^ o1


def randint_batched(keys, n: int, minval: int, maxval: int):
    """vmap(lambda k: randint(k, (n,), minval, maxval))(keys) -> int32[B,n]."""
    ks = split_batched(keys, 2)
    hi = random_bits_batched(ks[:, 0], n)
    lo = random_bits_batched(ks[:, 1], n)
    span = _U32(max(1, maxval - minval))
    mult = _U32(((65536 % int(span)) ** 2) % int(span))
    with np.errstate(over="ignore"):
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int32(minval) + off.astype(np.int32)).astype(np.int32)


# ---- jax.random.fold_in / normal / truncated_normal (parameter initialisers; checked against the library's kernels) ----
def fold_in(key, data: int):
    """jax.random.fold_in(key, data) = threefry2x32(key, threefry_seed(data)) with threefry_seed(uint32 d) = (0, d)."""
    key = np.asarray(key, dtype=_U32)
    o0, o1 = threefry2x32(key, np.zeros(1, _U32), np.array([data], _U32))
    return np.array([o0[0], o1[0]], dtype=_U32)


def erf_inv_f32(x):
    """XLA's single-precision erf_inv expansion (two degree-8 polynomials in w = -log1p(-x^2))."""
    f = np.float32
    x = np.asarray(x, f)
    w = (-np.log1p(-x * x)).astype(f)
    lt = w < f(5.0)
    w = np.where(lt, w - f(2.5), np.sqrt(w) - f(3.0)).astype(f)
    lo = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
    hi = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]
    p = np.where(lt, f(lo[0]), f(hi[0])).astype(f)
    for a, b in zip(lo[1:], hi[1:]):
        p = (np.where(lt, f(a), f(b)) + p * w).astype(f)
    return (p * x).astype(f)


def _uniform_range(key, shape, minval, maxval):
    f = np.float32
    bits = random_bits(key, shape)
    u = ((bits >> _U32(9)) | _U32(0x3F800000)).view(f) - f(1.0)
    return np.maximum(f(minval), (u * (f(maxval) - f(minval)) + f(minval)).astype(f))


def normal(key, shape):
    """jax.random.normal(key, shape, float32) (jax 0.6.0 `_normal_real`)."""
    f = np.float32
    u = _uniform_range(key, shape, np.nextafter(f(-1.0), f(0.0)), f(1.0))
    return (f(np.sqrt(2.0)) * erf_inv_f32(u)).astype(f)


def truncated_normal(key, lower, upper, shape):
    """jax.random.truncated_normal(key, lower, upper, shape, float32)."""
    import math

    f = np.float32
    s2 = f(np.sqrt(2.0))
    a, b = f(math.erf(float(f(lower) / s2))), f(math.erf(float(f(upper) / s2)))
    out = (s2 * erf_inv_f32(_uniform_range(key, shape, a, b))).astype(f)
    return np.clip(out, np.nextafter(f(lower), f(np.inf)), np.nextafter(f(upper), f(-np.inf)))
