"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs. the CPU oracle)."""
import ctypes as C

import numpy as np
import torch

from magpo_b200 import _lib as L


def dt(x, dev, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev).contiguous()


def u32(x, dev):
    """uint32 numpy -> int32 device tensor with the same bits."""
    return dt(np.asarray(x, np.uint32).view(np.int32), dev)


def as_u32(t):
    return t.cpu().numpy().view(np.uint32)


def to_time_major(x, T, A):
    """oracle [N, T*A, ...] -> CUDA [T, N, A, ...]"""
    x = np.asarray(x)
    N = x.shape[0]
    return np.ascontiguousarray(np.swapaxes(x.reshape(N, T, A, *x.shape[2:]), 0, 1))


def from_time_major(x):
    """CUDA [T, N, A, ...] -> oracle [N, T*A, ...]"""
    x = np.asarray(x)
    T, N, A = x.shape[:3]
    return np.swapaxes(x, 0, 1).reshape(N, T * A, *x.shape[3:])


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def sync():
    torch.cuda.synchronize()
