#!/usr/bin/env python
"""Forward / gradient errors of the general guider path per network shape (tests/test_gpu_general_shapes.py without the asserts).
Usage (GPU box): python tools/diag_general.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_networks as tgn  # noqa: E402
from gpu_util import from_time_major, rel_err, sync  # noqa: E402
from magpo_b200 import _lib as L  # noqa: E402
from oracle import nets as onets  # noqa: E402

dev = torch.device("cuda:0")
A, d, a, T, N = 3, 9, 7, 10, 5
for shape in [(32, 4, 1), (32, 4, 2), (32, 2, 1), (64, 4, 1), (128, 4, 1), (32, 1, 2), (64, 2, 2)]:
    cfg, net, gp, ap, (gt, ng, gflat), _ = tgn.setup_nets(A, d, a, dev, shape=shape)
    mb = tgn.make_case(1, 1, N, T, A, d, a, shape=shape)
    mbs, keep = tgn.device_minibatch(mb, T, A, dev)
    ws, nbytes = tgn.workspace(net, T, N, dev)
    value, logits = torch.zeros(T, N, A, device=dev), torch.zeros(T, N, A, a, device=dev)
    L.call("magpo_guider_forward", L.context(), L.stream_ptr(), C.byref(net.c_struct()), L.ptr(gflat), mbs, L.ptr(value), L.ptr(logits),
           L.ptr(ws), C.c_size_t(nbytes))
    outs = {}
    for dt_ in (torch.float32, torch.float64):
        p = onets.to_torch(gp, dt_)
        v_ref, _, _, l_ref = onets.sable_apply(p, cfg, torch.tensor(mb["obs"], dtype=dt_), torch.tensor(mb["action_mask"]),
                                               torch.tensor(mb["step_count"]), torch.tensor(mb["action"]),
                                               tuple(torch.tensor(h, dtype=dt_) for h in mb["prev_hstates"]), torch.tensor(mb["done"]), T)
        outs[dt_] = (v_ref.numpy(), l_ref.numpy())
    sync()
    lg, legal = from_time_major(logits.cpu().numpy()), mb["action_mask"]
    v = from_time_major(value.cpu().numpy())
    print(shape, "value: cuda-vs-o64 %.2e  o32-vs-o64 %.2e | logits: cuda-vs-o64 %.2e  o32-vs-o64 %.2e" % (
        rel_err(v, outs[torch.float64][0]), rel_err(outs[torch.float32][0], outs[torch.float64][0]),
        rel_err(lg[legal], outs[torch.float64][1][legal]), rel_err(outs[torch.float32][1][legal], outs[torch.float64][1][legal])), flush=True)
for shape, Ns in [((128, 4, 1), 4), ((128, 4, 1), 40), ((128, 1, 1), 40), ((32, 4, 2), 6)]:
    try:
        tgn.test_minibatch_grads(dev, 3, 9, 7, 8, Ns, 2, 2e-4, shape=shape)
        print(shape, Ns, "grads pass")
    except AssertionError as e:
        rep = open(os.path.join(ROOT, "gpurun_out", "grad_report_A3_U2.txt")).read().splitlines()
        rows = sorted((l for l in rep if not l.startswith("loss")), key=lambda l: -float(l.split()[0]))
        print(shape, Ns, "grads FAIL", str(e)[:100])
        for l in rows[:6]:
            print("    ", l)
