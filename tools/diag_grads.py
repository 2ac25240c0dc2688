#!/usr/bin/env python
"""Bisect of a minibatch-gradient mismatch: runs tests/test_gpu_networks.py::test_minibatch_grads for one shape with the
tensor-core GEMMs / chunkwise retention / persistent GRU scans switched off one at a time and prints the worst tensors.
Usage: python tools/diag_grads.py A d a T Ns U"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from magpo_b200 import _lib as L  # noqa: E402
import test_gpu_networks as tgn  # noqa: E402

shape = [int(x) for x in sys.argv[1:7]] if len(sys.argv) >= 7 else [3, 4, 10, 9, 50, 2]
dev = torch.device("cuda:0")
lib = L.lib()
modes = [("default", {}), ("all off", dict(tc=0, scan=1, gru=1))]
base = list(shape)
for name, m, ns in [(n, m, ns) for ns in (5, 12, 25, 50, 100) for n, m in modes]:
    shape = base[:4] + [ns] + base[5:]
    name = f"Ns={ns} {name}"
    lib.magpo_set_tensor_cores(m.get("tc", 1))
    lib.magpo_debug_force_retention_scan(m.get("scan", 0))
    lib.magpo_debug_force_gru_stepwise(m.get("gru", 0))
    try:
        tgn.test_minibatch_grads.__wrapped__(dev, *shape) if hasattr(tgn.test_minibatch_grads, "__wrapped__") else \
            tgn.test_minibatch_grads(dev, *shape, 2e-4)
        verdict = "pass"
    except AssertionError as e:
        verdict = f"FAIL {str(e)[:80]}"
    A, U = shape[0], shape[5]
    rep = open(os.path.join(ROOT, "gpurun_out", f"grad_report_A{A}_U{U}.txt")).read().splitlines()
    rows = sorted((l for l in rep if not l.startswith("loss")), key=lambda l: -float(l.split()[0]))
    print(f"=== {name}: {verdict}")
    for l in rows[:4]:
        print("   ", l)
