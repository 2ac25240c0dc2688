#!/usr/bin/env python
"""A/B of the persistent tensor-core GRU scans against the per-timestep path: gradients of one minibatch and timing.
Usage: python tools/diag_gru.py [--num-envs 50] [--rollout-length 9]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpo_b200 import _lib as L  # noqa: E402
from magpo_b200 import init as minit  # noqa: E402
from magpo_b200.learner import CoordSumVec, MagpoLearner, SystemConfig, param_views  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--num-envs", type=int, default=50)
ap.add_argument("--update-batch-size", type=int, default=2)
ap.add_argument("--rollout-length", type=int, default=9)
args = ap.parse_args()
dev = torch.device("cuda:0")
env = CoordSumVec(num_agents=3, num_actions=10, time_limit=100, maxval=30)
sysc = SystemConfig(num_envs=args.num_envs, update_batch_size=args.update_batch_size, rollout_length=args.rollout_length,
                    num_minibatches=1)
lrn = MagpoLearner(env, sysc, device=dev)
lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 0), minit.init_actor(env.obs_dim, env.action_dim, 1))
env_keys, step_key, _ = minit.setup_keys(42, 1, sysc.update_batch_size, sysc.num_envs, dev)
lrn.reset(env_keys[0], step_key)
lib = L.lib()
lrn.rollout(); lrn.gae(); lrn.epoch_indices(True)
out = {}
for mode in (1, 0):
    lib.magpo_debug_force_gru_stepwise(mode)
    lrn.minibatch_grads(0)
    torch.cuda.synchronize()
    out[mode] = lrn.grads.clone()
ga = {m: param_views(out[m][lrn.n_g:lrn.n_g + lrn.n_a], lrn.a_table) for m in out}
for k in ga[0]:
    a, b = ga[0][k], ga[1][k]
    print(f"{k:50s} |stepwise|max {b.abs().max().item():.3e}  max diff {(a - b).abs().max().item():.3e}")
print("losses stepwise", out[1][-8:].tolist())
print("losses scan    ", out[0][-8:].tolist())

import ctypes as C
CATS = ["gemm_nn", "gemm_tn", "colsum", "rowops", "retention_fwd", "retention_bwd", "gru_pointwise", "loss", "pack", "optim"]
for mode in (1, 0):
    lib.magpo_debug_force_gru_stepwise(mode)
    lrn.minibatch_grads(0)
    torch.cuda.synchronize()
    lib.magpo_prof_enable(1)
    lrn.minibatch_grads(0)
    torch.cuda.synchronize()
    lib.magpo_prof_enable(0)
    line = []
    for i, name in enumerate(CATS):
        ms, work, cnt = C.c_double(), C.c_double(), C.c_int64()
        lib.magpo_prof_read(i, C.byref(ms), C.byref(work), C.byref(cnt))
        line.append(f"{name} {ms.value:.3f}ms/{cnt.value}")
    print("stepwise" if mode else "scan    ", " | ".join(line))

if os.environ.get("GRU_TIMELINE"):
    import numpy as np
    buf = torch.zeros(3 * 1024, dtype=torch.int64, device=dev)
    lib.magpo_debug_force_gru_stepwise(0)
    lib.magpo_debug_gru_timeline(C.c_void_p(buf.data_ptr()))
    lrn.minibatch_grads(0)
    torch.cuda.synchronize()
    lib.magpo_debug_gru_timeline(None)
    b = buf.cpu().numpy().reshape(3, 1024)
    t0 = b[b > 0].min()
    rel = lambda x: (x - t0) / 1000.0 if x > 0 else float("nan")
    for t in range(1, 4):
        print(f"--- step {t}  (us since first stamp)")
        print("  MMA: a_ready %.2f | " % rel(b[1, t * 16]) + " | ".join(f"jb{jb}: first B piece {rel(b[1, t*16+1+2*jb]):.2f} commit issued {rel(b[1, t*16+2+2*jb]):.2f}" for jb in range(4)))
        for jb in range(4):
            g = b[2, (t * 4 + jb) * 6:(t * 4 + jb) * 6 + 6]
            print(f"  gates jb{jb}: loads issued {rel(g[0]):.2f}  acc ready {rel(g[1]):.2f}  tmem read {rel(g[2]):.2f}  math done {rel(g[3]):.2f}  stores issued {rel(g[4]):.2f}  A written {rel(g[5]):.2f}")
        print("  TMA waits passed: " + " ".join(f"{rel(b[0, t*16+i]):.2f}" for i in range(16)))
