#!/usr/bin/env python
"""Phase timing at the bench workload (CUDA events): rollout, GAE, and one minibatch's guider / learner forward and
backward (by skipping parts of magpo_minibatch_grads through the debug hook).
Usage: python tools/bench_phases.py [--num-envs 4096] [--rollout-length 128]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpo_b200 import _lib as L  # noqa: E402
from magpo_b200 import init as minit  # noqa: E402
from magpo_b200.learner import CoordSumVec, MagpoLearner, SystemConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--update-batch-size", type=int, default=2)
ap.add_argument("--rollout-length", type=int, default=128)
ap.add_argument("--chunk-envs", type=int, default=4096)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
env = CoordSumVec(num_agents=3, num_actions=10, time_limit=100, maxval=30)
sysc = SystemConfig(num_envs=args.num_envs, update_batch_size=args.update_batch_size, rollout_length=args.rollout_length,
                    chunk_envs=args.chunk_envs)
lrn = MagpoLearner(env, sysc, device=dev)
lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 0), minit.init_actor(env.obs_dim, env.action_dim, 1))
env_keys, step_key, _ = minit.setup_keys(42, 1, sysc.update_batch_size, sysc.num_envs, dev)
lrn.reset(env_keys[0], step_key)
lib = L.lib()


def timeit(fn, reps=args.reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"E={args.num_envs} U={args.update_batch_size} T={args.rollout_length} chunk={lrn.chunk}")
print(f"rollout+bootstrap      {timeit(lrn.rollout):9.3f} ms")
import time
torch.cuda.synchronize()
t0 = time.perf_counter(); lrn.rollout(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"  host enqueue time of one rollout call {1e3 * (t1 - t0):9.3f} ms (then {1e3 * (t2 - t1):.3f} ms until the GPU drained)")
print(f"gae                    {timeit(lrn.gae):9.3f} ms")
lrn.epoch_indices(True)
res = {}
for name, mask in (("all", 0), ("no guider", 1), ("no learner", 2), ("no guider bwd", 4), ("no learner bwd", 8), ("losses+pack only", 3)):
    lib.magpo_debug_set_skip(mask)
    res[name] = timeit(lambda: lrn.minibatch_grads(0))
lib.magpo_debug_set_skip(0)
base = res["losses+pack only"]
print(f"minibatch_grads        {res['all']:9.3f} ms   (x{sysc.ppo_epochs * sysc.num_minibatches} per step)")
print(f"  pack + stats + loss  {base:9.3f} ms")
print(f"  guider fwd           {res['all'] - res['no guider'] - (res['all'] - res['no guider bwd']):9.3f} ms")
print(f"  guider bwd           {res['all'] - res['no guider bwd']:9.3f} ms")
print(f"  learner fwd          {res['all'] - res['no learner'] - (res['all'] - res['no learner bwd']):9.3f} ms")
print(f"  learner bwd          {res['all'] - res['no learner bwd']:9.3f} ms")
lib.magpo_debug_force_gru_stepwise(1)
lib.magpo_debug_set_skip(1)
a = timeit(lambda: lrn.minibatch_grads(0))
lib.magpo_debug_set_skip(1 | 8)
b = timeit(lambda: lrn.minibatch_grads(0))
lib.magpo_debug_set_skip(3)
c = timeit(lambda: lrn.minibatch_grads(0))
print(f"  (per-timestep GRU path: learner fwd {b - c:9.3f} ms, bwd {a - b:9.3f} ms)")
lib.magpo_debug_force_gru_stepwise(0)
lib.magpo_debug_set_skip(0)
print(f"apply_grads            {timeit(lrn.apply_grads):9.3f} ms")
