#!/usr/bin/env python
"""One minibatch of the update (pack, guider/learner forward + backward, clip+Adam) at the bench size between
cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists and --set full captures.
Usage: [ncu ... --profile-from-start off] python tools/profile_update.py [--num-envs 4096] [--rollout-steps 0]
--rollout-steps K > 0 profiles a K-step rollout instead of the minibatch."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpo_b200 import init as minit  # noqa: E402
from magpo_b200.learner import CoordSumVec, LbfVec, MagpoLearner, SystemConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="lbf", choices=["lbf", "coordsum"])
ap.add_argument("--num-envs", type=int, default=4096)
ap.add_argument("--update-batch-size", type=int, default=2)
ap.add_argument("--rollout-length", type=int, default=128)
ap.add_argument("--rollout-steps", type=int, default=0)
ap.add_argument("--gae", action="store_true", help="profile the GAE kernel on a full-length trajectory instead")
args = ap.parse_args()
dev = torch.device("cuda:0")
env = LbfVec() if args.env == "lbf" else CoordSumVec(num_agents=3, num_actions=10, time_limit=100, maxval=30)
T = args.rollout_steps or args.rollout_length
sysc = SystemConfig(num_envs=args.num_envs, update_batch_size=args.update_batch_size, rollout_length=T)
lrn = MagpoLearner(env, sysc, device=dev, graph_rollout=False)
lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 0), minit.init_actor(env.obs_dim, env.action_dim, 1))
env_keys, step_key, _ = minit.setup_keys(42, 1, sysc.update_batch_size, sysc.num_envs, dev)
lrn.reset(env_keys[0], step_key)
lrn.rollout(); lrn.gae(); lrn.epoch_indices(True)
lrn.minibatch_grads(0); lrn.apply_grads()
torch.cuda.synchronize()
torch.cuda.profiler.start()
if args.gae:
    lrn.gae()
elif args.rollout_steps:
    lrn.rollout()
else:
    lrn.minibatch_grads(1); lrn.apply_grads()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
