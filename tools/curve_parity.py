#!/usr/bin/env python
"""Learning-curve parity of the CUDA path against the CPU oracle on BASELINE.json configs[0] (CoordSum 3x10-30, num_envs=16,
rollout_length=128, U=2, P=4, M=2) over several seeds: per update the fraction of identical sampled actions, the mean reward per
agent-step, the six loss terms of both sides and the largest parameter deviation relative to the parameter scale.
Usage (GPU box): python tools/curve_parity.py --seeds 5 --updates 8 --out gpurun_out/curve_parity.json [--env lbf]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magpo_b200.learner import CoordSumVec, LbfVec, MagpoLearner, SystemConfig  # noqa: E402
from oracle import coordsum as ocs, lbf as olbf, learner as olr, nets as onets, prng as oprng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=5)
ap.add_argument("--updates", type=int, default=8)
ap.add_argument("--num-envs", type=int, default=16)
ap.add_argument("--rollout-length", type=int, default=128)
ap.add_argument("--env", default="coordsum", choices=["coordsum", "lbf"])
ap.add_argument("--out", default="gpurun_out/curve_parity.json")
args = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)
E, U, T = args.num_envs, 2, args.rollout_length
if args.env == "lbf":
    kw = olbf.SCENARIOS["2s-8x8-2p-2f-coop"]
    spec, vec = olbf.LbfSpec(**kw), LbfVec(**kw)
else:
    kw = ocs.SCENARIOS["3x10-30-v0"]
    spec, vec = ocs.CoordSumSpec(**kw), CoordSumVec(**kw)
ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T)
NAMES = ("total_loss", "value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy")
runs = []
for seed in range(42, 42 + args.seeds):
    state = olr.learner_setup(spec, ncfg, osys, seed=seed, param_seed=seed)
    lrn = MagpoLearner(vec, SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T), device="cuda:0")
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(seed), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    rows = []
    for upd in range(args.updates):
        t0 = time.time()
        rec = {}
        _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
        t_cpu = time.time() - t0
        _, losses = lrn.update_step()
        torch.cuda.synchronize()
        li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
        act = lrn.traj["action"].cpu().numpy()
        rew = lrn.traj["reward"].cpu().numpy()
        same = np.mean([(act[:, u * E:(u + 1) * E] == rec["traj"][u]["action"]).mean() for u in range(U)])
        r_ref = float(np.mean([rec["traj"][u]["reward"].mean() for u in range(U)]))
        gp, apar = lrn.get_params()
        dev = 0.0
        for new, ref in ((gp, state["guider_params"]), (apar, state["actor_params"])):
            for name, r in ref.items():
                dev = max(dev, float(np.abs(new[name].cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-3)))
        row = dict(update=upd, actions_identical=float(same), reward_cuda=float(rew.mean()), reward_oracle=r_ref, max_param_rel_dev=dev,
                   oracle_s=round(t_cpu, 2))
        for n in NAMES:
            row[n + "_cuda"] = float(li[n].mean())
            row[n + "_oracle"] = float(np.mean([i[n] for i in infos]))
        rows.append(row)
        print(seed, json.dumps(row), flush=True)
    runs.append(dict(seed=seed, updates=rows))
summary = dict(config=dict(env=args.env, num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=4, num_minibatches=2),
               runs=runs)
with open(args.out, "w") as f:
    json.dump(summary, f, indent=1)
