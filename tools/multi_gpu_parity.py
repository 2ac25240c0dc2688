#!/usr/bin/env python
"""Data-parallel parity: N ranks x (U=1 slot of E envs) against one process with U=N slots of the same envs.

The reference averages gradients over the "batch" (slot) axis and then over the "device" axis (rec_magpo.py:395-409); the env keys
are laid out (device, slot, env) (:642-653) and the step key is shared. So N devices with one slot each and one device with N
slots see the same envs, sample the same actions and must reach the same parameters after an update (up to fp32 summation order).
Run: torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_parity.py
Rank 0 prints one JSON line (and writes gpurun_out/multi_gpu_parity.json)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magpo_b200 import init as minit  # noqa: E402
from magpo_b200.comm import NcclComm  # noqa: E402
from magpo_b200.learner import LbfVec, MagpoLearner, SystemConfig  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)  # harness only (gathering the ranks' results); the gradient exchange is NcclComm's
comm = NcclComm.from_env(dev)
E, T, UPDATES = 64, 32, 3
env = LbfVec()


def make(U, n_dev, r, attach):
    sysc = SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T)
    lrn = MagpoLearner(env, sysc, device=dev, world_size=n_dev)
    if attach:
        comm.attach(lrn)  # magpo_minibatch_grads all-reduces through the library's communicator
    lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 0), minit.init_actor(env.obs_dim, env.action_dim, 1))
    env_keys, step_key, _ = minit.setup_keys(42, n_dev, U, E, dev)
    lrn.reset(env_keys[r], step_key)
    return lrn


dp = make(1, world, rank, True)
single = make(world, 1, 0, False) if rank == 0 else None
rows = []
for upd in range(UPDATES):
    dp.update_step()
    torch.cuda.synchronize()
    acts = [torch.zeros_like(dp.traj["action"]) for _ in range(world)]
    dist.all_gather(acts, dp.traj["action"].contiguous())
    if rank == 0:
        single.update_step()
        torch.cuda.synchronize()
        ref = single.traj["action"]  # [T, world*E, A], slot-major
        same = float(np.mean([(acts[r] == ref[:, r * E:(r + 1) * E]).float().mean().item() for r in range(world)]))
        gp, ap = dp.get_params()
        gs, as_ = single.get_params()
        dev_max = 0.0
        for a, b in ((gp, gs), (ap, as_)):
            for k in a:
                dev_max = max(dev_max, float((a[k] - b[k]).abs().max() / max(float(b[k].abs().max()), 1e-3)))
        rows.append(dict(update=upd, actions_identical=same, max_param_rel_dev=dev_max))
# every rank must hold the same parameters
flat = torch.cat([dp.guider, dp.actor])
ref = flat.clone()
dist.broadcast(ref, 0)
replica_dev = torch.tensor([float((flat - ref).abs().max())], device=dev)
dist.all_reduce(replica_dev, op=dist.ReduceOp.MAX)
if rank == 0:
    out = dict(world=world, envs_per_rank=E, rollout_length=T, updates=rows, max_abs_param_diff_between_ranks=float(replica_dev))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "multi_gpu_parity.json"), "w"), indent=1)
    print(json.dumps(out))
dist.barrier()
comm.close()
dist.destroy_process_group()
