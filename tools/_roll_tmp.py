import sys, os, time
sys.path.insert(0, '.')
import torch
from magpo_b200 import init as minit, _lib as L
from magpo_b200.learner import LbfVec, MagpoLearner, SystemConfig
dev = torch.device("cuda:0")
env = LbfVec()
def run(overlap, sable_only, graph=True):
    L.lib().magpo_debug_set_overlap(overlap)
    sysc = SystemConfig(num_envs=4096, update_batch_size=2, rollout_length=128, sable_only=sable_only) if sable_only else SystemConfig(num_envs=4096, update_batch_size=2, rollout_length=128)
    lrn = MagpoLearner(env, sysc, device=dev, graph_rollout=graph)
    lrn.set_params(minit.init_guider(env.num_agents, env.obs_dim, env.action_dim, 0), minit.init_actor(env.obs_dim, env.action_dim, 1))
    env_keys, step_key, _ = minit.setup_keys(42, 1, sysc.update_batch_size, sysc.num_envs, dev)
    lrn.reset(env_keys[0], step_key)
    for _ in range(3): lrn.rollout()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lrn.rollout()
    e1.record(); torch.cuda.synchronize()
    print("overlap", overlap, "sable_only", sable_only, "graph", graph, "rollout ms", e0.elapsed_time(e1) / 5, flush=True)
    del lrn
run(1, False); run(0, False)
try:
    run(1, True)
except Exception as ex:
    print("sable_only failed", ex)
