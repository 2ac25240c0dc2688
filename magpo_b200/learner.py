"""Device-resident state and the rollout+update step of rec_magpo on one GPU.

Mirrors `_update_step` of `mava/systems/gpo/anakin/rec_magpo.py:106-499` for the U update-batch slots
of one device at once (B = U*E envs, slot-major); the data-parallel mean over devices (the
`pmean(..., "device")` of :399-409) is one all-reduce of the flat gradient buffer supplied by the caller.
torch only owns memory/streams here; all compute is in libmagpo_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib as L


@dataclass
class SystemConfig:
    """configs/system/gpo/rec_magpo.yaml:12-25 + arch.num_envs."""
    num_envs: int = 16
    update_batch_size: int = 2
    rollout_length: int = 128
    ppo_epochs: int = 4
    num_minibatches: int = 2
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_eps: float = 0.2
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    clip_gpo: float = 1.5
    alpha: float = 1.0
    actor_lr: float = 2.5e-4
    chunk_envs: int = 0  # envs differentiated per pass (0 = whole minibatch); gradients add up exactly
    sable_only: bool = False  # rec_sable (systems/sable/anakin/rec_sable.py): the guider network alone under PPO
    decay_learning_rates: bool = False  # utils/training.py:48-64: linear schedule over num_updates (needs num_updates)
    num_updates: int = 0

    def c_struct(self) -> L.SysCfg:
        return L.SysCfg(self.num_envs, self.update_batch_size, self.rollout_length, self.ppo_epochs,
                        self.num_minibatches, self.gamma, self.gae_lambda, self.clip_eps, self.ent_coef,
                        self.vf_coef, self.max_grad_norm, self.clip_gpo, self.alpha, self.actor_lr, int(self.sable_only))


@dataclass
class NetworkConfig:
    """configs/network/magpo.yaml + env dims."""
    n_agents: int
    obs_dim: int
    action_dim: int
    max_step_count: int
    embed_dim: int = 64
    n_head: int = 1
    n_block: int = 1
    hidden: int = 128
    timestep_pe: bool = True
    decay_scaling_factor: float = 0.8

    @property
    def state_shape(self) -> tuple:
        """One env's Sable hidden state: [n_head, n_block, head_size, head_size] (get_init_hstates.py:20-43)."""
        hs = self.embed_dim // self.n_head
        return (self.n_head, self.n_block, hs, hs)

    def c_struct(self) -> L.NetCfg:
        return L.NetCfg(self.n_agents, self.obs_dim, self.action_dim, self.embed_dim, self.n_head, self.n_block,
                        self.hidden, int(self.timestep_pe), self.decay_scaling_factor, self.max_step_count)


def param_table(net: NetworkConfig, which: int):
    """[(flax path, offset, dim0, dim1, ld)] of network `which` (0 guider, 1 learner) and the flat length."""
    lib = L.lib()
    cfg = net.c_struct()
    n = lib.magpo_param_num_tensors(C.byref(cfg), which)
    if n < 0:
        raise L.MagpoError("unsupported network configuration")
    out = []
    for i in range(n):
        name, off = C.c_char_p(), C.c_int64()
        d0, d1, ld = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(lib.magpo_param_tensor(C.byref(cfg), which, i, C.byref(name), C.byref(off), C.byref(d0), C.byref(d1),
                                       C.byref(ld)), "magpo_param_tensor")
        out.append((name.value.decode(), off.value, d0.value, d1.value, ld.value))
    return out, int(lib.magpo_param_count(C.byref(cfg), which))


def param_views(flat: torch.Tensor, table) -> dict:
    """The flax-named tensors as (strided) views of the flat buffer."""
    views = {}
    for name, off, d0, d1, ld in table:
        if d1 == 0:
            views[name] = flat[off:off + d0]
        else:
            views[name] = torch.as_strided(flat, (d0, d1), (ld, 1), flat.storage_offset() + off)
    return views


def load_params(flat: torch.Tensor, table, params: dict) -> None:
    views = param_views(flat, table)
    missing = set(views) - set(params)
    if missing:
        raise KeyError(f"missing parameters: {sorted(missing)}")
    for k, v in views.items():
        v.copy_(torch.as_tensor(np.asarray(params[k]), dtype=torch.float32).reshape(v.shape))


class CoordSumVec:
    """B CoordSum envs under RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(CoordSumWrapper(.))))
    (mava/coordsum/env.py, mava/utils/make_env.py:90-104,202-218) as device arrays."""

    kind = L.ENV_COORDSUM
    reset_fn, step_fn = "magpo_coordsum_reset", "magpo_coordsum_step"

    def __init__(self, num_agents: int, num_actions: int, time_limit: int = 100, maxval: int | None = None):
        self.num_agents, self.num_actions, self.time_limit = num_agents, num_actions, time_limit
        self.maxval = maxval or num_actions
        self.cfg = L.CoordSumCfg(num_agents, num_actions, time_limit, self.maxval)

    @property
    def obs_dim(self):
        return self.num_agents + 1

    @property
    def action_dim(self):
        return self.num_actions

    def alloc_state(self, B: int, dev) -> dict:
        i32, f32, u32 = torch.int32, torch.float32, torch.int32  # uint32 keys are stored in int32 tensors
        TL, a = self.time_limit, self.num_actions
        return dict(step_count=torch.zeros(B, dtype=i32, device=dev), target=torch.zeros(B, TL + 1, dtype=i32, device=dev),
                    record=torch.zeros(B, a, TL, dtype=i32, device=dev), key=torch.zeros(B, 2, dtype=u32, device=dev),
                    metrics_key=torch.zeros(B, 2, dtype=u32, device=dev),
                    running_return=torch.zeros(B, dtype=f32, device=dev), running_length=torch.zeros(B, dtype=i32, device=dev),
                    episode_return=torch.zeros(B, dtype=f32, device=dev), episode_length=torch.zeros(B, dtype=i32, device=dev))

    def state_struct(self, st: dict) -> L.CoordSumState:
        return L.struct_of(L.CoordSumState, **st)


class LbfVec:
    """B LevelBasedForaging envs (jumanji 1.1.0 `RandomGenerator(**task_config)`, `time_limit` from env.kwargs) under
    RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(LbfWrapper(.)))) (mava/utils/make_env.py:90-135, wrappers/jumanji.py:171-208)
    as device arrays. The dynamics are restated from the un-vendored dependency; see oracle/lbf.py."""

    kind = L.ENV_LBF
    reset_fn, step_fn = "magpo_lbf_reset", "magpo_lbf_step"

    def __init__(self, grid_size: int = 8, fov: int = 2, num_agents: int = 2, num_food: int = 2, max_agent_level: int = 2,
                 force_coop: bool = True, time_limit: int = 100, agent_mask_rows: bool = True):
        self.grid_size, self.fov, self.num_agents, self.num_food = grid_size, fov, num_agents, num_food
        self.max_agent_level, self.force_coop, self.time_limit = max_agent_level, bool(force_coop), time_limit
        self.cfg = L.LbfCfg(grid_size, fov, num_agents, num_food, max_agent_level, int(bool(force_coop)), time_limit,
                            int(bool(agent_mask_rows)))

    @property
    def obs_dim(self):
        return self.num_agents + 3 * (self.num_food + self.num_agents)

    @property
    def action_dim(self):
        return 6

    def alloc_state(self, B: int, dev) -> dict:
        i32, f32, u8 = torch.int32, torch.float32, torch.uint8
        A, F = self.num_agents, self.num_food
        z = lambda *s, dt=i32: torch.zeros(*s, dtype=dt, device=dev)
        return dict(agent_pos=z(B, A, 2), agent_level=z(B, A), agent_loading=z(B, A, dt=u8), food_pos=z(B, F, 2),
                    food_level=z(B, F), food_eaten=z(B, F, dt=u8), step_count=z(B), key=z(B, 2), metrics_key=z(B, 2),
                    running_return=z(B, dt=f32), running_length=z(B), episode_return=z(B, dt=f32), episode_length=z(B))

    def state_struct(self, st: dict) -> L.LbfState:
        return L.struct_of(L.LbfState, **st)


class RwareVec:
    """B RobotWarehouse envs (jumanji 1.1.0 `RandomGenerator(**task_config)`, `time_limit` from env.kwargs) under
    RecordEpisodeMetrics(AutoResetWrapper(AgentIDWrapper(RwareWrapper(.)))) (mava/utils/make_env.py:90-135, wrappers/jumanji.py:137-168)
    as device arrays. The dynamics are restated from the un-vendored dependency; see oracle/rware.py."""

    kind = L.ENV_RWARE
    reset_fn, step_fn = "magpo_rware_reset", "magpo_rware_step"

    def __init__(self, column_height: int = 8, shelf_rows: int = 1, shelf_columns: int = 3, num_agents: int = 4,
                 sensor_range: int = 1, request_queue_size: int = 4, time_limit: int = 500):
        self.num_agents, self.sensor_range, self.request_queue_size, self.time_limit = num_agents, sensor_range, request_queue_size, time_limit
        self.grid_size = ((column_height + 1) * shelf_rows + 2, 3 * shelf_columns + 1)
        self.cfg = L.RwareCfg(column_height, shelf_rows, shelf_columns, num_agents, sensor_range, request_queue_size, time_limit)
        self.num_shelves = int(L.lib().magpo_rware_num_shelves(C.byref(self.cfg)))
        L.check(min(self.num_shelves, 0), "magpo_rware_num_shelves")

    @property
    def obs_dim(self):
        return self.num_agents + 8 + 7 * (2 * self.sensor_range + 1) ** 2

    @property
    def action_dim(self):
        return 5

    def alloc_state(self, B: int, dev) -> dict:
        i32, f32, u8 = torch.int32, torch.float32, torch.uint8
        A, S, Q = self.num_agents, self.num_shelves, self.request_queue_size
        H, W = self.grid_size
        z = lambda *s, dt=i32: torch.zeros(*s, dtype=dt, device=dev)
        return dict(grid=z(B, 2, H, W), agent_pos=z(B, A, 2), agent_dir=z(B, A), agent_carry=z(B, A, dt=u8), shelf_pos=z(B, S, 2),
                    shelf_req=z(B, S, dt=u8), request_queue=z(B, Q), step_count=z(B), action_mask=z(B, A, 5, dt=u8), key=z(B, 2),
                    metrics_key=z(B, 2), running_return=z(B, dt=f32), running_length=z(B), episode_return=z(B, dt=f32),
                    episode_length=z(B))

    def state_struct(self, st: dict) -> L.RwareState:
        return L.struct_of(L.RwareState, **st)


def alloc_timestep(B, A, d, a, dev) -> dict:
    f32, i32, u8 = torch.float32, torch.int32, torch.uint8
    return dict(step_type=torch.zeros(B, dtype=torch.int8, device=dev), reward=torch.zeros(B, A, dtype=f32, device=dev),
                discount=torch.zeros(B, A, dtype=f32, device=dev), agents_view=torch.zeros(B, A, d, dtype=f32, device=dev),
                action_mask=torch.zeros(B, A, a, dtype=u8, device=dev), step_count=torch.zeros(B, A, dtype=i32, device=dev),
                next_agents_view=torch.zeros(B, A, d, dtype=f32, device=dev),
                next_step_count=torch.zeros(B, A, dtype=i32, device=dev),
                episode_return=torch.zeros(B, dtype=f32, device=dev), episode_length=torch.zeros(B, dtype=i32, device=dev),
                is_terminal_step=torch.zeros(B, dtype=u8, device=dev))


class MagpoLearner:
    """All device buffers of one rank + `update_step()` (rollout, GAE, P epochs x M minibatches)."""

    def __init__(self, env: CoordSumVec, sys: SystemConfig, device="cuda:0", allreduce=None, world_size: int = 1,
                 graph_rollout: bool = True, net: "NetworkConfig | None" = None):
        self.env, self.sys, self.dev = env, sys, torch.device(device)
        self.ctx = L.create_context(self.dev)  # this learner's MagpoContext: side streams + tensor-core weight images
        if sys.decay_learning_rates and sys.num_updates < 1:
            raise ValueError("decay_learning_rates needs system.num_updates >= 1 (utils/training.py:38-44)")
        # The T-step rollout is ~90 short launches per env step: after the first (eager) call it is replayed from one
        # CUDA graph, the way XLA runs command-buffer-compatible FFI handlers. All its operands are device-resident.
        self.graph_rollout, self._rollout_graph, self._rollout_graph_launches = graph_rollout, None, 0
        self.net = net or NetworkConfig(env.num_agents, env.obs_dim, env.action_dim, env.time_limit)
        if (self.net.n_agents, self.net.obs_dim, self.net.action_dim) != (env.num_agents, env.obs_dim, env.action_dim):
            raise ValueError("network configuration does not match the environment's (agents, obs_dim, action_dim)")
        # data-parallel exchange: `comm.NcclComm.attach(self)` (the library reduces inside magpo_minibatch_grads), or a host callable
        # `allreduce(flat)` summing the flat gradient buffer in place (CPU tests with gloo)
        self.allreduce, self.world_size, self.comm = allreduce, world_size, None
        self.c_net, self.c_sys = self.net.c_struct(), sys.c_struct()
        dev, f32, i32, u8 = self.dev, torch.float32, torch.int32, torch.uint8
        A, d, a = self.net.n_agents, self.net.obs_dim, self.net.action_dim
        T, E, U, M = sys.rollout_length, sys.num_envs, sys.update_batch_size, sys.num_minibatches
        if E % M:
            raise ValueError("num_envs must be divisible by num_minibatches")
        B = self.B = U * E
        self.g_table, self.n_g = param_table(self.net, 0)
        self.a_table, self.n_a = param_table(self.net, 1)
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.guider, self.actor = z(self.n_g), z(self.n_a)
        self.g_mu, self.g_nu, self.a_mu, self.a_nu = z(self.n_g), z(self.n_g), z(self.n_a), z(self.n_a)
        self.g_count, self.a_count = z(1, dt=i32), z(1, dt=i32)
        self.grads = z(self.n_g + self.n_a + 8)
        self.key = z(2, dt=i32)
        self.env_state = env.alloc_state(B, dev)
        self.ts = alloc_timestep(B, A, d, a, dev)
        ss = self.net.state_shape
        self.hs = dict(encoder=z(B, *ss), decoder_self=z(B, *ss), decoder_cross=z(B, *ss))
        self.policy_h = z(B, A, 128)
        self.traj = dict(done=z(T + 1, B, dt=u8), agents_view=z(T + 1, B, A, d), action_mask=z(T + 1, B, A, a, dt=u8),
                         step_count=z(T + 1, B, A, dt=i32), action=z(T, B, A, dt=i32), value=z(T, B, A), reward=z(T, B, A),
                         log_prob=z(T, B, A), policy_h0=z(B, A, 128),
                         sable_h0=dict(encoder=z(B, *ss), decoder_self=z(B, *ss), decoder_cross=z(B, *ss)),
                         episode_return=z(T, B), episode_length=z(T, B, dt=i32), is_terminal_step=z(T, B, dt=u8),
                         last_value=z(B, A))
        self.adv, self.targets = z(T, B, A), z(T, B, A)
        # update-side buffers
        Nmb = U * (E // M)
        self.chunk = Nmb if sys.chunk_envs <= 0 else min(sys.chunk_envs, Nmb)
        n = self.chunk
        self.mb = dict(agents_view=z(T, n, A, d), action_mask=z(T, n, A, a, dt=u8), step_count=z(T, n, A, dt=i32),
                       done=z(T, n, dt=u8), action=z(T, n, A, dt=i32), value=z(T, n, A), log_prob=z(T, n, A),
                       advantages=z(T, n, A), targets=z(T, n, A), policy_h0=z(n, A, 128),
                       sable_h0=dict(encoder=z(n, *ss), decoder_self=z(n, *ss), decoder_cross=z(n, *ss)))
        self.hs_perm, self.batch_perm, self.agent_perm = z(E, dt=i32), z(E, dt=i32), z(A, dt=i32)
        self.env_index, self.hs_index, self.env_slot = z(M * Nmb, dt=i32), z(M * Nmb, dt=i32), z(M * Nmb, dt=i32)
        self.idx_scratch = z(4 + 2 * max(E, A) + E + 16, dt=i32)
        self.stats, self.stats_scratch = z(U, 2), z(2 * U, dt=torch.float64)
        self.adam_scratch = z(1024)
        lib = L.lib()
        lib.magpo_rollout_workspace_bytes.restype = C.c_size_t
        lib.magpo_update_workspace_bytes.restype = C.c_size_t
        wr = lib.magpo_rollout_workspace_bytes(C.byref(self.c_net), B, T)
        wu = lib.magpo_update_workspace_bytes(C.byref(self.c_net), T, n)
        if int(wr) == 0 or int(wu) == 0:  # the size queries return 0 for shapes the kernels do not cover
            raise L.MagpoError(f"unsupported network / batch configuration: {self.net}")
        self.ws_bytes = max(int(wr), int(wu))
        self.workspace = torch.empty(self.ws_bytes, dtype=u8, device=dev)
        self.first_rollout = True
        self.loss_log: list[torch.Tensor] = []

    def __del__(self):
        try:
            torch.cuda.synchronize(self.dev)
            L.destroy_context(self.ctx)
        except Exception:
            pass

    # ------------------------------------------------------------------ structs
    def _ts_struct(self):
        return L.struct_of(L.TimeStep, **self.ts)

    def _traj_struct(self):
        t = dict(self.traj)
        t["sable_h0"] = L.struct_of(L.SableHState, **self.traj["sable_h0"])
        return L.struct_of(L.Trajectory, **t)

    def _mb_struct(self, n):
        m = {k: v for k, v in self.mb.items() if k != "sable_h0"}
        s = L.struct_of(L.Minibatch, **m)
        s.sable_h0 = L.struct_of(L.SableHState, **self.mb["sable_h0"])
        s.T, s.N = self.sys.rollout_length, n
        return s

    # ------------------------------------------------------------------ setup
    def set_params(self, guider: dict, actor: dict) -> None:
        load_params(self.guider, self.g_table, guider)
        load_params(self.actor, self.a_table, actor)

    def set_opt_state(self, guider_opt: dict, actor_opt: dict) -> None:
        """Load both Adam states: dicts with `count`, `mu`, `nu` (the moments keyed like the parameters)."""
        for opt, table, mu, nu, cnt in ((guider_opt, self.g_table, self.g_mu, self.g_nu, self.g_count),
                                        (actor_opt, self.a_table, self.a_mu, self.a_nu, self.a_count)):
            load_params(mu, table, opt["mu"])
            load_params(nu, table, opt["nu"])
            cnt.fill_(int(opt["count"]))

    def get_params(self):
        return ({k: v.clone() for k, v in param_views(self.guider, self.g_table).items()},
                {k: v.clone() for k, v in param_views(self.actor, self.a_table).items()})

    def reset(self, env_keys, step_key) -> None:
        """vmap(env.reset)(env_keys) for this rank's U*E envs + the shared step key (rec_magpo.py:642-673)."""
        B = self.B
        keys = torch.as_tensor(np.asarray(env_keys, dtype=np.uint32).view(np.int32).reshape(B, 2)).to(self.dev)
        self.key.copy_(torch.as_tensor(np.asarray(step_key, dtype=np.uint32).view(np.int32)))
        ts = dict(self.ts)
        ts.update(agents_view=self.traj["agents_view"][0], action_mask=self.traj["action_mask"][0],
                  step_count=self.traj["step_count"][0])
        L.call(self.env.reset_fn, L.stream_ptr(), C.byref(self.env.cfg), B, L.ptr(keys),
               self.env.state_struct(self.env_state), L.struct_of(L.TimeStep, **ts))
        self.traj["done"][0].zero_()
        for h in self.hs.values():
            h.zero_()
        self.policy_h.zero_()
        self.first_rollout = True
        torch.cuda.current_stream().synchronize()  # `keys` must outlive the launch

    # ------------------------------------------------------------------ the step
    def _rollout_call(self, carry_over: int) -> None:
        L.call("magpo_rollout", self.ctx, L.stream_ptr(), C.byref(self.c_net), C.byref(self.c_sys), self.env.kind, C.byref(self.env.cfg),
               C.byref(self.env.state_struct(self.env_state)), self._ts_struct(), L.ptr(self.guider), L.ptr(self.actor),
               L.ptr(self.key), L.struct_of(L.SableHState, **self.hs), L.ptr(self.policy_h), self._traj_struct(),
               carry_over, L.ptr(self.workspace), C.c_size_t(self.ws_bytes))

    def rollout(self) -> None:
        if self.first_rollout or not self.graph_rollout:
            self._rollout_call(0 if self.first_rollout else 1)
            self.first_rollout = False
            return
        lib = L.lib()
        if self._rollout_graph is None:
            lib.magpo_launch_count.restype = C.c_int64
            n0 = lib.magpo_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._rollout_call(1)
            self._rollout_graph_launches = int(lib.magpo_launch_count() - n0)
            lib.magpo_launch_count_add(C.c_int64(-self._rollout_graph_launches))  # capturing launched nothing
            self._rollout_graph = g
        self._rollout_graph.replay()
        lib.magpo_launch_count_add(C.c_int64(self._rollout_graph_launches))

    def gae(self) -> None:
        T, B, A = self.sys.rollout_length, self.B, self.net.n_agents
        L.call("magpo_gae", L.stream_ptr(), T, B, A, L.ptr(self.traj["reward"]), L.ptr(self.traj["value"]),
               L.ptr(self.traj["done"]), L.ptr(self.traj["last_value"]), L.ptr(self.traj["done"][T]),
               C.c_double(self.sys.gamma), C.c_double(self.sys.gae_lambda), L.ptr(self.adv), L.ptr(self.targets))

    def minibatch_grads(self, m: int) -> None:
        """Zeroes self.grads and accumulates the slot-averaged gradients + loss sums of minibatch m."""
        sysc, s = self.sys, L.stream_ptr()
        T, A = sysc.rollout_length, self.net.n_agents
        Nmb = sysc.update_batch_size * (sysc.num_envs // sysc.num_minibatches)
        env_index = self.env_index[m * Nmb:(m + 1) * Nmb]
        hs_index = self.hs_index[m * Nmb:(m + 1) * Nmb]
        env_slot = self.env_slot[m * Nmb:(m + 1) * Nmb]
        self.grads.zero_()
        L.call("magpo_adv_stats", s, T, self.B, A, L.ptr(self.adv), L.ptr(env_index), Nmb, sysc.update_batch_size,
               L.ptr(self.stats_scratch), L.ptr(self.stats))
        inv_tokens = 1.0 / (Nmb * T * A)
        for c0 in range(0, Nmb, self.chunk):
            n = min(self.chunk, Nmb - c0)
            mbs = self._mb_struct(n)
            L.call("magpo_pack_minibatch", s, C.byref(self.c_net), C.byref(self.c_sys), self._traj_struct(),
                   L.ptr(self.adv), L.ptr(self.targets), L.ptr(env_index[c0:c0 + n]), L.ptr(hs_index[c0:c0 + n]),
                   L.ptr(self.agent_perm), n, mbs)
            last = c0 + n >= Nmb  # the attached communicator reduces once the whole minibatch has been accumulated
            L.call("magpo_minibatch_grads", self.ctx, s, C.byref(self.c_net), C.byref(self.c_sys), L.ptr(self.guider),
                   L.ptr(self.actor), mbs, L.ptr(env_slot[c0:c0 + n]), L.ptr(self.stats), C.c_float(inv_tokens),
                   L.ptr(self.grads), 1 if (last and self.comm is not None) else 0, L.ptr(self.workspace), C.c_size_t(self.ws_bytes))

    def apply_grads(self) -> None:
        sysc, s = self.sys, L.stream_ptr()
        if self.allreduce is not None and self.comm is None:
            self.allreduce(self.grads)  # sum over ranks; the mean is taken by grad_scale
        scale = 1.0 / self.world_size
        g, a = self.grads[:self.n_g], self.grads[self.n_g:self.n_g + self.n_a]
        # both optimisers run on actor_lr (rec_magpo.py:581-589); the schedule is evaluated on the device from the count
        period = sysc.ppo_epochs * sysc.num_minibatches if sysc.decay_learning_rates else 0
        L.call("magpo_clip_adam_sched", s, C.c_int64(self.n_g), L.ptr(self.guider), L.ptr(g), L.ptr(self.g_mu), L.ptr(self.g_nu),
               L.ptr(self.g_count), C.c_float(scale), C.c_float(sysc.actor_lr), period, int(sysc.num_updates),
               C.c_float(sysc.max_grad_norm), L.ptr(self.adam_scratch))
        L.call("magpo_clip_adam_sched", s, C.c_int64(self.n_a), L.ptr(self.actor), L.ptr(a), L.ptr(self.a_mu), L.ptr(self.a_nu),
               L.ptr(self.a_count), C.c_float(scale), C.c_float(sysc.actor_lr), period, int(sysc.num_updates),
               C.c_float(sysc.max_grad_norm), L.ptr(self.adam_scratch[512:]))

    def epoch_indices(self, first: bool) -> None:
        L.call("magpo_epoch_indices", L.stream_ptr(), C.byref(self.c_sys), self.net.n_agents, L.ptr(self.key),
               L.ptr(self.hs_perm), 1 if first else 0, L.ptr(self.batch_perm), L.ptr(self.agent_perm),
               L.ptr(self.env_index), L.ptr(self.hs_index), L.ptr(self.env_slot), L.ptr(self.idx_scratch))

    def update_step(self):
        """One `_update_step`. Returns (episode metrics dict of [T,B] tensors (views), loss tensor [P,M,8])."""
        sysc = self.sys
        self.rollout()
        self.gae()
        losses = torch.empty(sysc.ppo_epochs, sysc.num_minibatches, 8, device=self.dev)
        for p in range(sysc.ppo_epochs):
            self.epoch_indices(p == 0)
            for m in range(sysc.num_minibatches):
                self.minibatch_grads(m)
                self.apply_grads()
                losses[p, m] = self.grads[self.n_g + self.n_a:] * (1.0 / self.world_size)
        metrics = dict(episode_return=self.traj["episode_return"], episode_length=self.traj["episode_length"],
                       is_terminal_step=self.traj["is_terminal_step"])
        return metrics, losses

    @staticmethod
    def loss_info(losses: torch.Tensor, sys: SystemConfig) -> dict:
        """The six train metrics of rec_magpo.py:427-434 from the [.., 8] loss sums."""
        gl, ent, vl, kl, al, kla = (losses[..., i] for i in (1, 2, 3, 4, 6, 7))
        total_g = gl + kl - sys.ent_coef * ent + sys.vf_coef * vl
        total_a = al * sys.alpha + kla
        return dict(total_loss=total_g + total_a, value_loss=vl, actor_loss=al, guider_loss=gl, kl_loss=kl, entropy=ent)

    def sable_hidden_state(self) -> dict:
        """LearnerState.hstates.sable_hidden_state: reset where the last step was terminal (rec_magpo.py:165-169)."""
        done = self.traj["done"][self.sys.rollout_length].bool().view(-1, 1, 1, 1, 1)
        return {k: torch.where(done, torch.zeros_like(v), v) for k, v in self.hs.items()}
