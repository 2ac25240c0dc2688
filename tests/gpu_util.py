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


# ----------------------------------------------------------------------------- BASELINE configs at real T + fp32/fp64 control
def elementwise_dev(new: dict, ref: dict, floor: float = 1e-3) -> dict:
    """max over elements of |new - ref| / max(|ref|, floor), per tensor (the element-wise form of north_star's
    'updated parameters within rtol 1e-4 per update'; `floor` keeps exactly-zero parameters comparable)."""
    out = {}
    for k, r in ref.items():
        r = np.asarray(r, np.float64)
        n = np.asarray(new[k].cpu().numpy() if torch.is_tensor(new[k]) else new[k], np.float64)
        out[k] = float((np.abs(n - r) / np.maximum(np.abs(r), floor)).max())
    return out


def update_dev(new: dict, ref: dict, lr: float, rtol: float = 1e-4, atol_lr: float = 0.1) -> dict:
    """Per tensor: (max |new - ref| in units of the learning rate, max |new - ref| / (rtol |ref| + atol_lr lr)). The second is the
    allclose-style violation of `|dp| <= rtol |p| + atol_lr * lr` (<= 1 passes): north_star's rtol 1e-4 on the parameter plus a
    tenth of ONE optimiser step's reach — an update consists of P*M Adam steps of at most ~lr each, and a parameter that starts at
    zero (biases, SwiGLU) has no scale of its own for a relative bound."""
    out = {}
    for k, r in ref.items():
        r = np.asarray(r, np.float64)
        n = np.asarray(new[k].cpu().numpy() if torch.is_tensor(new[k]) else new[k], np.float64)
        d = np.abs(n - r)
        out[k] = (float(d.max()) / lr, float((d / (rtol * np.abs(r) + atol_lr * lr)).max()))
    return out


def baseline_case(env: str, dev, E=16, U=2, T=128, P=4, M=2, seed=42):
    """(spec, ncfg, osys, oracle state, CUDA learner) of a BASELINE.json config at its real rollout length:
    'coordsum' = configs[0] exactly (CoordSum 3x10-30, num_envs=16, rollout_length=128, U=2, P=4, M=2); 'lbf' = configs[1]'s
    env and 'rware' = configs[2]'s (tiny-4ag) at the same small num_envs so that the CPU oracle finishes in seconds."""
    from magpo_b200.learner import CoordSumVec, LbfVec, MagpoLearner, RwareVec, SystemConfig
    from oracle import coordsum as ocs, lbf as olbf, learner as olr, nets as onets, prng as oprng, rware as orw

    if env == "coordsum":
        kw = ocs.SCENARIOS["3x10-30-v0"]
        spec, vec = ocs.CoordSumSpec(**kw), CoordSumVec(**kw)
    elif env == "lbf":
        kw = olbf.SCENARIOS["2s-8x8-2p-2f-coop"]
        spec, vec = olbf.LbfSpec(**kw), LbfVec(**kw)
    else:
        kw = orw.SCENARIOS["tiny-4ag"]
        spec, vec = orw.RwareSpec(**kw), RwareVec(**kw)
    ncfg = onets.NetCfg(spec.num_agents, spec.obs_dim, spec.action_dim)
    osys = olr.SysCfg(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M)
    state = olr.learner_setup(spec, ncfg, osys, seed=seed)
    lrn = MagpoLearner(vec, SystemConfig(num_envs=E, update_batch_size=U, rollout_length=T, ppo_epochs=P, num_minibatches=M),
                       device=dev)
    lrn.set_params(state["guider_params"], state["actor_params"])
    ks = oprng.split(oprng.prng_key(seed), 4)
    allk = oprng.split(ks[0], U * E + 1)
    lrn.reset(allk[1:], oprng.split(allk[0])[1])
    return spec, ncfg, osys, state, lrn


def run_baseline_updates(env: str, dev, updates: int = 2, with_fp64: bool = True, resync: bool = True, **kw):
    """`updates` consecutive `_update_step`s of the CUDA path and of the fp32 oracle from the same state; per update the
    trajectory comparison (exact), the losses, and the element-wise parameter deviations
        cuda_vs_o32, and with `with_fp64`: cuda_vs_o64 next to o32_vs_o64 (the control: what fp32 arithmetic alone costs
        against a double-precision update from the identical state).
    `resync`: after each update the CUDA learner's parameters and Adam state are reloaded from the fp32 oracle, so that every
    update is measured from an identical start (env / hidden states are never touched: they stay bit-identical on their own)."""
    import copy

    from magpo_b200.learner import MagpoLearner
    from oracle import learner as olr

    spec, ncfg, osys, state, lrn = baseline_case(env, dev, **kw)
    U, E, T = osys.update_batch_size, osys.num_envs, osys.rollout_length
    rows = []
    for upd in range(updates):
        s64 = copy.deepcopy(state) if with_fp64 else None
        rec = {}
        _, infos = olr.update_step(state, spec, ncfg, osys, record=rec)
        if with_fp64:
            olr.update_step(s64, spec, ncfg, osys, dtype=torch.float64)
        _, losses = lrn.update_step()
        sync()
        row = dict(update=upd, actions_exact=True, rewards_exact=True, obs_exact=True)
        for u in range(U):
            sl = slice(u * E, (u + 1) * E)
            row["actions_exact"] &= bool((lrn.traj["action"].cpu().numpy()[:, sl] == rec["traj"][u]["action"]).all())
            row["rewards_exact"] &= bool((lrn.traj["reward"].cpu().numpy()[:, sl] == rec["traj"][u]["reward"]).all())
            row["obs_exact"] &= bool((lrn.traj["agents_view"].cpu().numpy()[:T, sl] == rec["traj"][u]["obs"].astype(np.float32)).all())
            row["value_rel"] = max(row.get("value_rel", 0.0), rel_err(lrn.traj["value"].cpu().numpy()[:, sl], rec["traj"][u]["value"]))
            row["logp_rel"] = max(row.get("logp_rel", 0.0), rel_err(lrn.traj["log_prob"].cpu().numpy()[:, sl], rec["traj"][u]["log_prob"]))
        li = MagpoLearner.loss_info(losses.cpu(), lrn.sys)
        worst, k = 0.0, 0
        for p in range(osys.ppo_epochs):
            for m in range(osys.num_minibatches):
                for name in ("value_loss", "actor_loss", "guider_loss", "kl_loss", "entropy", "total_loss"):
                    ref, got = infos[k][name], float(li[name][p, m])
                    worst = max(worst, abs(got - ref) / max(1.0, abs(ref)))
                k += 1
        row["loss_dev"] = worst
        gp, ap = lrn.get_params()
        cuda = {**{"g/" + k: v for k, v in gp.items()}, **{"a/" + k: v for k, v in ap.items()}}
        o32 = {**{"g/" + k: v for k, v in state["guider_params"].items()}, **{"a/" + k: v for k, v in state["actor_params"].items()}}
        d = elementwise_dev(cuda, o32)
        row["cuda_vs_o32"] = max(d.values())
        row["cuda_vs_o32_tensor"] = max(d, key=d.get)
        # max-norm over the tensors that have a scale of their own (zero-initialised ones are a few lr large after one update)
        row["cuda_vs_o32_maxnorm"] = max(rel_err(cuda[k].cpu().numpy(), o32[k]) for k in o32 if np.abs(o32[k]).max() > 0.02)
        lr = osys.actor_lr
        u = update_dev(cuda, o32, lr)
        row["cuda_vs_o32_lr"] = max(v[0] for v in u.values())
        row["cuda_vs_o32_viol"] = max(v[1] for v in u.values())
        if with_fp64:
            o64 = {**{"g/" + k: v for k, v in s64["guider_params"].items()}, **{"a/" + k: v for k, v in s64["actor_params"].items()}}
            dc, do = elementwise_dev(cuda, o64), elementwise_dev(o32, o64)
            row["cuda_vs_o64"], row["o32_vs_o64"] = max(dc.values()), max(do.values())
            row["cuda_vs_o64_tensor"], row["o32_vs_o64_tensor"] = max(dc, key=dc.get), max(do, key=do.get)
            uc, uo = update_dev(cuda, o64, lr), update_dev(o32, o64, lr)
            row["cuda_vs_o64_lr"], row["o32_vs_o64_lr"] = max(v[0] for v in uc.values()), max(v[0] for v in uo.values())
            row["cuda_vs_o64_viol"], row["o32_vs_o64_viol"] = max(v[1] for v in uc.values()), max(v[1] for v in uo.values())
        rows.append(row)
        if resync:
            lrn.set_params(state["guider_params"], state["actor_params"])
            lrn.set_opt_state(state["guider_opt"], state["actor_opt"])
    return rows
