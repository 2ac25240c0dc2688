"""The rec_magpo Anakin learner (rollout + GAE + update) restated on CPU — oracle only.

Follows `mava/systems/gpo/anakin/rec_magpo.py`:
  * `_env_step`            :126-187      -> `rollout`
  * bootstrap value        :202-208
  * `calculate_gae`        `mava/utils/multistep.py:24-68` -> `gae`
  * `_update_epoch` shuffle:437-462      -> `make_minibatches`
  * `_guider_loss_fn`      :222-311,  `_actor_loss_fn` :313-370
  * pmean over "batch"     :395-409      -> mean over the U update-batch slots
  * optimisers             :581-589,412-423 (optax 0.2.4 clip_by_global_norm + adam, Appendix A11)
Network math runs in torch on the CPU (autograd mirrors jax.value_and_grad); integer work in NumPy.
One "device" only; multi-device = the same thing per shard with the gradient mean over shards.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import coordsum, nets, prng


def env_module(spec):
    """The env restatement matching `spec` (make_env.py:202-218 picks the maker by env name)."""
    if isinstance(spec, coordsum.CoordSumSpec):
        return coordsum
    from . import lbf

    if isinstance(spec, lbf.LbfSpec):
        return lbf
    from . import rware

    if isinstance(spec, rware.RwareSpec):
        return rware
    raise TypeError(f"no env restatement for {type(spec).__name__}")


@dataclass
class SysCfg:
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
    sable_only: bool = False  # rec_sable (mava/systems/sable/anakin/rec_sable.py): Sable alone under PPO, no learner
    decay_learning_rates: bool = False  # mava/utils/training.py:48-64
    num_updates: int = 0


def make_learning_rate(sys: "SysCfg", f=np.float32):
    """mava/utils/training.py:20-64 (rec_magpo.py:581): constant `actor_lr`, or the linear schedule
    count -> init_lr * (1 - (count // (ppo_epochs * num_minibatches)) / num_updates) in float32 like the traced jnp code."""
    if not sys.decay_learning_rates:
        return sys.actor_lr
    period = sys.ppo_epochs * sys.num_minibatches

    def linear_schedule(count: int):
        frac = f(1.0) - f(count // period) / f(sys.num_updates)
        return f(sys.actor_lr) * frac

    return linear_schedule


# ----------------------------------------------------------------------------- GAE
def gae(done, value, reward, last_val, last_done, gamma, lam):
    """calculate_gae (multistep.py:51-68). done bool[T,...], value/reward f32[T,...]. fp32 op order kept."""
    T = value.shape[0]
    g = np.float32(gamma)
    gl = np.float32(gamma * lam)  # Python-float product, rounded once (weak-typed constant)
    adv = np.zeros_like(value, dtype=np.float32)
    acc = np.zeros_like(last_val, dtype=np.float32)
    nv = last_val.astype(np.float32)
    nd = last_done.astype(np.float32)
    one = np.float32(1.0)
    for t in range(T - 1, -1, -1):
        delta = reward[t] + g * nv * (one - nd) - value[t]
        acc = (delta + gl * (one - nd) * acc).astype(np.float32)
        adv[t] = acc
        nv, nd = value[t], done[t].astype(np.float32)
    return adv, (adv + value).astype(np.float32)


# ----------------------------------------------------------------------------- optimiser
def clip_adam_step(params, grads, opt, lr, max_norm, b1=0.9, b2=0.999, eps=1e-5, f=np.float32):
    """optax.chain(clip_by_global_norm, adam(lr, eps=1e-5)) + apply_updates (Appendix A11). In place, fp32
    (`f=np.float64`: the same arithmetic in double, for the fp32-vs-fp64 tolerance control only). `lr` may be a
    callable count -> learning rate (optax schedule, evaluated at the count BEFORE the increment)."""
    sq = f(0.0)
    for k in grads:
        sq += f(np.sum(grads[k].astype(f) ** 2, dtype=f))
    gnorm = f(np.sqrt(sq))
    step_lr = lr(int(opt["count"])) if callable(lr) else lr
    opt["count"] = int(opt["count"]) + 1
    c = opt["count"]
    bc1 = f(1) - f(np.power(f(b1), c, dtype=f))
    bc2 = f(1) - f(np.power(f(b2), c, dtype=f))
    for k in params:
        g = grads[k].astype(f)
        if not (gnorm < f(max_norm)):
            g = (g / gnorm) * f(max_norm)
        mu = opt["mu"][k] = (f(b1) * opt["mu"][k] + f(1 - b1) * g).astype(f)
        nu = opt["nu"][k] = (f(b2) * opt["nu"][k] + f(1 - b2) * g * g).astype(f)
        u = (mu / bc1) / (np.sqrt(nu / bc2) + f(eps))
        params[k] = (params[k] + u * f(-step_lr)).astype(f)
    return gnorm


def init_opt(params):
    return dict(count=0, mu={k: np.zeros_like(v) for k, v in params.items()}, nu={k: np.zeros_like(v) for k, v in params.items()})


# ----------------------------------------------------------------------------- losses
def _kl_cat(logp1, logp2):
    """distrax _kl_divergence_categorical_categorical (Appendix A10)."""
    p1 = torch.exp(logp1)
    return (torch.where(p1 == 0, torch.zeros_like(p1), logp1 - logp2) * p1).sum(-1)


def _norm_adv(adv):
    return (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)


def guider_loss(g_logits, value, g_ent, a_logits, mb, sys: SysCfg):
    """_guider_loss_fn body after the network calls (rec_magpo.py:251-311). a_logits is detached."""
    act = mb["action"].to(torch.int64)[..., None]
    lg_all = nets.log_softmax(g_logits)
    la_all = nets.log_softmax(a_logits.detach())
    kl = _kl_cat(lg_all, la_all)
    lg = lg_all.gather(-1, act)[..., 0]
    la = la_all.gather(-1, act)[..., 0]
    lo, hi = math.log(1 / sys.clip_gpo), math.log(sys.clip_gpo)
    diff = lg - la
    ratio = torch.exp(lg - mb["log_prob"])
    clipped_ratio = torch.exp(torch.clamp(diff, lo, hi) + la - mb["log_prob"])
    mask = ((diff < lo) | (diff > hi)).to(value.dtype)
    kl_loss = (kl * mask).mean()
    adv = _norm_adv(mb["adv"])
    l1 = ratio * adv
    l2 = torch.clamp(clipped_ratio, 1.0 - sys.clip_eps, 1.0 + sys.clip_eps) * adv
    g_loss = -torch.minimum(l1, l2).mean()
    ent = g_ent.mean()
    vclip = mb["value"] + (value - mb["value"]).clamp(-sys.clip_eps, sys.clip_eps)
    v_loss = 0.5 * torch.maximum((value - mb["targets"]) ** 2, (vclip - mb["targets"]) ** 2).mean()
    total = g_loss + kl_loss - sys.ent_coef * ent + sys.vf_coef * v_loss
    return total, (g_loss, ent, v_loss, kl_loss)


def sable_ppo_loss(g_logits, value, g_ent, mb, sys: SysCfg):
    """rec_sable.py:196-226 (`_loss_fn` after the network call): clipped PPO + clipped value loss - entropy bonus."""
    act = mb["action"].to(torch.int64)[..., None]
    log_prob = nets.log_softmax(g_logits).gather(-1, act)[..., 0]
    ratio = torch.exp(log_prob - mb["log_prob"])
    adv = _norm_adv(mb["adv"])
    a_loss = -torch.minimum(ratio * adv, torch.clamp(ratio, 1.0 - sys.clip_eps, 1.0 + sys.clip_eps) * adv).mean()
    ent = g_ent.mean()
    vclip = mb["value"] + (value - mb["value"]).clamp(-sys.clip_eps, sys.clip_eps)
    v_loss = 0.5 * torch.maximum((value - mb["targets"]) ** 2, (vclip - mb["targets"]) ** 2).mean()
    total = a_loss - sys.ent_coef * ent + sys.vf_coef * v_loss
    return total, (a_loss, ent, v_loss)


def actor_loss(a_logits, g_logits, mb, sys: SysCfg):
    """_actor_loss_fn body after the network calls (rec_magpo.py:338-370). g_logits is detached."""
    act = mb["action"].to(torch.int64)[..., None]
    la_all = nets.log_softmax(a_logits)
    lg_all = nets.log_softmax(g_logits.detach())
    la = la_all.gather(-1, act)[..., 0]
    kl_loss = _kl_cat(lg_all, la_all).mean()
    ratio = torch.exp(la - mb["log_prob"])
    adv = _norm_adv(mb["adv"])
    l1 = ratio * adv
    l2 = torch.clamp(ratio, 1.0 - sys.clip_eps, 1.0 + sys.clip_eps) * adv
    a_loss = -torch.minimum(l1, l2).mean()
    total = a_loss * sys.alpha + kl_loss
    return total, (a_loss, kl_loss)


def forward_reshape(x, A):  # rec_magpo.py:60-75  (N, T*A, ...) -> (T, N, A, ...)
    n, ta = x.shape[:2]
    x = x.reshape(n, ta // A, A, *x.shape[2:])
    return x.transpose(0, 1) if isinstance(x, torch.Tensor) else np.swapaxes(x, 0, 1)


def backward_reshape(x):  # rec_magpo.py:78-88  (T, N, A, ...) -> (N, T*A, ...)
    x = x.transpose(0, 1)
    return x.reshape(x.shape[0], x.shape[1] * x.shape[2], *x.shape[3:])


def minibatch_losses_and_grads(gp_np, ap_np, mb_np, ncfg: nets.NetCfg, sys: SysCfg, dtype=torch.float32):
    """One slot's `guider_grad_fn` + `actor_grad_fn` (rec_magpo.py:374-391). mb_np: dict of numpy arrays
    [N, C, ...] (+ prev_hstates [N,...], policy_h0 [N,A,H]). Returns (guider grads, actor grads, info dict)."""
    T = sys.rollout_length
    A = ncfg.n_agents
    gp = nets.to_torch(gp_np, dtype, requires_grad=True)
    ap = nets.to_torch(ap_np, dtype, requires_grad=True)
    mb = {}
    for k, v in mb_np.items():
        if k == "prev_hstates":
            mb[k] = tuple(torch.tensor(x, dtype=dtype) for x in v)
        elif v.dtype.kind == "f":
            mb[k] = torch.tensor(v, dtype=dtype)
        else:
            mb[k] = torch.tensor(v)
    obs = mb["obs"].to(dtype)
    value, g_logp, g_ent, g_logits = nets.sable_apply(
        gp, ncfg, obs, mb["action_mask"], mb["step_count"], mb["action"], mb["prev_hstates"], mb["done"], T
    )
    if sys.sable_only:  # rec_sable.py:172-258: one network, one loss
        tot, (a_loss, ent, v_loss) = sable_ppo_loss(g_logits, value, g_ent, mb, sys)
        gg = torch.autograd.grad(tot, list(gp.values()), allow_unused=True)
        g_grads = {k: (np.zeros_like(gp_np[k]) if g is None else g.detach().to(torch.float32).numpy()) for k, g in zip(gp, gg)}
        a_grads = {k: np.zeros_like(v) for k, v in ap_np.items()}
        info = dict(total_loss=float(tot.detach()), value_loss=float(v_loss.detach()), actor_loss=float(a_loss.detach()), guider_loss=float(a_loss.detach()),
                    kl_loss=0.0, entropy=float(ent.detach()), total_guider=float(tot.detach()), total_actor=0.0, actor_kl=0.0)
        return g_grads, a_grads, info, dict(value=value.detach().numpy(), g_logits=g_logits.detach().numpy(), a_logits=None)
    _, a_logits_t = nets.actor_apply(
        ap, ncfg, mb["policy_h0"], forward_reshape(obs, A), forward_reshape(mb["done"], A), forward_reshape(mb["action_mask"], A)
    )
    a_logits = backward_reshape(a_logits_t)
    tot_g, (g_loss, ent, v_loss, kl_g) = guider_loss(g_logits, value, g_ent, a_logits, mb, sys)
    tot_a, (a_loss, kl_a) = actor_loss(a_logits, g_logits, mb, sys)
    gg = torch.autograd.grad(tot_g, list(gp.values()), allow_unused=True, retain_graph=True)
    ga = torch.autograd.grad(tot_a, list(ap.values()), allow_unused=True)
    g_grads = {k: (np.zeros_like(gp_np[k]) if g is None else g.detach().to(dtype).numpy()) for k, g in zip(gp, gg)}
    a_grads = {k: (np.zeros_like(ap_np[k]) if g is None else g.detach().to(dtype).numpy()) for k, g in zip(ap, ga)}
    info = dict(
        total_loss=float(tot_g.detach()) + float(tot_a.detach()), value_loss=float(v_loss.detach()), actor_loss=float(a_loss.detach()),
        guider_loss=float(g_loss.detach()), kl_loss=float(kl_g.detach()), entropy=float(ent.detach()),
        total_guider=float(tot_g.detach()), total_actor=float(tot_a.detach()), actor_kl=float(kl_a.detach()),
    )
    aux = dict(value=value.detach().numpy(), g_logits=g_logits.detach().numpy(), a_logits=a_logits.detach().numpy())
    return g_grads, a_grads, info, aux


# ----------------------------------------------------------------------------- rollout
def init_hstates(ncfg: nets.NetCfg, E):
    hs = ncfg.head_size
    shp = (E, ncfg.n_head, ncfg.n_block, hs, hs)
    return dict(sable=tuple(np.zeros(shp, np.float32) for _ in range(3)), policy=np.zeros((E, ncfg.n_agents, ncfg.hidden), np.float32))


def rollout(spec, ncfg: nets.NetCfg, sys: SysCfg, gp_np, ap_np, slot):
    """lax.scan(_env_step, length=T) for one update-batch slot (rec_magpo.py:126-197). Mutates `slot`
    (key, env_state, timestep, dones, hstates) and returns the trajectory dict [T,E,...] + metrics."""
    gp, ap = nets.to_torch(gp_np), nets.to_torch(ap_np)
    T, E, A = sys.rollout_length, sys.num_envs, ncfg.n_agents
    traj = {k: [] for k in ("done", "action", "value", "reward", "log_prob", "obs", "action_mask", "step_count")}
    metrics = {k: [] for k in ("episode_return", "episode_length", "is_terminal_step")}
    policy_h0 = slot["hstates"]["policy"].copy()
    with torch.no_grad():
        for _ in range(T):
            ks = prng.split(slot["key"])
            slot["key"], policy_key = ks[0], ks[1]
            ts = slot["timestep"]
            ob = ts["observation"]
            obs_f = torch.tensor(ob["agents_view"].astype(np.float32))
            hs_t = tuple(torch.tensor(h) for h in slot["hstates"]["sable"])
            action, logp, value, new_hs = nets.sable_get_actions(
                gp, ncfg, obs_f, torch.tensor(ob["action_mask"]), torch.tensor(ob["step_count"]), hs_t, policy_key
            )
            last_done = slot["dones"]
            if sys.sable_only:  # rec_sable.py:86-120 has no learner
                ph = torch.tensor(slot["hstates"]["policy"])
            else:
                ph, _ = nets.actor_apply(
                    ap, ncfg, torch.tensor(slot["hstates"]["policy"]), obs_f[None], torch.tensor(last_done)[None], torch.tensor(ob["action_mask"])[None]
                )
            prev_done = np.repeat((ts["step_type"] == coordsum.STEP_LAST)[:, None], A, axis=1)
            env_state, new_ts = env_module(spec).step(spec, slot["env_state"], action)
            done = new_ts["step_type"] == coordsum.STEP_LAST
            sable = tuple(np.where(done[:, None, None, None, None], 0.0, h.numpy()).astype(np.float32) for h in new_hs)
            traj["done"].append(prev_done)
            traj["action"].append(action)
            traj["value"].append(value.numpy().astype(np.float32))
            traj["reward"].append(new_ts["reward"])
            traj["log_prob"].append(logp.numpy().astype(np.float32))
            traj["obs"].append(ob["agents_view"])
            traj["action_mask"].append(ob["action_mask"])
            traj["step_count"].append(ob["step_count"])
            for k in metrics:
                metrics[k].append(new_ts["extras"]["episode_metrics"][k])
            slot["env_state"], slot["timestep"] = env_state, new_ts
            slot["dones"] = np.repeat(done[:, None], A, axis=1)
            slot["hstates"] = dict(sable=sable, policy=ph.numpy().astype(np.float32))
    traj = {k: np.stack(v) for k, v in traj.items()}
    traj["policy_h0"] = policy_h0
    return traj, {k: np.stack(v) for k, v in metrics.items()}


def bootstrap_value(ncfg, gp_np, slot):
    """rec_magpo.py:202-208: key split + full get_actions, keep the value."""
    gp = nets.to_torch(gp_np)
    ks = prng.split(slot["key"])
    slot["key"], last_val_key = ks[0], ks[1]
    ob = slot["timestep"]["observation"]
    with torch.no_grad():
        _, _, v, _ = nets.sable_get_actions(
            gp, ncfg, torch.tensor(ob["agents_view"].astype(np.float32)), torch.tensor(ob["action_mask"]),
            torch.tensor(ob["step_count"]), tuple(torch.tensor(h) for h in slot["hstates"]["sable"]), last_val_key,
        )
    return v.numpy().astype(np.float32)


def make_minibatches(traj, adv, targets, prev_hs, batch_perm, agent_perm, M):
    """The shuffle of `_update_epoch` (rec_magpo.py:441-462): take(axis=1) by env, take(axis=2) by agent,
    (T,E,A,..)->(E,T*A,..), split into M minibatches. Returns a list of M dicts."""
    def prep(x):
        x = np.take(np.take(x, batch_perm, axis=1), agent_perm, axis=2)
        x = np.moveaxis(x, 0, 1)
        return x.reshape(x.shape[0], x.shape[1] * x.shape[2], *x.shape[3:])

    full = {k: prep(traj[k]) for k in ("done", "action", "value", "reward", "log_prob", "obs", "action_mask", "step_count")}
    full["adv"], full["targets"] = prep(adv), prep(targets)
    E = batch_perm.shape[0]
    N = E // M
    h0 = np.take(np.take(traj["policy_h0"], batch_perm, axis=0), agent_perm, axis=1)
    phs = tuple(np.take(h, batch_perm, axis=0) for h in prev_hs)
    out = []
    for m in range(M):
        sl = slice(m * N, (m + 1) * N)
        mb = {k: v[sl] for k, v in full.items()}
        mb["policy_h0"] = h0[sl]
        mb["prev_hstates"] = tuple(h[sl] for h in phs)
        out.append(mb)
    return out


# ----------------------------------------------------------------------------- setup + learn
def learner_setup(spec, ncfg: nets.NetCfg, sys: SysCfg, seed: int = 42, n_devices: int = 1, device: int = 0, param_seed: int = 0):
    """learner_setup (rec_magpo.py:533-685) for one device shard: keys, env resets, replicated state."""
    ks = prng.split(prng.prng_key(seed), 4)  # run_experiment :699-701
    key = ks[0]
    U, E = sys.update_batch_size, sys.num_envs
    all_keys = prng.split(key, n_devices * U * E + 1)  # :642-644
    key, env_keys = all_keys[0], all_keys[1:].reshape(n_devices, U, E, 2)
    ks2 = prng.split(key)  # :660
    step_key = ks2[1]
    gp = nets.init_guider_params(ncfg, param_seed)
    ap = nets.init_actor_params(ncfg, param_seed + 1)
    slots = []
    for u in range(U):
        env_state, ts = env_module(spec).reset(spec, env_keys[device, u])
        slots.append(dict(key=step_key.copy(), env_state=env_state, timestep=ts,
                          dones=np.zeros((E, ncfg.n_agents), bool), hstates=init_hstates(ncfg, E)))
    return dict(guider_params=gp, actor_params=ap, guider_opt=init_opt(gp), actor_opt=init_opt(ap), slots=slots)


def update_step(state, spec, ncfg: nets.NetCfg, sys: SysCfg, grad_allreduce=None, record=None, dtype=torch.float32):
    """_update_step (rec_magpo.py:106-499) for all U slots of one device. `grad_allreduce(flat)->flat`
    stands in for the pmean over "device". `record` (dict) collects intermediates for parity tests.
    `dtype=torch.float64` runs the update's network math, gradient mean and Adam in double (the rollout stays fp32):
    the reference point of the fp32-vs-fp64 tolerance control (tools/tolerance_control.py), never a parity target."""
    f = np.float64 if dtype == torch.float64 else np.float32
    lr = make_learning_rate(sys, f)
    U, P, M = sys.update_batch_size, sys.ppo_epochs, sys.num_minibatches
    trajs, advs, tgts, prevs, mets = [], [], [], [], []
    for slot in state["slots"]:
        prev_hs = tuple(h.copy() for h in slot["hstates"]["sable"])
        traj, met = rollout(spec, ncfg, sys, state["guider_params"], state["actor_params"], slot)
        last_val = bootstrap_value(ncfg, state["guider_params"], slot)
        adv, tgt = gae(traj["done"], traj["value"], traj["reward"], last_val, slot["dones"], sys.gamma, sys.gae_lambda)
        trajs.append(traj); advs.append(adv); tgts.append(tgt); prevs.append(prev_hs); mets.append(met)
    if record is not None:
        record.update(traj=trajs, adv=advs, targets=tgts, minibatches=[], grads=[])
    loss_infos = []
    keys = [slot["key"] for slot in state["slots"]]
    for _ in range(P):
        mbs = []
        for u in range(U):
            k4 = prng.split(keys[u], 4)  # :439
            keys[u] = k4[0]
            batch_perm = prng.permutation(k4[1], sys.num_envs)
            agent_perm = prng.permutation(k4[2], ncfg.n_agents)
            mbs.append(make_minibatches(trajs[u], advs[u], tgts[u], prevs[u], batch_perm, agent_perm, M))
            # quirk kept: `_update_epoch` returns the *permuted* prev_hstates in update_state
            # (rec_magpo.py:447,471), so from the 2nd epoch on the stored Sable states are a
            # composition of all earlier env permutations while the trajectory is not.
            prevs[u] = tuple(np.take(h, batch_perm, axis=0) for h in prevs[u])
        for m in range(M):
            gsum, asum, infos = None, None, []
            for u in range(U):
                gg, ga, info, aux = minibatch_losses_and_grads(state["guider_params"], state["actor_params"], mbs[u][m], ncfg, sys, dtype)
                gsum = gg if gsum is None else {k: gsum[k] + gg[k] for k in gg}
                asum = ga if asum is None else {k: asum[k] + ga[k] for k in ga}
                infos.append(info)
            gmean = {k: (v / f(U)).astype(f) for k, v in gsum.items()}
            amean = {k: (v / f(U)).astype(f) for k, v in asum.items()}
            info = {k: float(np.mean([i[k] for i in infos])) for k in infos[0]}
            if grad_allreduce is not None:
                gmean, amean, info = grad_allreduce(gmean, amean, info)
            if record is not None:
                record["grads"].append((gmean, amean))
            clip_adam_step(state["guider_params"], gmean, state["guider_opt"], lr, sys.max_grad_norm, f=f)
            clip_adam_step(state["actor_params"], amean, state["actor_opt"], lr, sys.max_grad_norm, f=f)
            loss_infos.append(info)
    for u, slot in enumerate(state["slots"]):
        slot["key"] = keys[u]
    return mets, loss_infos


def learn(state, spec, ncfg, sys, num_updates: int):
    """learner_fn (rec_magpo.py:501-528): `num_updates` sequential update steps."""
    ep, tr = [], []
    for _ in range(num_updates):
        m, l = update_step(state, spec, ncfg, sys)
        ep.append(m); tr.append(l)
    return ep, tr
