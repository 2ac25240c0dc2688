"""TEST INFRASTRUCTURE ONLY — CPU restatement of Jumanji's LevelBasedForaging as Mava's rec_magpo uses it.

The dynamics are NOT in /root/reference: they live in the un-vendored dependency jumanji 1.1.0 @ git 9ced6b8 (uv.lock:1217-1219),
`jumanji/environments/routing/lbf/{env,generator,observer,utils,constants}.py`, which is absent from this image and cannot be
installed (no network). This file restates that package's published algorithm FROM MEMORY — **parity unpinned**: it was never
compared with a run of Jumanji. It is anchored on the reference's own call sites:
  * construction        mava/utils/make_env.py:107-135 (`RandomGenerator(**scenario.task_config)`, `jumanji.make(name, generator=,
                        time_limit=)`), scenarios mava/configs/env/scenario/{2s-8x8-2p-2f-coop,8x8-2p-2f-coop,2s-10x10-3p-3f,...}.yaml
  * Mava wrapper        mava/wrappers/jumanji.py:171-208 (LbfWrapper: obs cast to float, team reward = sum repeated — the
                        `aggregate_rewards` config flag is never passed, make_env.py:131, so the default True always applies)
  * stack               oracle/wrappers.py

Restated algorithm (positions are (row, col); actions NOOP, UP(-1,0), DOWN(+1,0), LEFT(0,-1), RIGHT(0,+1), LOAD):
  generator   key_food, key_agents, key_food_level, key_agent_level, key = split(key, 5)
              food: per item i, key_i = split(key_food, F)[i]; `jax.random.choice(key_i, G*G, (), p=mask)` (cumsum / uniform /
              searchsorted-left) over the interior cells; the chosen cell and its 4 neighbours leave the mask.
              agents: `choice(key_agents, G*G, (A,), replace=False, p=mask)` = top-A of gumbel(key,(G*G,)) + log(mask);
              the mask is built as `ones((G,G)).at[food_positions].set(False)` which indexes the FIRST axis with every food
              coordinate value, i.e. it clears whole rows r for r in {food rows} U {food cols} (known unknown: set
              AGENT_MASK_CLEARS_ROWS = False for the per-cell reading).
              levels: agents randint(key_agent_level, (A,), 1, max_agent_level+1); max_food_level = sum of the 3 lowest agent
              levels; food = max_food_level if force_coop else randint(key_food_level, (F,), 1, max_food_level+1).
  step        every agent proposes pos + MOVES[action]; the move is refused when out of bounds, onto an un-eaten food, or onto
              another agent's CURRENT cell; agents whose proposed cells coincide all stay (one pass); loading = action == LOAD;
              per food: adjacent (L1 distance 1) loading agents' levels are summed, food is eaten when the sum >= its level;
              reward[a] = sum_f level_a * level_f * eaten_f / (sum_adj_f * sum_f' level_f') (nan -> 0), fp32;
              step_count += 1; terminate = all eaten (discount 0), truncate = step_count >= time_limit (discount 1).
  observation VectorObserver: per agent [food (row', col', level)] * F, self, the other agents in id order; an entity is visible when
              both coordinate distances are <= fov (food: and not eaten); visible coordinates are pos - self + min(fov, self);
              invisible entries are (-1, -1, 0). action_mask[a] = move a lands in bounds on a free cell (NOOP always legal), LOAD
              legal iff an un-eaten food is adjacent.
"""
from dataclasses import dataclass

import numpy as np

from . import prng, wrappers

AGENT_MASK_CLEARS_ROWS = True  # see the header: `mask.at[food_positions].set(False)` on a 2-D mask

NOOP, UP, DOWN, LEFT, RIGHT, LOAD = range(6)
MOVES = np.array([[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1], [0, 0]], np.int32)

STEP_FIRST, STEP_MID, STEP_LAST = wrappers.STEP_FIRST, wrappers.STEP_MID, wrappers.STEP_LAST


@dataclass(frozen=True)
class LbfSpec:
    grid_size: int = 8
    fov: int = 2
    num_agents: int = 2
    num_food: int = 2
    max_agent_level: int = 2
    force_coop: bool = True
    time_limit: int = 100

    @property
    def obs_dim(self):  # agent-id one-hot + 3 * (food + agents)
        return self.num_agents + 3 * (self.num_food + self.num_agents)

    @property
    def action_dim(self):
        return 6


# mava/configs/env/scenario/*.yaml task_config of the LevelBasedForaging scenarios
SCENARIOS = {
    "2s-8x8-2p-2f-coop": dict(grid_size=8, fov=2, num_agents=2, num_food=2, max_agent_level=2, force_coop=True),
    "8x8-2p-2f-coop": dict(grid_size=8, fov=8, num_agents=2, num_food=2, max_agent_level=2, force_coop=True),
    "2s-10x10-3p-3f": dict(grid_size=10, fov=2, num_agents=3, num_food=3, max_agent_level=2, force_coop=False),
    "10x10-3p-3f": dict(grid_size=10, fov=10, num_agents=3, num_food=3, max_agent_level=2, force_coop=False),
    "15x15-3p-5f": dict(grid_size=15, fov=15, num_agents=3, num_food=5, max_agent_level=2, force_coop=False),
    "15x15-4p-3f": dict(grid_size=15, fov=15, num_agents=4, num_food=3, max_agent_level=2, force_coop=False),
    "15x15-4p-5f": dict(grid_size=15, fov=15, num_agents=4, num_food=5, max_agent_level=2, force_coop=False),
}


def _randint_span(key, n, minval, span):
    """jax.random.randint(key, (n,), minval, minval + span) with a run-time span (Appendix A4)."""
    k1, k2 = prng.split(key)
    hi, lo = prng.random_bits(k1, (n,)), prng.random_bits(k2, (n,))
    span = np.uint32(max(1, span))
    mult = np.uint32(((65536 % int(span)) ** 2) % int(span))
    with np.errstate(over="ignore"):
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int32(minval) + off.astype(np.int32)).astype(np.int32)


def _generate_one(spec: LbfSpec, key):
    G, A, F = spec.grid_size, spec.num_agents, spec.num_food
    ks = prng.split(key, 5)
    key_food, key_agents, key_food_level, key_agent_level, key = ks
    # --- sample_food
    mask = np.ones((G, G), bool)
    mask[0, :] = mask[-1, :] = False
    mask[:, 0] = mask[:, -1] = False
    mask = mask.ravel()
    food_flat = np.zeros(F, np.int64)
    for i, k in enumerate(prng.split(key_food, F)):
        cum = np.cumsum(mask.astype(np.float32), dtype=np.float32)
        u = prng.uniform(k, ())
        r = np.float32(cum[-1] * np.float32(np.float32(1.0) - u))
        pos = int(np.searchsorted(cum, r, side="left"))
        food_flat[i] = pos
        for adj in (pos, pos + 1, pos - 1, pos + G, pos - G):
            if 0 <= adj < G * G:
                mask[adj] = False
    food_pos = np.stack(np.divmod(food_flat, G), axis=1).astype(np.int32)
    # --- sample_agents
    amask = np.ones((G, G), bool)
    if AGENT_MASK_CLEARS_ROWS:
        amask[food_pos.ravel()] = False
    else:
        amask[food_pos[:, 0], food_pos[:, 1]] = False
    g = prng.gumbel(key_agents, (G * G,))
    with np.errstate(divide="ignore"):
        score = (g + np.log(amask.ravel().astype(np.float32))).astype(np.float32)
    agent_flat = np.argsort(-score, kind="stable")[:A]  # lax.top_k: descending, ties to the lower index
    agent_pos = np.stack(np.divmod(agent_flat, G), axis=1).astype(np.int32)
    # --- levels
    agent_level = prng.randint(key_agent_level, (A,), 1, spec.max_agent_level + 1)
    max_food_level = int(np.sort(agent_level)[:3].sum())
    if spec.force_coop:
        food_level = np.full(F, max_food_level, np.int32)
    else:
        food_level = _randint_span(key_food_level, F, 1, max_food_level)
    return dict(agent_pos=agent_pos, agent_level=agent_level.astype(np.int32), agent_loading=np.zeros(A, bool),
                food_pos=food_pos, food_level=food_level, food_eaten=np.zeros(F, bool), step_count=np.int32(0),
                key=np.asarray(key, np.uint32))


def base_reset(spec: LbfSpec, keys):
    keys = np.asarray(keys, np.uint32).reshape(-1, 2)
    envs = [_generate_one(spec, k) for k in keys]
    return {f: np.stack([e[f] for e in envs]) for f in envs[0]}


def _adjacent(agent_pos, food_pos):
    """[B,A,2], [B,F,2] -> bool[B,F,A]: L1 distance exactly 1."""
    dist = np.abs(agent_pos[:, None, :, :] - food_pos[:, :, None, :]).sum(-1)
    return dist == 1


def base_step(spec: LbfSpec, base, actions):
    G, A, F = spec.grid_size, spec.num_agents, spec.num_food
    pos = base["agent_pos"]
    B = pos.shape[0]
    prop = pos + MOVES[actions]  # [B,A,2]
    oob = ((prop < 0) | (prop >= G)).any(-1)
    same_cell = (prop[:, :, None, :] == pos[:, None, :, :]).all(-1)  # [B, mover, other]
    other = ~np.eye(A, dtype=bool)[None]
    agent_at = (same_cell & other).any(-1)
    food_at = ((prop[:, :, None, :] == base["food_pos"][:, None, :, :]).all(-1) & ~base["food_eaten"][:, None, :]).any(-1)
    moved = np.where((oob | agent_at | food_at)[..., None], pos, prop)
    dup = ((moved[:, :, None, :] == moved[:, None, :, :]).all(-1) & other).any(-1)  # flag_duplicates
    new_pos = np.where(dup[..., None], pos, moved).astype(np.int32)
    loading = actions == LOAD
    # eat_food, vmapped over food
    adj = _adjacent(new_pos, base["food_pos"])  # [B,F,A]
    adj_levels = np.where(adj & loading[:, None, :] & ~base["food_eaten"][:, :, None], base["agent_level"][:, None, :], 0)
    sum_adj = adj_levels.sum(-1)  # [B,F]
    eaten_now = sum_adj >= base["food_level"]
    food_eaten = base["food_eaten"] | eaten_now
    # get_reward (normalize_reward=True, penalty=0)
    total_food_level = base["food_level"].sum(-1)  # [B]
    num = (adj_levels * eaten_now[:, :, None].astype(np.int32) * base["food_level"][:, :, None]).astype(np.float32)
    den = (sum_adj * total_food_level[:, None]).astype(np.float32)[:, :, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        per_food = np.nan_to_num((num / den).astype(np.float32), nan=0.0)
    reward = np.zeros((B, A), np.float32)
    for f in range(F):
        reward = (reward + per_food[:, f]).astype(np.float32)
    # LbfWrapper: team reward repeated
    team = reward[:, 0].copy()
    for i in range(1, A):
        team = (team + reward[:, i]).astype(np.float32)
    rewards = np.repeat(team[:, None], A, axis=1)
    step_count = (base["step_count"] + 1).astype(np.int32)
    terminate = food_eaten.all(-1)
    truncate = step_count >= spec.time_limit
    new_base = dict(agent_pos=new_pos, agent_level=base["agent_level"], agent_loading=loading, food_pos=base["food_pos"],
                    food_level=base["food_level"], food_eaten=food_eaten, step_count=step_count, key=base["key"])
    return new_base, rewards, terminate, truncate


def observe(spec: LbfSpec, base):
    G, A, F, fov = spec.grid_size, spec.num_agents, spec.num_food, spec.fov
    pos, fpos = base["agent_pos"], base["food_pos"]
    B = pos.shape[0]
    view = np.tile(np.array([-1, -1, 0], np.int32), (B, A, F + A))
    offs = np.minimum(fov, pos)  # [B,A,2]
    vis_f = (np.abs(pos[:, :, None, :] - fpos[:, None, :, :]) <= fov).all(-1) & ~base["food_eaten"][:, None, :]  # [B,A,F]
    tf = fpos[:, None, :, :] - pos[:, :, None, :] + offs[:, :, None, :]
    for f in range(F):
        v = vis_f[:, :, f]
        view[:, :, 3 * f + 0] = np.where(v, tf[:, :, f, 0], -1)
        view[:, :, 3 * f + 1] = np.where(v, tf[:, :, f, 1], -1)
        view[:, :, 3 * f + 2] = np.where(v, base["food_level"][:, None, f], 0)
    vis_a = (np.abs(pos[:, :, None, :] - pos[:, None, :, :]) <= fov).all(-1)  # [B, observer, other]
    ta = pos[:, None, :, :] - pos[:, :, None, :] + offs[:, :, None, :]
    for i in range(A):
        order = [i] + [j for j in range(A) if j != i]
        for slot, j in enumerate(order):
            c = 3 * (F + slot)
            v = vis_a[:, i, j]
            view[:, i, c + 0] = np.where(v, ta[:, i, j, 0], -1)
            view[:, i, c + 1] = np.where(v, ta[:, i, j, 1], -1)
            view[:, i, c + 2] = np.where(v, base["agent_level"][:, j], 0)
    # compute_action_mask
    nxt = pos[:, :, None, :] + MOVES[None, None]  # [B,A,6,2]
    oob = ((nxt < 0) | (nxt >= G)).any(-1)
    other = ~np.eye(A, dtype=bool)
    agent_occ = ((nxt[:, :, :, None, :] == pos[:, None, None, :, :]).all(-1) & other[None, :, None, :]).any(-1)
    food_occ = ((nxt[:, :, :, None, :] == fpos[:, None, None, :, :]).all(-1) & ~base["food_eaten"][:, None, None, :]).any(-1)
    mask = ~(oob | agent_occ | food_occ)
    adj = _adjacent(pos, fpos) & ~base["food_eaten"][:, :, None]  # [B,F,A]
    mask[:, :, LOAD] &= adj.any(1)
    return view.astype(np.float32), mask


def reset(spec: LbfSpec, keys):
    return wrappers.reset(spec, _MOD, keys)


def step(spec: LbfSpec, state, actions):
    return wrappers.step(spec, _MOD, state, actions)


class _Mod:
    base_reset = staticmethod(base_reset)
    base_step = staticmethod(base_step)
    observe = staticmethod(observe)


_MOD = _Mod
