"""TEST INFRASTRUCTURE ONLY — CPU restatement of Jumanji's RobotWarehouse as Mava's rec_magpo uses it.

The dynamics are NOT in /root/reference: they live in the un-vendored dependency jumanji 1.1.0 @ git 9ced6b8 (uv.lock:1217-1219),
`jumanji/environments/routing/robot_warehouse/{env,generator,utils,utils_agent,utils_shelf,utils_spawn}.py`, absent from this image
and not installable (no network). This file restates that package's published algorithm FROM MEMORY —
**parity unpinned**, never compared with a run of Jumanji. Anchors in the reference: construction mava/utils/make_env.py:107-135
(`RandomGenerator(**scenario.task_config)`, `jumanji.make("RobotWarehouse-v0", generator=, time_limit=)`), scenarios
mava/configs/env/scenario/{tiny-2ag,tiny-4ag,tiny-4ag-easy,small-4ag,...}.yaml, RwareWrapper mava/wrappers/jumanji.py:137-168 (float
obs, scalar reward / discount repeated over agents); the wrapper stack is oracle/wrappers.py.

Restated algorithm. Positions are (x, y) = (row, col), grid int32[2, H, W]: channel 0 = shelf id + 1, channel 1 = agent id + 1.
  layout      H = (column_height + 1) * shelf_rows + 2, W = 3 * shelf_columns + 1. highway(row, col) = col % 3 == 0 or
              row % (column_height + 1) == 0 or row == H - 1 or (row > H - (column_height + 3) and col in {W//2 - 1, W//2}).
              Shelves sit on every non-highway cell, ids in row-major order; goals = cells (H - 1, W//2 - 1), (H - 1, W//2).
  generator   key, agent_key, dir_key, queue_key = split(key, 4); agents on `choice(agent_key, H*W, (A,), replace=False)`
              (= permutation(key, H*W)[:A]) cells, directions randint(dir_key, (A,), 0, 4) (UP 0, RIGHT 1, DOWN 2, LEFT 3), not
              carrying; request queue = permutation(queue_key, S)[:Q]; State.key = key.
  step        actions NOOP 0, FORWARD 1, LEFT 2 (dir - 1), RIGHT 3 (dir + 1), TOGGLE_LOAD 4; an action whose mask bit is off becomes
              NOOP. Agents are updated sequentially in id order on the shared grid: FORWARD moves the agent (and the shelf it
              carries) one cell, clipped to the grid; TOGGLE_LOAD picks up the shelf on the agent's cell, or puts the carried one
              down unless the cell is a highway. collision = some agent's cell no longer holds its own id (two agents entered one
              cell). Then per goal, in order: a requested shelf on the goal gives reward += 1, `key, request_key = split(key)`, a
              new request = argmax(gumbel(request_key, (S,)) + log(not requested)) replaces it in the queue.
              step_count += 1; done = collision or step_count >= time_limit -> termination (discount 0), else transition.
  mask        only FORWARD is ever illegal: next cell outside the grid, holding an agent, or (carrying and holding a shelf).
  observation per agent 8 + 7 * (2 * sensor_range + 1)^2 int32: [x, y, carrying, one_hot(dir, 4), on_highway] then, for every cell of
              the zero-padded window in row-major order, [agent present, one_hot(that agent's dir, 4), shelf present, shelf requested].
Known unknowns: the order of the key splits, (row, col) vs (col, row) in the first two features, the
collision rule's exact form, and the sampler used for the new request.
"""
from dataclasses import dataclass

import numpy as np

from . import prng, wrappers

NOOP, FORWARD, LEFT, RIGHT, TOGGLE_LOAD = range(5)
DIR_DELTA = np.array([[-1, 0], [0, 1], [1, 0], [0, -1]], np.int32)  # UP, RIGHT, DOWN, LEFT

STEP_FIRST, STEP_MID, STEP_LAST = wrappers.STEP_FIRST, wrappers.STEP_MID, wrappers.STEP_LAST


@dataclass(frozen=True)
class RwareSpec:
    column_height: int = 8
    shelf_rows: int = 1
    shelf_columns: int = 3
    num_agents: int = 4
    sensor_range: int = 1
    request_queue_size: int = 4
    time_limit: int = 500

    @property
    def grid_size(self):
        return (self.column_height + 1) * self.shelf_rows + 2, 3 * self.shelf_columns + 1

    @property
    def highways(self):
        H, W = self.grid_size
        r, c = np.mgrid[0:H, 0:W]
        return ((c % 3 == 0) | (r % (self.column_height + 1) == 0) | (r == H - 1)
                | ((r > H - (self.column_height + 3)) & ((c == W // 2 - 1) | (c == W // 2))))

    @property
    def shelf_cells(self):
        return np.argwhere(~self.highways).astype(np.int32)  # row-major

    @property
    def goals(self):
        H, W = self.grid_size
        return [(H - 1, W // 2 - 1), (H - 1, W // 2)]

    @property
    def view_dim(self):
        return 8 + 7 * (2 * self.sensor_range + 1) ** 2

    @property
    def obs_dim(self):
        return self.num_agents + self.view_dim

    @property
    def action_dim(self):
        return 5


# mava/configs/env/scenario/*.yaml task_config of the RobotWarehouse scenarios
def _sc(h, r, c, n, q):
    return dict(column_height=h, shelf_rows=r, shelf_columns=c, num_agents=n, sensor_range=1, request_queue_size=q)


SCENARIOS = {
    "tiny-2ag": _sc(8, 1, 3, 2, 2), "tiny-2ag-hard": _sc(8, 1, 3, 2, 1), "tiny-4ag": _sc(8, 1, 3, 4, 4),
    "tiny-4ag-easy": _sc(8, 1, 3, 4, 8), "tiny-4ag-hard": _sc(8, 1, 3, 4, 2), "small-4ag": _sc(8, 2, 3, 4, 4),
    "small-4ag-hard": _sc(8, 2, 3, 4, 2), "medium-4ag": _sc(8, 2, 5, 4, 4), "medium-4ag-hard": _sc(8, 2, 5, 4, 2),
    "medium-6ag": _sc(8, 2, 5, 6, 6),
}


def _action_mask(spec, grid, agent_pos, agent_dir, agent_carry):
    H, W = spec.grid_size
    A = spec.num_agents
    mask = np.ones((A, 5), bool)
    for i in range(A):
        nx, ny = agent_pos[i] + DIR_DELTA[agent_dir[i]]
        ok = 0 <= nx < H and 0 <= ny < W
        if ok:
            ok = grid[1, nx, ny] == 0 and not (agent_carry[i] and grid[0, nx, ny] > 0)
        mask[i, FORWARD] = ok
    return mask


def _generate_one(spec: RwareSpec, key):
    H, W = spec.grid_size
    A, Q = spec.num_agents, spec.request_queue_size
    cells = spec.shelf_cells
    S = len(cells)
    key, agent_key, dir_key, queue_key = prng.split(key, 4)
    flat = prng.permutation(agent_key, H * W)[:A]
    agent_pos = np.stack(np.divmod(flat, W), axis=1).astype(np.int32)
    agent_dir = prng.randint(dir_key, (A,), 0, 4)
    queue = prng.permutation(queue_key, S)[:Q].astype(np.int32)
    shelf_req = np.zeros(S, bool)
    shelf_req[queue] = True
    grid = np.zeros((2, H, W), np.int32)
    grid[0, cells[:, 0], cells[:, 1]] = np.arange(1, S + 1)
    grid[1, agent_pos[:, 0], agent_pos[:, 1]] = np.arange(1, A + 1)
    carry = np.zeros(A, bool)
    return dict(grid=grid, agent_pos=agent_pos, agent_dir=agent_dir.astype(np.int32), agent_carry=carry, shelf_pos=cells.copy(),
                shelf_req=shelf_req, request_queue=queue, step_count=np.int32(0),
                action_mask=_action_mask(spec, grid, agent_pos, agent_dir, carry), key=np.asarray(key, np.uint32))


def base_reset(spec: RwareSpec, keys):
    keys = np.asarray(keys, np.uint32).reshape(-1, 2)
    envs = [_generate_one(spec, k) for k in keys]
    return {f: np.stack([e[f] for e in envs]) for f in envs[0]}


def _step_one(spec: RwareSpec, e, actions):
    H, W = spec.grid_size
    A = spec.num_agents
    hw = spec.highways
    grid, pos, dirs, carry = e["grid"].copy(), e["agent_pos"].copy(), e["agent_dir"].copy(), e["agent_carry"].copy()
    spos, sreq, queue = e["shelf_pos"].copy(), e["shelf_req"].copy(), e["request_queue"].copy()
    key = e["key"]
    for i in range(A):  # sequential update in agent order
        act = int(actions[i])
        if not (0 <= act < 5) or not e["action_mask"][i, act]:
            act = NOOP
        if act == LEFT:
            dirs[i] = (dirs[i] - 1) % 4
        elif act == RIGHT:
            dirs[i] = (dirs[i] + 1) % 4
        elif act == FORWARD:
            x, y = pos[i]
            nx = min(max(x + DIR_DELTA[dirs[i], 0], 0), H - 1)
            ny = min(max(y + DIR_DELTA[dirs[i], 1], 0), W - 1)
            grid[1, x, y] = 0
            grid[1, nx, ny] = i + 1
            if carry[i]:
                sid = grid[0, x, y]
                grid[0, x, y] = 0
                grid[0, nx, ny] = sid
                if sid > 0:
                    spos[sid - 1] = (nx, ny)
            pos[i] = (nx, ny)
        elif act == TOGGLE_LOAD:
            x, y = pos[i]
            if not carry[i]:
                carry[i] = grid[0, x, y] > 0
            elif not hw[x, y]:
                carry[i] = False
    collision = any(grid[1, pos[i, 0], pos[i, 1]] != i + 1 for i in range(A))
    reward = np.float32(0.0)
    for gx, gy in spec.goals:
        sid = grid[0, gx, gy]
        if sid > 0 and sreq[sid - 1]:
            reward = np.float32(reward + np.float32(1.0))
            key, request_key = prng.split(key)
            with np.errstate(divide="ignore"):
                score = (prng.gumbel(request_key, (len(sreq),)) + np.log((~sreq).astype(np.float32))).astype(np.float32)
            new_id = int(np.argmax(score))  # top-1, ties to the lower index
            queue[queue == sid - 1] = new_id
            sreq[sid - 1] = False
            sreq[new_id] = True
    steps = np.int32(e["step_count"] + 1)
    done = collision or steps >= spec.time_limit
    new = dict(grid=grid, agent_pos=pos, agent_dir=dirs, agent_carry=carry, shelf_pos=spos, shelf_req=sreq, request_queue=queue,
               step_count=steps, action_mask=_action_mask(spec, grid, pos, dirs, carry), key=np.asarray(key, np.uint32))
    return new, reward, done


def base_step(spec: RwareSpec, base, actions):
    B, A = actions.shape
    outs = [_step_one(spec, {k: v[b] for k, v in base.items()}, actions[b]) for b in range(B)]
    new_base = {f: np.stack([o[0][f] for o in outs]) for f in base}
    reward = np.array([o[1] for o in outs], np.float32)
    done = np.array([o[2] for o in outs], bool)
    # RwareWrapper: scalar reward repeated; termination for both collision and the horizon (discount 0)
    return new_base, np.repeat(reward[:, None], A, axis=1), done, np.zeros(B, bool)


def observe(spec: RwareSpec, base):
    H, W = spec.grid_size
    A, sr = spec.num_agents, spec.sensor_range
    hw = spec.highways
    B = base["grid"].shape[0]
    rf = 2 * sr + 1
    view = np.zeros((B, A, spec.view_dim), np.int32)
    for b in range(B):
        grid = np.pad(base["grid"][b], ((0, 0), (sr, sr), (sr, sr)))
        for i in range(A):
            x, y = base["agent_pos"][b, i]
            o = view[b, i]
            o[0], o[1], o[2] = x, y, int(base["agent_carry"][b, i])
            o[3 + base["agent_dir"][b, i]] = 1
            o[7] = int(hw[x, y])
            k = 8
            for dx in range(rf):
                for dy in range(rf):
                    aid, sid = grid[1, x + dx, y + dy], grid[0, x + dx, y + dy]
                    if aid > 0:
                        o[k] = 1
                        o[k + 1 + base["agent_dir"][b, aid - 1]] = 1
                    if sid > 0:
                        o[k + 5] = 1
                        o[k + 6] = int(base["shelf_req"][b, sid - 1])
                    k += 7
    return view.astype(np.float32), base["action_mask"].copy()


def reset(spec: RwareSpec, keys):
    return wrappers.reset(spec, _MOD, keys)


def step(spec: RwareSpec, state, actions):
    return wrappers.step(spec, _MOD, state, actions)


class _MOD:
    base_reset = staticmethod(base_reset)
    base_step = staticmethod(base_step)
    observe = staticmethod(observe)
