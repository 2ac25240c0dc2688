"""CPU: the LevelBasedForaging restatement (oracle/lbf.py) against hand-computed cases and generator invariants.
These pin the restatement, not a Jumanji run (parity unpinned, see the oracle's header)."""
import numpy as np
import pytest

from oracle import lbf, prng, wrappers


def _base(spec, agent_pos, agent_level, food_pos, food_level, eaten=None, step=0):
    A, F = spec.num_agents, spec.num_food
    return dict(agent_pos=np.array([agent_pos], np.int32), agent_level=np.array([agent_level], np.int32),
                agent_loading=np.zeros((1, A), bool), food_pos=np.array([food_pos], np.int32),
                food_level=np.array([food_level], np.int32), food_eaten=np.array([eaten or [False] * F], bool),
                step_count=np.array([step], np.int32), key=np.array([[1, 2]], np.uint32))


@pytest.mark.parametrize("scenario", list(lbf.SCENARIOS))
def test_generator_invariants(scenario):
    spec = lbf.LbfSpec(**lbf.SCENARIOS[scenario])
    G = spec.grid_size
    b = lbf.base_reset(spec, prng.split(prng.prng_key(3), 200))
    fp, ap = b["food_pos"], b["agent_pos"]
    assert ((fp >= 1) & (fp <= G - 2)).all(), "food on the border"
    for i in range(spec.num_food):
        for j in range(i + 1, spec.num_food):
            assert (np.abs(fp[:, i] - fp[:, j]).sum(-1) >= 2).all(), "adjacent or coincident food"
    assert ((ap >= 0) & (ap < G)).all()
    for i in range(spec.num_agents):
        for j in range(i + 1, spec.num_agents):
            assert (ap[:, i] != ap[:, j]).any(-1).all(), "two agents on one cell"
        assert (ap[:, i, None, :] != fp).any(-1).all(), "agent on food"
        # the row-clearing agent mask: no agent row equals any food coordinate
        assert (ap[:, i, 0, None] != fp.reshape(len(fp), -1)).all()
    assert ((b["agent_level"] >= 1) & (b["agent_level"] <= spec.max_agent_level)).all()
    cap = np.sort(b["agent_level"], -1)[:, :3].sum(-1)
    if spec.force_coop:
        assert (b["food_level"] == cap[:, None]).all()
    else:
        assert ((b["food_level"] >= 1) & (b["food_level"] <= cap[:, None])).all()
    assert len({tuple(k) for k in b["key"]}) == 200
    # both levels and every interior cell occur
    assert set(np.unique(b["agent_level"])) == set(range(1, spec.max_agent_level + 1))
    if G <= 8:
        assert len(np.unique(fp[..., 0] * G + fp[..., 1])) == (G - 2) ** 2


def test_choice_is_kth_open_cell():
    """searchsorted(cumsum(mask), total * (1 - u)) == index of the ceil(r)-th open cell (the form the CUDA kernel uses)."""
    rng = np.random.default_rng(0)
    for _ in range(300):
        mask = rng.random(64) < 0.4
        mask[rng.integers(64)] = True
        u = np.float32(rng.random())
        cum = np.cumsum(mask.astype(np.float32), dtype=np.float32)
        r = np.float32(cum[-1] * np.float32(np.float32(1) - u))
        want = int(np.searchsorted(cum, r, side="left"))
        k = int(np.ceil(r))
        assert want == np.nonzero(mask)[0][k - 1]


def test_movement_rules():
    spec = lbf.LbfSpec(grid_size=8, fov=2, num_agents=3, num_food=2)
    #  agent0 (0,0) moves UP -> out of bounds; agent1 (3,3) moves RIGHT onto food (3,4) -> refused; agent2 (5,5) moves LEFT -> ok
    b = _base(spec, [[0, 0], [3, 3], [5, 5]], [1, 1, 2], [[3, 4], [6, 6]], [2, 2])
    nb, r, term, trunc = lbf.base_step(spec, b, np.array([[lbf.UP, lbf.RIGHT, lbf.LEFT]], np.int32))
    assert nb["agent_pos"][0].tolist() == [[0, 0], [3, 3], [5, 4]]
    assert (r == 0).all() and not term[0] and not trunc[0] and nb["step_count"][0] == 1
    # an eaten food no longer blocks
    b["food_eaten"][0, 0] = True
    nb, *_ = lbf.base_step(spec, b, np.array([[lbf.NOOP, lbf.RIGHT, lbf.NOOP]], np.int32))
    assert nb["agent_pos"][0, 1].tolist() == [3, 4]
    # moving onto another agent's CURRENT cell is refused even if that agent moves away
    b = _base(spec, [[2, 2], [2, 3], [7, 7]], [1, 1, 1], [[5, 5], [5, 1]], [1, 1])
    nb, *_ = lbf.base_step(spec, b, np.array([[lbf.RIGHT, lbf.RIGHT, lbf.NOOP]], np.int32))
    assert nb["agent_pos"][0].tolist() == [[2, 2], [2, 4], [7, 7]]
    # two agents targeting the same free cell both stay
    b = _base(spec, [[2, 2], [2, 4], [7, 7]], [1, 1, 1], [[5, 5], [5, 1]], [1, 1])
    nb, *_ = lbf.base_step(spec, b, np.array([[lbf.RIGHT, lbf.LEFT, lbf.NOOP]], np.int32))
    assert nb["agent_pos"][0].tolist() == [[2, 2], [2, 4], [7, 7]]


def test_loading_reward_and_termination():
    spec = lbf.LbfSpec(grid_size=8, fov=2, num_agents=2, num_food=2)
    # food0 level 3 at (3,3) with agents (level 1) at (3,2) and (level 2) at (2,3); food1 level 3 far away
    b = _base(spec, [[3, 2], [2, 3]], [1, 2], [[3, 3], [6, 6]], [3, 3])
    both = np.array([[lbf.LOAD, lbf.LOAD]], np.int32)
    nb, r, term, trunc = lbf.base_step(spec, b, both)
    assert nb["food_eaten"][0].tolist() == [True, False] and not term[0]
    # per agent: level_a * 3 / ((1 + 2) * (3 + 3)) -> 1/6 and 2/6; LbfWrapper: team sum repeated
    want = np.float32(np.float32(3) / np.float32(18)) + np.float32(np.float32(6) / np.float32(18))
    assert (r[0] == want).all() and abs(float(want) - 0.5) < 1e-6
    # one loader alone is too weak: nothing eaten, no reward
    nb2, r2, *_ = lbf.base_step(spec, b, np.array([[lbf.LOAD, lbf.NOOP]], np.int32))
    assert not nb2["food_eaten"][0].any() and (r2 == 0).all() and nb2["agent_loading"][0].tolist() == [True, False]
    # eating the last food terminates (discount 0); the wrapper stack auto-resets and publishes the return
    b["food_eaten"][0, 1] = True
    st = dict(env_state=b, key=np.array([[9, 9]], np.uint32), running_count_episode_return=np.array([0.25], np.float32),
              running_count_episode_length=np.array([7], np.int32), episode_return=np.zeros(1, np.float32),
              episode_length=np.zeros(1, np.int32))
    nst, ts = lbf.step(spec, st, both)
    assert ts["step_type"][0] == lbf.STEP_LAST and (ts["discount"][0] == 0).all()
    assert ts["extras"]["episode_metrics"]["episode_return"][0] == np.float32(0.25) + want
    assert ts["extras"]["episode_metrics"]["episode_length"][0] == 8
    assert nst["env_state"]["step_count"][0] == 0 and not nst["env_state"]["food_eaten"].any()
    assert (nst["env_state"]["key"][0] != b["key"][0]).any()
    assert ts["observation"]["step_count"][0].tolist() == [0, 0] and ts["extras"]["real_next_obs"]["step_count"][0].tolist() == [1, 1]
    # the reset used key, _ = split(state.key)
    rb = lbf.base_reset(spec, prng.split(b["key"][0])[:1])
    assert (rb["agent_pos"] == nst["env_state"]["agent_pos"]).all() and (rb["key"] == nst["env_state"]["key"]).all()


def test_truncation_keeps_discount():
    spec = lbf.LbfSpec(grid_size=8, fov=2, num_agents=2, num_food=2, time_limit=100)
    b = _base(spec, [[0, 0], [7, 7]], [1, 1], [[3, 3], [5, 5]], [2, 2], step=99)
    st = dict(env_state=b, key=np.array([[9, 9]], np.uint32), running_count_episode_return=np.zeros(1, np.float32),
              running_count_episode_length=np.array([99], np.int32), episode_return=np.zeros(1, np.float32),
              episode_length=np.zeros(1, np.int32))
    _, ts = lbf.step(spec, st, np.zeros((1, 2), np.int32))
    assert ts["step_type"][0] == lbf.STEP_LAST and (ts["discount"][0] == 1).all()
    assert ts["extras"]["episode_metrics"]["episode_length"][0] == 100


def test_observation_and_mask():
    spec = lbf.LbfSpec(grid_size=8, fov=2, num_agents=2, num_food=2)
    # agent0 (level 2) at (1,5): offsets min(fov, pos) = (1, 2); food0 (3,4) visible -> (3-1+1, 4-5+2) = (3, 1); food1 (6,6) not
    # visible; agent1 (level 1) at (2,7) visible from agent0 -> (2-1+1, 7-5+2) = (2, 4)
    b = _base(spec, [[1, 5], [2, 7]], [2, 1], [[3, 4], [6, 6]], [3, 3])
    view, mask = lbf.observe(spec, b)
    assert view[0, 0].tolist() == [3, 1, 3, -1, -1, 0, 1, 2, 2, 2, 4, 1]
    # agent1: offsets (2, 2); food0 |2-3|<=2, |7-4|=3 -> invisible; self (2,2,1); agent0 -> (1-2+2, 5-7+2) = (1, 0)
    assert view[0, 1].tolist() == [-1, -1, 0, -1, -1, 0, 2, 2, 1, 1, 0, 2]
    assert mask[0, 0].tolist() == [True, True, True, True, True, False]
    assert mask[0, 1].tolist() == [True, True, True, True, False, False]  # RIGHT leaves the grid
    # LOAD legal next to an un-eaten food, moving onto it is not; the AgentID one-hot comes first in the wrapped observation
    b = _base(spec, [[3, 3], [0, 0]], [2, 1], [[3, 4], [6, 6]], [3, 3])
    _, mask = lbf.observe(spec, b)
    assert mask[0, 0].tolist() == [True, True, True, True, False, True]
    obs = wrappers._observation(spec, lbf._MOD, b)
    assert obs["agents_view"].shape == (1, 2, 14) and obs["agents_view"][0, :, :2].tolist() == [[1, 0], [0, 1]]
    assert obs["agents_view"].dtype == np.float32
