"""Generates tests/golden/*.json|npz from the CPU oracle.

The reference ships no golden vectors and cannot be imported here (JAX is absent), so these
fixtures pin the ORACLE's current outputs (regression anchors shared by the CPU and the GPU tests); the
published Random123 / jax.random KATs are asserted separately in tests/test_oracle.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import coordsum as ocs  # noqa: E402
from oracle import prng  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    key = prng.prng_key(42)
    gold = dict(seed=42, split4=prng.split(key, 4).tolist(), bits8=prng.random_bits(key, (8,)).tolist(),
                randint8_0_30=prng.randint(key, (8,), 0, 30).tolist(), perm16=prng.permutation(key, 16).tolist())
    with open(os.path.join(HERE, "prng.json"), "w") as f:
        json.dump(gold, f, indent=1)
    # CoordSum trace: 4 envs, 105 steps (through one auto-reset) with a fixed action stream
    spec = ocs.CoordSumSpec(**ocs.SCENARIOS["3x10-30-v0"])
    keys = prng.split(prng.prng_key(5), 4)
    st, ts = ocs.reset(spec, keys)
    rng = np.random.default_rng(0)
    acts, rewards, views, steps = [], [], [ts["observation"]["agents_view"].copy()], []
    for _ in range(105):
        a = rng.integers(0, 10, (4, 3)).astype(np.int32)
        st, ts = ocs.step(spec, st, a)
        acts.append(a); rewards.append(ts["reward"].copy()); views.append(ts["observation"]["agents_view"].copy())
        steps.append(ts["step_type"].copy())
    np.savez_compressed(os.path.join(HERE, "coordsum_trace.npz"), keys=keys, actions=np.stack(acts), rewards=np.stack(rewards),
                        agents_view=np.stack(views), step_type=np.stack(steps), final_target=st["env_state"]["target"],
                        final_key=st["env_state"]["key"], episode_return=st["episode_return"])


if __name__ == "__main__":
    main()
