"""CPU oracle for the rec_magpo hot path — TEST INFRASTRUCTURE ONLY.

A NumPy (integer/PRNG/env) + torch-CPU-fp32 (network math, autograd for the
gradients) restatement of the reference's Anakin rollout+update loop
(`/root/reference/mava/systems/gpo/anakin/rec_magpo.py`) and of the third-party
arithmetic it relies on (jax.random threefry2x32, flax GRUCell/RMSNorm/GroupNorm,
distrax Categorical, optax clip+adam).  Every function cites the reference
file:line (or the SURVEY.md appendix) it follows.

PARITY UNPINNED: the reference ships no tests / golden vectors and JAX cannot be
installed in this environment, so nothing here was checked against a run of the
reference itself.  What pins it instead (tests/test_oracle_*.py): Random123
threefry KATs, the `split(PRNGKey(0))` KATs, recurrent == chunkwise retention,
GAE vs. the O(T^2) direct sum, closed-form Adam, hand-computed CoordSum / LBF / RWARE cases.
lbf.py and rware.py restate the un-vendored jumanji 1.1.0 envs (see their headers); wrappers.py is the
shared training wrapper stack.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`magpo_b200/`) never does.
"""
