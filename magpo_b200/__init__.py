"""magpo_b200 — B200-native (sm_100a) implementation of MAGPO's `rec_magpo` Anakin rollout+update path.

Host side mirrors `mava/systems/gpo/anakin/rec_magpo.py`; all compute runs in hand-written CUDA kernels
behind the C ABI of include/magpo_b200.h (magpo_b200/lib/libmagpo_b200.so). No CPU fallback exists.
"""
__version__ = "0.1"
