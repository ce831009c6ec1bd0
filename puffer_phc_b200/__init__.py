"""puffer_phc_b200 -- B200-native (sm_100a) implementation of puffer-phc's per-step rollout hot path.

Drop-in modules mirror the reference's call surface:

* ``puffer_phc_b200.motion_lib.MotionLibSMPL.get_motion_state``   (reference puffer_phc/motion_lib.py:549-626)
* ``puffer_phc_b200.envs.common.compute_*``                        (reference puffer_phc/envs/common.py)
* ``puffer_phc_b200.policies.running_norm.RunningNorm``            (reference puffer_phc/policies/running_norm.py)
* ``puffer_phc_b200.c_gae.compute_gae``                            (reference puffer_phc/c_gae.pyx)
* ``puffer_phc_b200.fused_step.FusedStep``                         (the whole post-physics step in one kernel)

All arithmetic runs in hand-written CUDA kernels behind the C-ABI library ``libphc_b200.so``
(include/phc_b200.h).  There is no CPU fallback: every entry point raises if the library is missing.
"""
from ._ffi import set_reference_device  # noqa: E402,F401  ("cuda" default | "cpu": whose torch rounding flags reproduce)



def install_c_gae_shim() -> str:
    """Put ``puffer_phc_b200/shims`` at the front of ``sys.path`` so that the reference's ``from c_gae import compute_gae``
    (reference puffer_phc/clean_pufferl/core.py:36) imports the CUDA-backed drop-in without an edit.  Returns the directory."""
    import os
    import sys
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
    if d not in sys.path:
        sys.path.insert(0, d)
    return d


__version__ = "0.1.3"
