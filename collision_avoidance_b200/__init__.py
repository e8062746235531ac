"""B200-native batched ORCA simulator and environment step.

Drop-in for the hot path of navallo/collision_avoidance: the ``rvo2.PyRVOSimulator`` calls
and the gym / ALAN per-step arithmetic around them, over a batch of independent worlds.
Python reaches hand-written sm_100a CUDA kernels through ``liborca_b200.so`` (ctypes, see
``include/orca_b200.h``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import (DONE_GOAL_RADIUS, DONE_NONE, DONE_X_BELOW, POLICY_ALAN, POLICY_EXTERNAL, POLICY_GOAL,  # noqa: F401
                   POLICY_RL, OrcaLibraryError)

__all__ = ["BatchedRVOSimulator", "scenarios", "POLICY_EXTERNAL", "POLICY_GOAL", "POLICY_RL", "POLICY_ALAN",
           "DONE_NONE", "DONE_GOAL_RADIUS", "DONE_X_BELOW", "OrcaLibraryError"]


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import collision_avoidance_b200` stays cheap
    if name == "BatchedRVOSimulator":
        from .sim import BatchedRVOSimulator
        return BatchedRVOSimulator
    if name in ("scenarios", "sim", "rvo2_compat", "envs", "alan", "actfile", "dist", "mcmc", "policy"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
