"""marl_llm_b200 — B200-native batched simulator for the step() hot path of MARL-LLM's assembly swarm env.

Only what the path needs lives here: csrc/ (sm_100a kernels + C ABI), the ctypes binding, the batched host
class and the drop-in mirror of the reference's env class.  See DESIGN.md / INTEGRATION.md."""
from ._lib import SwarmError  # noqa: F401

__all__ = ["SwarmError", "BatchedAssemblySim"]


def __getattr__(name):
    if name == "BatchedAssemblySim":
        from .batched import BatchedAssemblySim
        return BatchedAssemblySim
    raise AttributeError(name)
