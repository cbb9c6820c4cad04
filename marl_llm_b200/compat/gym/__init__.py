"""Minimal stand-in for the reference's vendored Gym 0.19 fork (cus_gym/gym) — only the surface the MARL-LLM
scripts touch: `gym.make('AssemblySwarm-v0').unwrapped`, `gym.wrappers.AssemblySwarmWrapper`,
`gym.spaces.Box`, `gym.Env`, `gym.Wrapper`.  Put `marl_llm_b200/compat` on PYTHONPATH *instead of* `cus_gym`
to run train_assembly.py / eval_assembly.py on the GPU simulator:

    PYTHONPATH=/path/to/repo:/path/to/repo/marl_llm_b200/compat python marl_llm/train/train_assembly.py

The Gym fork itself (registry, spaces zoo, upstream envs) is out of scope."""
from marl_llm_b200.assembly_env import AssemblySwarmEnv, AssemblySwarmWrapper, Box  # noqa: F401
from . import spaces, wrappers  # noqa: F401

__version__ = "0.19.0+swarm_b200"


class Env:
    """Base class kept for isinstance checks / subclassing by user code (cus_gym/gym/core.py:8)."""
    metadata = {"render.modes": []}

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    """cus_gym/gym/core.py:212-273"""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)


_REGISTRY = {"AssemblySwarm-v0": AssemblySwarmEnv}     # cus_gym/gym/envs/__init__.py:14-19


def make(env_id, **kwargs):
    """cus_gym/gym/envs/registration.py:99-110.  The reference wraps in TimeLimit and the scripts immediately take
    `.unwrapped` (train_assembly.py:49); the bare env is returned (its `.unwrapped` is itself)."""
    if env_id not in _REGISTRY:
        raise KeyError(f"{env_id}: only {sorted(_REGISTRY)} exist (the reference ships no other swarm env, SURVEY.md §0.2)")
    return _REGISTRY[env_id](**kwargs)
