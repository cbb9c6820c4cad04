from marl_llm_b200.assembly_env import AssemblySwarmWrapper  # noqa: F401
