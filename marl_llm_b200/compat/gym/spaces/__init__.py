from marl_llm_b200.assembly_env import Box  # noqa: F401
