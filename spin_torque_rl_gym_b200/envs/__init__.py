from .vector_env import SpinTorqueVectorEnv  # noqa: F401
