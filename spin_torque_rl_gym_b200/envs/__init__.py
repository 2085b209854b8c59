from .vector_env import SpinTorqueVectorEnv  # noqa: F401
from .array_env import SpinTorqueArrayVectorEnv  # noqa: F401
