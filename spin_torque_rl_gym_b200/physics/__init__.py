from .thermal_model import ThermalFluctuations  # noqa: F401
from .llgs_solver import LLGSSolver  # noqa: F401
from .simple_solver import RobustLLGSSolver, SimpleLLGSSolver  # noqa: F401
from .energy_landscape import EnergyLandscape  # noqa: F401
from .vectorized_operations import VectorizedMagneticsOperations, VectorizedSolver  # noqa: F401
