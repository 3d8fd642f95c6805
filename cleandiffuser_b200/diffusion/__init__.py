from .basic import DiffusionModel
from .sde import BaseDiffusionSDE, DiscreteDiffusionSDE, ContinuousDiffusionSDE
from .consistency import ContinuousConsistencyModel
from .solvers import SUPPORTED_SOLVERS
