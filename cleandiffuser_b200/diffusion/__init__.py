from .basic import DiffusionModel
from .sde import BaseDiffusionSDE, DiscreteDiffusionSDE, ContinuousDiffusionSDE
from .consistency import ContinuousConsistencyModel
from .edm import ContinuousEDM
from .legacy import DDPM, EDM
from .rectifiedflow import DiscreteRectifiedFlow, ContinuousRectifiedFlow
from .solvers import SUPPORTED_SOLVERS
