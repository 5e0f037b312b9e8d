"""Import paths of the reference's `evotorch.neuroevolution.net` for the pieces on the hot path."""

from ..policy import Policy, count_parameters, fill_parameters, parameter_vector
from ..runningnorm import CollectedStats, ObsNormLayer, RunningNorm
from . import multilayered, parser, runningnorm, runningstat, vecrl
from .runningstat import RunningStat
from .multilayered import MultiLayered
from .parser import NetParsingError, str_to_net

__all__ = ["Policy", "count_parameters", "fill_parameters", "parameter_vector", "RunningNorm", "ObsNormLayer", "CollectedStats", "runningnorm",
           "vecrl", "multilayered", "parser", "MultiLayered", "NetParsingError", "str_to_net", "runningstat", "RunningStat"]
