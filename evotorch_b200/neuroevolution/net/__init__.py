"""Import paths of the reference's `evotorch.neuroevolution.net` for the pieces on the hot path."""

from ..policy import Policy, count_parameters, fill_parameters, parameter_vector
from ..runningnorm import CollectedStats, ObsNormLayer, RunningNorm
from . import runningnorm, vecrl

__all__ = ["Policy", "count_parameters", "fill_parameters", "parameter_vector", "RunningNorm", "ObsNormLayer", "CollectedStats", "runningnorm",
           "vecrl"]
