from .policy import Policy, count_parameters, fill_parameters, parameter_vector
from .rollout import RolloutResult, rollout
from .runningnorm import CollectedStats, ObsNormLayer, RunningNorm
from .vecne import VecNE

__all__ = ["Policy", "count_parameters", "fill_parameters", "parameter_vector", "RunningNorm", "ObsNormLayer", "CollectedStats", "rollout",
           "RolloutResult", "VecNE"]
