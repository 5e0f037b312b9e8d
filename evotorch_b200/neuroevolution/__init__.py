from .neproblem import BaseNEProblem, NEProblem
from .policy import Policy, count_parameters, fill_parameters, parameter_vector
from .rollout import RolloutResult, rollout
from .runningnorm import CollectedStats, ObsNormLayer, RunningNorm
from .supervisedne import SupervisedNE
from .vecne import VecNE

__all__ = ["Policy", "count_parameters", "fill_parameters", "parameter_vector", "RunningNorm", "ObsNormLayer", "CollectedStats", "rollout",
           "RolloutResult", "VecNE", "NEProblem", "BaseNEProblem", "SupervisedNE"]
