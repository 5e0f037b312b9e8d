from . import functional
from .cmaes import CMAES
from .gaussian import CEM, PGPE, SNES, XNES, GaussianSearchAlgorithm
from .searchalgorithm import LazyReporter, LazyStatusDict, SearchAlgorithm, SinglePopulationAlgorithmMixin

__all__ = ["CMAES", "PGPE", "SNES", "CEM", "XNES", "GaussianSearchAlgorithm", "SearchAlgorithm", "LazyReporter", "LazyStatusDict",
           "SinglePopulationAlgorithmMixin", "functional"]
