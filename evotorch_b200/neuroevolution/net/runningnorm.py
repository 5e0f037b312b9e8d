"""`evotorch.neuroevolution.net.runningnorm` import path (the implementation lives in neuroevolution/runningnorm.py)."""

from ..runningnorm import CollectedStats, ObsNormLayer, RunningNorm

__all__ = ["RunningNorm", "ObsNormLayer", "CollectedStats"]
