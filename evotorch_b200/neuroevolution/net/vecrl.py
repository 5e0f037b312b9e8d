"""`evotorch.neuroevolution.net.vecrl` import path for `Policy` (the implementation lives in neuroevolution/policy.py)."""

from ..policy import Policy

__all__ = ["Policy"]
