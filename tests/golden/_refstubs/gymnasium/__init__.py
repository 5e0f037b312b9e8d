"""Import-time stand-in for `gymnasium` (absent in this image). Test fixture only."""


class Env:
    pass


class Wrapper:
    def __init__(self, *a, **k):
        pass


class ObservationWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass


class RewardWrapper(Wrapper):
    pass


from . import spaces, vector  # noqa: E402


def make(*a, **k):
    raise RuntimeError("gymnasium is not available (stub)")


__version__ = "1.0.0"
