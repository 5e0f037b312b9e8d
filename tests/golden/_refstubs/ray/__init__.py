"""Import-time stand-in for `ray` (absent in this image) so that the read-only
reference at /root/reference can be imported by tests/golden/gen_golden.py.
Test fixture only - never imported by the product."""


class ObjectRef:
    pass


def remote(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda f: f


def is_initialized():
    return False


def _unavailable(*a, **k):
    raise RuntimeError("ray is not available (stub)")


init = get = put = _unavailable


def kill(x):
    pass


def available_resources():
    return {}
