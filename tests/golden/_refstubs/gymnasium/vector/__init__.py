class VectorEnv:
    def __init__(self, *args, **kwargs):
        pass

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class SyncVectorEnv(VectorEnv):
    pass


class AsyncVectorEnv(VectorEnv):
    pass
