class VectorEnv:
    pass


class SyncVectorEnv(VectorEnv):
    pass


class AsyncVectorEnv(VectorEnv):
    pass
