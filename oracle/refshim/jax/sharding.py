"""refshim jax.sharding: names only (the sharded loops of the reference are not executed through the shim)."""


class Mesh:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class PartitionSpec(tuple):
    def __new__(cls, *a):
        return super().__new__(cls, a)


class NamedSharding:
    def __init__(self, *a, **k):
        pass
