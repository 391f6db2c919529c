import dataclasses


def field(pytree_node=True, **kw):
    return dataclasses.field(**kw)


def dataclass(cls):
    return dataclasses.dataclass(cls)
