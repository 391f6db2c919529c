from . import linen, struct  # noqa: F401


class optim:
    class Optimizer:
        pass
