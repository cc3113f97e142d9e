"""TEST-ONLY paramz.optimization stand-in (import-time only)."""


class Optimizer(object):
    pass
