"""TEST-ONLY paramz.core.pickleable stand-in."""


class Pickleable(object):
    def __init__(self, *a, **kw):
        pass
