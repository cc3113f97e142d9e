"""TEST-ONLY paramz.caching stand-in: no caching at all (correctness never depended on it)."""


class Cacher(object):
    def __init__(self, operation, limit=3, ignore_args=(), force_kwargs=()):
        self.operation = operation

    def __call__(self, *a, **kw):
        return self.operation(*a, **kw)


class Cache_this(object):
    def __init__(self, limit=5, ignore_args=(), force_kwargs=()):
        pass

    def __call__(self, f):
        return f
