"""TEST-ONLY paramz.transformations stand-in: Logexp / Logistic / __fixed__ (see paramz/__init__.py)."""
import numpy as np

_LIM = 36.0
__fixed__ = "fixed"


class Transformation(object):
    domain = None

    def f(self, x):
        raise NotImplementedError

    def finv(self, f):
        raise NotImplementedError

    def gradfactor(self, f, df):
        raise NotImplementedError

    def log_jacobian(self, f):
        raise NotImplementedError


class Logexp(Transformation):
    domain = "positive"

    def f(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > _LIM, x, np.log1p(np.exp(np.clip(x, -np.inf, _LIM))))

    def finv(self, f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM, f, np.log(np.expm1(f)))

    def gradfactor(self, f, df):
        f = np.asarray(f, dtype=np.float64)
        return df * np.where(f > _LIM, 1.0, -np.expm1(-f))

    def __str__(self):
        return "+ve"


class Logistic(Transformation):
    domain = "bounded"

    def __init__(self, lower=0.0, upper=1.0):
        self.lower, self.upper = float(lower), float(upper)
        self.difference = self.upper - self.lower

    def f(self, x):
        return self.lower + self.difference / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))

    def finv(self, f):
        return np.log(np.clip(f - self.lower, 1e-10, np.inf) / np.clip(self.upper - f, 1e-10, np.inf))


class Exponent(Transformation):
    domain = "positive"

    def f(self, x):
        return np.exp(x)

    def finv(self, f):
        return np.log(f)


class NegativeLogexp(Transformation):
    domain = "negative"


class LogexpNeg(Transformation):
    domain = "positive"


class Square(Transformation):
    domain = "positive"
