"""Minimal parameter plumbing for the host-side mirror of GPy's plugin interface.

GPy gets `Param`, `Parameterized`, `Logexp` and the optimizer loop from the third-party `paramz` package
(GPy setup.py:146), which is not part of the hot path and is not re-implemented here. What the hot path touches
is tiny: a float array that carries a `.gradient` of the same shape (`self.variance.gradient = ...`,
GPy/kern/src/stationary.py:199-213, GPy/likelihoods/gaussian.py:72-73), the `Logexp` positivity transform applied to
all three hyper-parameters (stationary.py:78-79, gaussian.py:43), and concatenated `param_array` / `gradient` views
in link order. This file provides exactly that so the mirror classes can be driven stand-alone; when the real GPy
is importable the classes in gpy_b200/gpy_plugin.py subclass GPy's own instead.
"""
import numpy as np

_LIM = 36.0


class Logexp(object):
    """theta = log(1 + exp(x)) — the transform paramz.transformations.Logexp applies (domain: positive)."""

    @staticmethod
    def f(x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > _LIM, x, np.log1p(np.exp(np.clip(x, -np.inf, _LIM))))

    @staticmethod
    def finv(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM, f, np.log(np.expm1(np.clip(f, 1e-300, np.inf))))

    @staticmethod
    def gradfactor(f):
        """d theta / d x evaluated at theta = f."""
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM, 1.0, -np.expm1(-f))


class Param(object):
    """A named positive parameter vector with a `.gradient` of the same shape (cf. paramz.Param)."""

    def __init__(self, name, value, transform=Logexp):
        self.name = name
        self.values = np.atleast_1d(np.array(value, dtype=np.float64)).copy()
        self.gradient = np.zeros_like(self.values)
        self.transform = transform
        self.is_fixed = False
        self._observers = []

    # array-ish behaviour ------------------------------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __len__(self):
        return self.values.size

    def __getitem__(self, idx):
        return self.values[idx]

    def __setitem__(self, idx, v):
        self.values[idx] = v
        self._notify()

    def __float__(self):
        return float(self.values.reshape(-1)[0]) if self.values.size == 1 else float(self.values)

    def _binary(self, other, op):
        return op(self.values, np.asarray(other))

    def __add__(self, o): return self.values + np.asarray(o)
    __radd__ = __add__
    def __mul__(self, o): return self.values * np.asarray(o)
    __rmul__ = __mul__
    def __sub__(self, o): return self.values - np.asarray(o)
    def __rsub__(self, o): return np.asarray(o) - self.values
    def __truediv__(self, o): return self.values / np.asarray(o)
    def __rtruediv__(self, o): return np.asarray(o) / self.values
    def __pow__(self, o): return self.values ** o
    def __neg__(self): return -self.values

    @property
    def size(self):
        return self.values.size

    @property
    def shape(self):
        return self.values.shape

    def fix(self):
        self.is_fixed = True

    def unfix(self):
        self.is_fixed = False

    def set(self, v):
        self.values[...] = v
        self._notify()

    def _notify(self):
        for cb in self._observers:
            cb()

    def __repr__(self):
        return "Param(%s=%s)" % (self.name, np.array2string(self.values, precision=6))


class Parameterized(object):
    """Ordered container of Params and sub-containers (cf. paramz.Parameterized.link_parameter).

    Writes to a linked leaf (`m.kern.variance[0] = 2.`, `Param.set`) travel up the link chain and end in the root's
    `parameters_changed()` — the observer contract of paramz (`Parameterizable._notify_parent_change` /
    `_trigger_params_changed`), so that `log_likelihood()`, `gradient` and `predict()` never describe an older theta.
    The root gates the re-evaluation on `update_model_flag` (paramz `update_model(False)`)."""

    _parent = None

    def __init__(self, name):
        self.name = name
        self._links = []

    def link_parameter(self, p):
        self._links.append(p)
        if isinstance(p, Param):
            p._observers.append(self._child_changed)
        else:
            p._parent = self

    def _child_changed(self):
        root = self
        while getattr(root, "_parent", None) is not None:
            root = root._parent
        if getattr(root, "update_model_flag", True) and getattr(root, "_initialised", True):
            root.parameters_changed()

    def parameters_changed(self):
        pass

    link_parameters = lambda self, *ps: [self.link_parameter(p) for p in ps]

    def flattened_parameters(self):
        out = []
        for p in self._links:
            if isinstance(p, Param):
                out.append(p)
            else:
                out.extend(p.flattened_parameters())
        return out

    @property
    def param_array(self):
        ps = self.flattened_parameters()
        return np.concatenate([p.values.reshape(-1) for p in ps]) if ps else np.zeros(0)

    @property
    def gradient(self):
        ps = self.flattened_parameters()
        return np.concatenate([np.asarray(p.gradient, dtype=np.float64).reshape(-1) * np.ones(p.size) for p in ps]) \
            if ps else np.zeros(0)

    def parameter_names(self):
        return ["%s.%s" % (self.name, p.name) for p in self.flattened_parameters()]
