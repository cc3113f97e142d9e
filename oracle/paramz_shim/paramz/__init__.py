"""TEST-ONLY stand-in for the third-party `paramz` package (GPy setup.py:146, not installed in this image, no network).

Purpose: let the UNMODIFIED reference modules under /root/reference/GPy (kernels, exact inference, Gaussian
likelihood, linalg) be imported and run in the build container so that tests/golden/make_golden.py can take its numbers
from the reference itself (oracle/ref_gpy.py). It is observer-free: nothing is cached (`Cache_this` is the identity) and
`parameters_changed()` is called explicitly by the caller. It implements none of paramz's optimisation machinery and is
never imported by the product (gpy_b200/).
"""
import numpy as np

from .transformations import Logexp, __fixed__  # noqa: F401
from .parameterized import Parameterized, ParametersChangedMeta  # noqa: F401
from .core.parameter_core import Parameterizable


class ObsAr(np.ndarray):
    """paramz.ObsAr: an ndarray (observability dropped)."""

    def __new__(cls, input_array, *a, **kw):
        return np.atleast_1d(np.require(input_array, dtype=np.float64, requirements=["W", "C"])).view(cls)

    @property
    def values(self):
        return self.view(np.ndarray)


class Param(np.ndarray, Parameterizable):
    """paramz.Param: float ndarray with a name, a `.gradient` of the same shape and a default constraint."""

    def __new__(cls, name, input_array, default_constraint=None, *a, **kw):
        obj = np.atleast_1d(np.array(input_array, dtype=np.float64)).view(cls)
        return obj

    def __init__(self, name, input_array, default_constraint=None, *a, **kw):
        Parameterizable.__init__(self, name=name)
        self._gradient_ = np.zeros(self.shape)
        self.default_constraint = default_constraint

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self._gradient_ = None
        self.default_constraint = getattr(obj, "default_constraint", None)

    def __array_wrap__(self, out_arr, context=None, return_scalar=False):
        # arithmetic on parameters yields plain arrays, as with paramz
        out = np.asarray(out_arr)
        return out[()] if out.ndim == 0 else out

    @property
    def gradient(self):
        return self._gradient_

    @gradient.setter
    def gradient(self, val):
        g = np.zeros(self.shape)
        g[...] = val
        self._gradient_ = g

    @property
    def values(self):
        return self.view(np.ndarray)

    def flattened_parameters(self):
        return [self]


class Model(Parameterized):
    """paramz.Model surface used by GPy/core/model.py (import-time only here)."""

    def __init__(self, name):
        super(Model, self).__init__(name)


def load(*a, **kw):
    raise NotImplementedError("paramz shim: load")
