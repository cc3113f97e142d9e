"""TEST-ONLY paramz.parameterized stand-in (see paramz/__init__.py): link order, param_array / gradient views, and the
metaclass that calls parameters_changed() once construction has finished."""
import numpy as np

from .core.parameter_core import Parameterizable


class ParametersChangedMeta(type):
    def __call__(cls, *args, **kw):
        inst = super(ParametersChangedMeta, cls).__call__(*args, **kw)
        inst._in_init_ = False
        try:
            inst.parameters_changed()
        except NotImplementedError:
            pass
        return inst


class Parameterized(Parameterizable, metaclass=ParametersChangedMeta):
    def __init__(self, name=None, parameters=(), *a, **kw):
        super(Parameterized, self).__init__(name=name)
        self.parameters = []
        self._in_init_ = True

    def link_parameter(self, p, index=None):
        self.parameters.append(p) if index is None else self.parameters.insert(index, p)
        p._parent_ = self

    def link_parameters(self, *ps):
        for p in ps:
            self.link_parameter(p)

    def unlink_parameter(self, p):
        self.parameters.remove(p)

    def parameters_changed(self):
        pass

    def flattened_parameters(self):
        out = []
        for p in self.parameters:
            out.extend(p.flattened_parameters())
        return out

    @property
    def param_array(self):
        ps = self.flattened_parameters()
        return np.concatenate([np.asarray(p).reshape(-1) for p in ps]) if ps else np.zeros(0)

    @property
    def gradient(self):
        ps = self.flattened_parameters()
        return np.concatenate([np.asarray(p.gradient, dtype=np.float64).reshape(-1) for p in ps]) if ps else np.zeros(0)

    @property
    def size(self):
        return sum(p.size for p in self.flattened_parameters())
