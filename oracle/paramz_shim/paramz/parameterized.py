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
        self._update_on_ = True

    # ---- observer contract (eager): parameter writes below this node end in the ROOT's parameters_changed() ----------
    def _trigger_params_changed(self):
        if getattr(self, "_in_init_", False) or not getattr(self, "_update_on_", True):
            return
        self.parameters_changed()

    def update_model(self, updates=None):
        """paramz Parameterizable.update_model: False defers re-evaluation while several things are written"""
        if updates is None:
            return self._update_on_
        was = self._update_on_
        self._update_on_ = bool(updates)
        if updates and not was:
            self._trigger_params_changed()

    @property
    def is_fixed(self):
        ps = self.flattened_parameters()
        return bool(ps) and all(p.is_fixed for p in ps)

    def fix(self, *a, **kw):
        for p in self.flattened_parameters():
            p._fixed_ = True

    def unfix(self):
        for p in self.flattened_parameters():
            p._fixed_ = False

    @property
    def optimizer_array(self):
        """free parameters in the optimizer's (transformed) space, link order"""
        out = []
        for p in self.flattened_parameters():
            if p.is_fixed:
                continue
            v = np.asarray(p, dtype=np.float64).reshape(-1)
            c = p.default_constraint
            out.append(v if c is None else c.finv(v))
        return np.concatenate(out) if out else np.zeros(0)

    @optimizer_array.setter
    def optimizer_array(self, x):
        x = np.asarray(x, dtype=np.float64)
        i = 0
        for p in self.flattened_parameters():
            if p.is_fixed:
                continue
            v = x[i:i + p.size]
            i += p.size
            c = p.default_constraint
            np.ndarray.__setitem__(p, Ellipsis, (v if c is None else c.f(v)).reshape(p.shape))   # silent write ...
        self._trigger_params_changed()                                                            # ... ONE evaluation

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

    @gradient.setter
    def gradient(self, val):
        """paramz keeps ONE flat gradient array and hands out views; `node.gradient = v` / `node.gradient += v` write through
        to the leaves in link order (core/sparse_gp.py:115 `self.kern.gradient += kerngrad`)"""
        val = np.asarray(val, dtype=np.float64).reshape(-1)
        i = 0
        for p in self.flattened_parameters():
            p.gradient = val[i:i + p.size].reshape(p.shape)
            i += p.size

    @property
    def size(self):
        return sum(p.size for p in self.flattened_parameters())
