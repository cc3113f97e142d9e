"""TEST-ONLY stand-in for the third-party `paramz` package (GPy setup.py:146, not installed in this image, no network).

Purpose: let the UNMODIFIED reference modules under /root/reference/GPy (kernels, exact inference, Gaussian
likelihood, linalg, and since round 2 core/gp.py, core/model.py, models/gp_regression.py) be imported and run in the build
container so that tests/golden/make_golden.py can take its numbers from the reference itself (oracle/ref_gpy.py) and the
plugin can be driven by the reference's own GP / GPRegression objects. Nothing is cached (`Cache_this` is the identity).
What it does provide of paramz's runtime (first six rows), in the simplest form that keeps the
contract: a write to a linked Param re-runs the root's `parameters_changed()` (eager observer), `update_model(False/True)`,
`fix()/unfix()`, the Logexp-transformed `optimizer_array`, and `Model.optimize` = scipy L-BFGS-B on it with paramz's
failure counting. It is never imported by the product (gpy_b200/).
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
        self._fixed_ = False

    # ---- observer contract: a write to a linked parameter re-evaluates the model it belongs to ------------------------
    def __setitem__(self, idx, val):
        np.ndarray.__setitem__(self, idx, val)
        self._notify_root()

    def _notify_root(self):
        root = self
        while getattr(root, "_parent_", None) is not None:
            root = root._parent_
        if root is not self and hasattr(root, "_trigger_params_changed"):
            root._trigger_params_changed()

    def fix(self, value=None, warning=True):
        if value is not None:
            self[:] = value
        self._fixed_ = True

    constrain_fixed = fix

    def unfix(self):
        self._fixed_ = False

    unconstrain_fixed = unfix

    @property
    def is_fixed(self):
        return bool(getattr(self, "_fixed_", False))

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self._gradient_ = None
        self.default_constraint = getattr(obj, "default_constraint", None)
        self._parent_ = None          # views / copies are not linked anywhere
        self._fixed_ = False

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
    """paramz.Model: what GPy/core/model.py and GPy/core/gp.py:663-684 call — optimize / checkgrad over the transformed
    `optimizer_array`, objective = objective_function() of the subclass (GPy/core/model.py:97-128)."""

    _allowed_failures = 10

    def __init__(self, name):
        super(Model, self).__init__(name)
        self.optimization_runs = []
        self._fail_count = 0
        self.obj_grads = None

    def _transform_gradients(self, g):
        """chain rule through the constraints, fixed entries dropped (paramz Parameterizable._transform_gradients)"""
        out, i = [], 0
        for p in self.flattened_parameters():
            gp = np.asarray(g[i:i + p.size], dtype=np.float64)
            i += p.size
            if p.is_fixed:
                continue
            c = p.default_constraint
            out.append(gp if c is None else c.gradfactor(np.asarray(p).reshape(-1), gp))
        return np.concatenate(out) if out else np.zeros(0)

    def _objective_grads(self, x):
        try:
            self.optimizer_array = x
            obj_f = self.objective_function()
            self.obj_grads = self._transform_gradients(self.objective_function_gradients())
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= self._allowed_failures:
                raise
            self._fail_count += 1
            obj_f = np.inf
            self.obj_grads = np.clip(self.obj_grads if self.obj_grads is not None else np.zeros_like(x), -1e10, 1e10)
        return obj_f, self.obj_grads

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, ipython_notebook=True,
                 clear_after_finish=False, **kwargs):
        """paramz Model.optimize with the default optimizer 'lbfgsb' (paramz.optimization.opt_lbfgsb ->
        scipy.optimize.fmin_l_bfgs_b(f_fp, x_init, maxfun=max_iters, maxiter=max_iters, pgtol=gtol, factr=bfgs_factor))."""
        from scipy.optimize import fmin_l_bfgs_b
        if optimizer not in (None, "lbfgsb", "lbfgs", "bfgs", "lbfgsb"):
            raise NotImplementedError("paramz shim: only the default L-BFGS-B optimizer")
        if self.size == 0 or all(p.is_fixed for p in self.flattened_parameters()):
            return None
        opt = {}
        if "gtol" in kwargs and kwargs["gtol"] is not None:
            opt["pgtol"] = kwargs["gtol"]
        if "bfgs_factor" in kwargs and kwargs["bfgs_factor"] is not None:
            opt["factr"] = kwargs["bfgs_factor"]
        x0 = self.optimizer_array.copy() if start is None else np.asarray(start, dtype=np.float64)
        self._n_evals = 0

        def f_fp(x):
            self._n_evals += 1
            f, g = self._objective_grads(x)
            if messages:
                print("eval %4d  objective %.10f" % (self._n_evals, f))
            return f, g

        x, f, d = fmin_l_bfgs_b(f_fp, x0, maxfun=max_iters, maxiter=max_iters, **opt)
        self.optimizer_array = x
        run = type("OptimizationRun", (object,), {})()
        run.x_opt, run.f_opt, run.funct_eval, run.status, run.info = x, f, d["funcalls"], d.get("task"), d
        self.optimization_runs.append(run)
        return run

    def checkgrad(self, verbose=False, step=1e-6, tolerance=1e-3):
        """central finite differences of the objective against the transformed analytic gradient"""
        x = self.optimizer_array.copy()
        _, g = self._objective_grads(x)
        g = g.copy()
        num = np.zeros_like(x)
        for i in range(x.size):
            xp, xm = x.copy(), x.copy()
            xp[i] += step
            xm[i] -= step
            num[i] = (self._objective_grads(xp)[0] - self._objective_grads(xm)[0]) / (2 * step)
        self.optimizer_array = x
        ratio = np.where(num != 0, g / np.where(num == 0, 1, num), 1.0)
        if verbose:
            print("analytic", g, "numeric", num)
        return bool(np.all(np.abs(1.0 - ratio) < tolerance) or np.allclose(g, num, atol=tolerance * 1e-2))


def load(*a, **kw):
    raise NotImplementedError("paramz shim: load")
