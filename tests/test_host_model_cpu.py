"""Host-side model logic of the stand-alone mirror (gpy_b200.GPRegression) on CPU: the C ABI is replaced by a test double
that answers from the oracle, so what is tested is the HOST logic only — observer propagation of parameter writes
(paramz contract: every write re-evaluates), the exact-content data key, posterior handles that must not outlive
their evaluation, `predict(kern=other)`, and the error mapping of the ctypes layer."""
import numpy as np
import pytest

from oracle import gpy_oracle as o


class FakeEngine(object):
    """test double of gpy_b200._ffi.Engine: same methods, numbers from the oracle."""

    def __init__(self, device=0):
        self.calls = []
        self.eval_serial = 0

    def set_data(self, X, Y):
        self.X, self.Y = np.array(X), np.array(Y)
        self.eval_serial += 1
        self.calls.append("set_data")

    def exact_eval(self, kind, ARD, variance, lengthscale, noise, jitter=1e-8, max_tries=5):
        self.calls.append("exact_eval")
        self.eval_serial += 1
        self.theta = (kind, ARD, variance, np.array(lengthscale, copy=True))
        lml, g, self.res = o.eval_lml_grad(self.X, self.Y, kind, ARD, variance, lengthscale, noise)
        return lml, g, 0.0

    def exact_eval_multi(self, parts, noise, jitter=1e-8, max_tries=5):
        self.calls.append("exact_eval_multi")
        self.eval_serial += 1
        self.parts = [dict(kind=k, ARD=a, term=t, dims=d, variance=v, lengthscale=l) for (k, a, t, d, v, l) in parts]
        lml, g, self.res = o.composite_eval_lml_grad(self.X, self.Y, self.parts, noise)
        self.theta = None
        return lml, g, 0.0

    def get(self, which):
        return {"L": self.res["L"], "alpha": self.res["alpha"], "Kinv": self.res["Wi"], "dL_dK": self.res["dL_dK"],
                "K": self.res["K"]}[which]

    def predict(self, Xnew, full_cov=False):
        if self.theta is None:                      # composite kernel: posterior.py:276-295 with the composite K
            kp = self.res["kparts"]
            Kx = o.composite_K(kp, self.X, Xnew)
            mu = Kx.T @ self.res["alpha"]
            tmp = o.dtrtrs(self.res["L"], Kx, lower=1)[0]
            if full_cov:
                return mu, o.composite_K(kp, Xnew) - tmp.T @ tmp
            return mu, (o.composite_Kdiag(kp, Xnew) - np.square(tmp).sum(0))[:, None]
        kind, ARD, var, ls = self.theta
        k = o.StationaryOracle(kind, self.X.shape[1], var, ls, ARD)
        return o.raw_predict(k, self.X, self.res["L"], self.res["alpha"], Xnew, full_cov)

    def total_launches(self):
        return 0


def _model(N=60, D=3, seed=2):
    import gpy_b200
    X, Y = o.synthetic(N, D, seed)
    eng = FakeEngine()
    m = gpy_b200.GPRegression(X, Y, gpy_b200.RBF(D, variance=1.3, lengthscale=[1.1, 1.6, 2.2], ARD=True), noise_var=0.05,
                              engine=eng)
    return m, eng, X, Y


def test_parameter_writes_re_evaluate_like_paramz():
    m, eng, X, Y = _model()
    assert eng.calls == ["set_data", "exact_eval"]           # one evaluation after construction, none during it
    l0 = m.log_likelihood()
    m.kern.variance[0] = 2.0                                   # a leaf write -> GP.parameters_changed()
    assert eng.calls[-1] == "exact_eval" and len(eng.calls) == 3
    ref, g, _ = o.eval_lml_grad(X, Y, "rbf", True, 2.0, np.array([1.1, 1.6, 2.2]), 0.05)
    assert m.log_likelihood() != l0 and abs(m.log_likelihood() - ref) < 1e-9
    np.testing.assert_allclose(m.gradient, g, rtol=1e-9)
    m.likelihood.variance.set(0.2)
    ref, g, _ = o.eval_lml_grad(X, Y, "rbf", True, 2.0, np.array([1.1, 1.6, 2.2]), 0.2)
    assert abs(m.log_likelihood() - ref) < 1e-9
    # update_model(False) defers, update_model(True) evaluates once
    n = len(eng.calls)
    m.update_model(False)
    m.kern.lengthscale[1] = 3.0
    m.kern.variance[0] = 0.7
    assert len(eng.calls) == n
    m.update_model(True)
    assert len(eng.calls) == n + 1
    ref, g, _ = o.eval_lml_grad(X, Y, "rbf", True, 0.7, np.array([1.1, 3.0, 2.2]), 0.2)
    assert abs(m.log_likelihood() - ref) < 1e-9


def test_data_key_sees_row_permutations_and_in_place_edits():
    from gpy_b200.inference import _DataKey
    rng = np.random.default_rng(0)
    X, Y = rng.standard_normal((1000, 4)), rng.standard_normal((1000, 1))
    k = _DataKey()
    assert not k.matches(X, Y)
    k.remember(X, Y)
    assert k.matches(X, Y) and k.matches(X.copy(), Y.copy())
    for trial in range(50):                                    # a moment / strided-sample fingerprint misses most of these
        i, j = rng.choice(1000, 2, replace=False)
        Yp = Y.copy(); Yp[[i, j]] = Yp[[j, i]]
        Xp = X.copy(); Xp[[i, j]] = Xp[[j, i]]
        assert not k.matches(X, Yp) and not k.matches(Xp, Y)
    X[3, 1] += 1e-9                                            # in-place edit of the caller's array: the key holds a copy
    assert not k.matches(X, Y)
    assert not k.matches(X[:999], Y[:999])


def test_shuffled_labels_are_uploaded_again():
    m, eng, X, Y = _model()
    Yp = Y.copy(); Yp[[0, 1]] = Yp[[1, 0]]
    post, lml, _ = m.inference_method.inference(m.kern, X, m.likelihood, Yp)
    assert eng.calls[-2:] == ["set_data", "exact_eval"]
    ref, _, _ = o.eval_lml_grad(X, Yp, "rbf", True, 1.3, np.array([1.1, 1.6, 2.2]), 0.05)
    assert abs(lml - ref) < 1e-9


def test_posterior_handle_does_not_outlive_its_evaluation():
    m, eng, X, Y = _model()
    old = m.posterior
    alpha_old = old.woodbury_vector.copy()                     # fetched while fresh: stays cached and valid
    m.kern.variance[0] = 2.5
    np.testing.assert_array_equal(old.woodbury_vector, alpha_old)
    with pytest.raises(RuntimeError):
        old.woodbury_chol                                      # not fetched before the new evaluation -> refuses
    assert m.posterior is not old
    assert m.posterior.woodbury_chol.shape == (60, 60)


def test_predict_with_another_kernel_uses_that_kernel():
    import gpy_b200
    m, eng, X, Y = _model()
    Xn = np.random.default_rng(1).uniform(-2, 2, (7, 3))
    mu, var = m.predict(Xn)
    k = o.StationaryOracle("rbf", 3, 1.3, np.array([1.1, 1.6, 2.2]), True)
    mu0, var0 = o.predict(k, X, eng.res["L"], eng.res["alpha"], Xn, 0.05)
    np.testing.assert_allclose(mu, mu0, rtol=1e-10)
    np.testing.assert_allclose(var, var0, rtol=1e-10)

    class HostRBF(gpy_b200.RBF):
        """answers K / Kdiag from the oracle (no device here)"""
        def K(self, X, X2=None):
            return o.StationaryOracle("rbf", self.input_dim, float(self.variance[0]), self.lengthscale.values, self.ARD).K(X, X2)
        def Kdiag(self, X):
            return np.full(X.shape[0], float(self.variance[0]))

    other = HostRBF(3, variance=0.4, lengthscale=[2.0, 2.0, 2.0], ARD=True)
    mu2, var2 = m.predict(Xn, kern=other, include_likelihood=False)
    k2 = o.StationaryOracle("rbf", 3, 0.4, np.array([2.0, 2.0, 2.0]), True)
    mu20, var20 = o.raw_predict(k2, X, eng.res["L"], eng.res["alpha"], Xn)      # posterior.py:276-295 with the GIVEN kernel
    np.testing.assert_allclose(mu2, mu20, rtol=1e-10)
    np.testing.assert_allclose(var2, var20, rtol=1e-10)
    assert not np.allclose(mu2, mu)


def test_argument_domain_errors_are_value_errors_and_optimize_survives():
    from gpy_b200 import _ffi
    assert issubclass(_ffi.GpxArgumentError, ValueError) and issubclass(_ffi.GpxArgumentError, _ffi.GpxError)
    with pytest.raises(ValueError):
        _ffi.check(-2, "gpx_exact_eval")
    with pytest.raises(np.linalg.LinAlgError):
        _ffi.check(3, "gpx_exact_eval")
    with pytest.raises(_ffi.GpxError):
        _ffi.check(-1, "gpx_exact_eval")
    m, eng, X, Y = _model(N=40)
    real = eng.exact_eval
    state = {"n": 0}

    def flaky(*a, **k):                                        # the 4th evaluation of the run fails in the argument check
        state["n"] += 1
        if state["n"] == 4:
            raise _ffi.GpxArgumentError("gpx_exact_eval: lengthscale must be positive")
        return real(*a, **k)

    eng.exact_eval = flaky
    f0 = m.objective_function()
    d = m.optimize(max_iters=15)                               # must not raise; the failed iterate counts as f = inf
    assert state["n"] > 4 and np.isfinite(d["f"]) and d["f"] < f0


def test_sum_and_product_kernels_take_the_fused_composite_path():
    """Add / Prod / White / Bias (add.py:60-99, prod.py:59-68,377-396, static.py:63-185): a sum of products of leaves is ONE
    engine call per evaluation (no N^2 host glue), gradients land on the leaves in paramz's link order, predict works."""
    import gpy_b200
    X, Y = o.synthetic(70, 4, 3)
    eng = FakeEngine()
    k = gpy_b200.Add([gpy_b200.Prod([gpy_b200.RBF(2, variance=1.2, lengthscale=[1.0, 2.0], ARD=True, active_dims=[0, 1]),
                                     gpy_b200.Matern32(2, variance=0.8, lengthscale=1.5, active_dims=[2, 3])]),
                      gpy_b200.Matern52(4, variance=0.5, lengthscale=2.0), gpy_b200.White(4, variance=0.05),
                      gpy_b200.Bias(4, variance=0.3)])
    m = gpy_b200.GPRegression(X, Y, k, noise_var=0.1, engine=eng)
    assert eng.calls == ["set_data", "exact_eval_multi"]
    parts = [dict(kind="rbf", term=0, dims=[0, 1], variance=1.2, lengthscale=np.array([1.0, 2.0]), ARD=True),
             dict(kind="matern32", term=0, dims=[2, 3], variance=0.8, lengthscale=1.5, ARD=False),
             dict(kind="matern52", term=1, dims=[0, 1, 2, 3], variance=0.5, lengthscale=2.0, ARD=False),
             dict(kind="white", term=2, dims=None, variance=0.05), dict(kind="bias", term=3, dims=None, variance=0.3)]
    lml0, g0, res = o.composite_eval_lml_grad(X, Y, parts, 0.1)
    assert abs(m.log_likelihood() - lml0) < 1e-9
    np.testing.assert_allclose(m.gradient, g0, rtol=1e-9)
    assert m.checkgrad()
    Xn = np.random.default_rng(2).uniform(-2, 2, (5, 4))
    mu, var = m.predict(Xn)
    Kx = o.composite_K(res["kparts"], X, Xn)
    np.testing.assert_allclose(mu, Kx.T @ res["alpha"], rtol=1e-10)
    # a kernel the flattening does not cover (product containing a sum) goes through the generic path, same numbers
    k2 = gpy_b200.Prod([gpy_b200.Add([gpy_b200.RBF(4), gpy_b200.Bias(4)]), gpy_b200.Matern32(4)])
    from gpy_b200.kern import flatten_parts
    assert flatten_parts(k2) is None and flatten_parts(k) is not None and len(flatten_parts(k)) == 5


def test_normalizer_and_pickle_round_trip():
    """normalizer=True -> Standardize (gp.py:49-66,355-363; normalizer.py:85-113): the engine sees the standardised Y, the
    predictions come back in the units of Y. Pickling drops the device handle (rbf.py:313-318 precedent) and the copy
    evaluates again after loading."""
    import pickle
    import gpy_b200
    X, Y = o.synthetic(50, 2, 4)
    Y = 3.0 + 5.0 * Y
    eng = FakeEngine()
    m = gpy_b200.GPRegression(X, Y, gpy_b200.RBF(2, lengthscale=1.4), noise_var=0.1, normalizer=True, engine=eng)
    Yn = (Y - Y.mean(0)) / Y.std(0)
    np.testing.assert_allclose(eng.Y, Yn, rtol=1e-14)
    lml0, g0, res = o.eval_lml_grad(X, Yn, "rbf", False, 1.0, 1.4, 0.1)
    assert abs(m.log_likelihood() - lml0) < 1e-9
    Xn = np.random.default_rng(0).uniform(-2, 2, (4, 2))
    mu, var = m.predict(Xn)
    k = o.StationaryOracle("rbf", 2, 1.0, 1.4, False)
    mu0, var0 = o.predict(k, X, res["L"], res["alpha"], Xn, 0.1)
    np.testing.assert_allclose(mu, mu0 * Y.std(0) + Y.mean(0), rtol=1e-10)
    np.testing.assert_allclose(var, var0 * Y.std(0) ** 2, rtol=1e-10)
    blob = pickle.dumps(m)
    m2 = pickle.loads(blob)
    assert m2.inference_method._engine is None and m2.posterior is None
    m2.inference_method._engine = FakeEngine()           # stands for the lazily re-created device context
    m2.parameters_changed()
    assert abs(m2.log_likelihood() - lml0) < 1e-9 and m2.inference_method._engine.calls == ["set_data", "exact_eval"]


class FakeSparseEngine(object):
    """test double for the sparse entry points of gpy_b200._ffi.Engine: numbers from the oracle's VarDTC."""

    def __init__(self):
        self.calls, self.sparse_serial = [], 0

    def sparse_set_data(self, X, Y):
        self.X, self.Y = np.array(X), np.array(Y)
        self.calls.append("sparse_set_data")

    def sparse_eval(self, kind, ARD, variance, lengthscale, Z, noise_variance):
        self.calls.append("sparse_eval")
        self.sparse_serial += 1
        lml, g, Zg, self.res = o.sparse_eval(self.X, self.Y, Z, kind, ARD, variance, lengthscale, noise_variance)
        return lml, g, Zg

    def sparse_eval_het(self, kind, ARD, variance, lengthscale, Z, noise_variances):
        self.calls.append("sparse_eval_het")
        self.sparse_serial += 1
        lml, g, Zg, self.res = o.sparse_eval(self.X, self.Y, Z, kind, ARD, variance, lengthscale,
                                             np.asarray(noise_variances).reshape(-1))
        nk = g.size - self.X.shape[0] * self.Y.shape[1]
        return lml, g[:nk], Zg, g[nk:].reshape(self.X.shape[0], self.Y.shape[1])

    def sparse_get(self, what):
        return {"woodbury_vector": self.res["woodbury_vector"], "woodbury_inv": self.res["woodbury_inv"],
                "Kmm": self.res["Kmm"], "Lm": self.res["Lm"]}[what]


def test_sparse_model_selects_the_heteroscedastic_branch_from_the_likelihood():
    """core/sparse_gp.py:76-119 with a HeteroscedasticGaussian likelihood: VarDTC.inference must take the het_noise route
    (precision.size > 1, var_dtc.py:82-84), hand dL_dR to the likelihood by output_index (gaussian.py:358-359), and the
    model's gradient must be the derivative of its own objective (finite differences in the optimizer space)."""
    import gpy_b200
    rng = np.random.default_rng(5)
    N, M, D = 90, 12, 2
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) + 0.1 * rng.standard_normal((N, 1))
    Z = X[rng.permutation(N)[:M]] + 0.01
    meta = {"output_index": np.arange(N)[:, None]}
    lik = gpy_b200.HeteroscedasticGaussian(meta)
    nv = rng.uniform(0.02, 0.3, N)
    lik.variance.values[...] = nv
    eng = FakeSparseEngine()

    class HostRBF(gpy_b200.RBF):
        """answers K / Kdiag from the oracle (no device here)"""
        def K(self, X, X2=None):
            return o.StationaryOracle("rbf", self.input_dim, float(self.variance[0]), self.lengthscale.values, self.ARD).K(X, X2)
        def Kdiag(self, X):
            return np.full(X.shape[0], float(self.variance[0]))

    m = gpy_b200.SparseGPRegression(X, Y, kernel=HostRBF(D, variance=1.2, lengthscale=[1.1, 1.8], ARD=True), Z=Z,
                                    engine=eng, likelihood=lik, Y_metadata=meta)
    assert "sparse_eval_het" in eng.calls and "sparse_eval" not in eng.calls
    lml0, g0, Zg0, _ = o.sparse_eval(X, Y, Z, "rbf", True, 1.2, np.array([1.1, 1.8]), nv)
    assert abs(m.log_likelihood() - lml0) <= 1e-10 * abs(lml0)
    g = np.concatenate([m.kern.variance.gradient, m.kern.lengthscale.gradient, m.likelihood.variance.gradient])
    np.testing.assert_allclose(g, g0, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(m.Z.gradient, Zg0, rtol=1e-12, atol=1e-12)
    assert m.optimizer_array.size == M * D + 1 + 2 + N
    assert m.checkgrad(step=1e-5)
    mu, var = m.predict(X[:4], include_likelihood=False)
    assert mu.shape == (4, 1) and var.shape == (4, 1)
    with pytest.raises(ValueError):
        m.predict(X[:4])                                           # no noise model away from the training points
    mu2, var2 = m.predict(X[:4], Y_metadata={"output_index": np.arange(4)[:, None]})
    np.testing.assert_allclose(var2, var + nv[:4, None], rtol=1e-12)
    # a scalar-noise model on the same engine class takes the homoscedastic entry point
    eng2 = FakeSparseEngine()
    gpy_b200.SparseGPRegression(X, Y, kernel=gpy_b200.RBF(D), Z=Z, engine=eng2)
    assert "sparse_eval" in eng2.calls and "sparse_eval_het" not in eng2.calls
