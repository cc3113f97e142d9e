"""The reference's OWN model loop driving the plugin: unmodified GPy/core/gp.py (GP.__init__, parameters_changed :269-282,
optimize :663-684, predict), GPy/core/model.py (objective :97-128) and GPy/models/gp_regression.py:29-36, executed from
/root/reference through oracle/ref_gpy.load_models() (paramz supplied by the test-only stand-in):

    m = GPRegression(X, Y, B.RBF(D, ARD=True)); m.inference_method = B.ExactGaussianInference(); m.optimize()

exactly the hook the reference's own tests use (GPy/testing/fitc.py:31). The CPU test replaces the C ABI by a test double
answering from the oracle (what is tested is that the stock GP / paramz loop runs UNCHANGED on the plugin pair and lands
where the stock pair lands); the GPU test runs the same thing on the device and is skipped where the reference tree does
not exist (the GPU box)."""
import os

import numpy as np
import pytest

from oracle import gpy_oracle as o

pytestmark = pytest.mark.skipif(not os.path.isdir(os.environ.get("GPX_REFERENCE", "/root/reference") + "/GPy"),
                                reason="reference tree not present")


def _plugin(ffi=None):
    from oracle import ref_gpy
    from gpy_b200 import gpy_plugin
    M = ref_gpy.load_models()
    G = M.G
    if ffi is None:
        B = gpy_plugin.make(G.RBF, G.Exponential, G.Matern32, G.Matern52, G.ExactGaussianInference)
    else:
        B = gpy_plugin.make(G.RBF, G.Exponential, G.Matern32, G.Matern52, G.ExactGaussianInference, ffi=ffi)
    return M, G, B


def _run_pair(M, G, B, N, D, kname, ARD, max_iters):
    X, Y = o.synthetic(N, D, seed=4)
    ls = (np.sqrt(D) * np.linspace(0.8, 1.3, D)) if ARD else float(np.sqrt(D))
    stock = M.GPRegression(X, Y, getattr(G, kname)(D, variance=1.2, lengthscale=ls, ARD=ARD), noise_var=0.1)
    plug = M.GPRegression(X, Y, getattr(B, kname)(D, variance=1.2, lengthscale=ls, ARD=ARD), noise_var=0.1)
    assert type(plug) is type(stock)                                     # the reference's own model class, untouched
    plug.inference_method = B.ExactGaussianInference()                   # GPy/testing/fitc.py:31
    plug.parameters_changed()
    assert abs(plug.log_likelihood() - stock.log_likelihood()) < 1e-8
    np.testing.assert_allclose(plug.gradient, stock.gradient, rtol=1e-6, atol=1e-9)
    # a parameter write goes through paramz's observer chain into GP.parameters_changed -> one device evaluation
    plug.kern.variance[:] = 0.9
    stock.kern.variance[:] = 0.9
    assert abs(plug.log_likelihood() - stock.log_likelihood()) < 1e-8
    assert plug.checkgrad()
    rs, rp = stock.optimize(max_iters=max_iters), plug.optimize(max_iters=max_iters)
    assert abs(plug.log_likelihood() - stock.log_likelihood()) <= 1e-5 * max(1.0, abs(stock.log_likelihood()))
    np.testing.assert_allclose(plug.param_array, stock.param_array, rtol=2e-3)
    assert plug.log_likelihood() > -plug.objective_function() - 1e-12    # objective = -LML (model.py:109)
    Xn = np.random.default_rng(0).uniform(-2, 2, (6, D))
    mu_s, var_s = stock.predict(Xn)
    mu_p, var_p = plug.predict(Xn)                                       # GP.predict -> posterior._raw_predict(kern=plugin kernel)
    np.testing.assert_allclose(mu_p, mu_s, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(var_p, var_s, rtol=1e-4, atol=1e-7)
    return plug, rp


def test_stock_gpregression_optimizes_through_the_plugin_cpu():
    from test_gpy_plugin_cpu import fake_ffi
    M, G, B = _plugin(fake_ffi())
    for (kname, ARD, D) in (("RBF", True, 3), ("Matern52", False, 4)):
        plug, run = _run_pair(M, G, B, 120, D, kname, ARD, 40)
        calls = plug.inference_method.engine.calls
        assert calls[0] == "set_data" and calls.count("set_data") == 1   # X, Y stayed resident across the optimisation
        assert calls.count("exact_eval") >= run.funct_eval


@pytest.mark.gpu
def test_stock_gpregression_optimizes_through_the_plugin_gpu():
    M, G, B = _plugin()
    for (kname, ARD, D, N) in (("RBF", True, 8, 1500), ("Matern52", False, 32, 900)):
        _run_pair(M, G, B, N, D, kname, ARD, 40)


def test_stock_gpregression_with_a_sum_product_kernel_through_the_plugin_cpu():
    """GPy's own Add / Prod (add.py, prod.py) over plugin leaves + the reference's White / Bias, inside the reference's own
    GPRegression: the plugin inference flattens the kernel and makes ONE engine call per evaluation; the stock model class
    distributes the gradients through (plugin) Add / Prod.update_gradients_full; same numbers as the all-stock model."""
    from test_host_model_cpu import FakeEngine as MultiFakeEngine
    from test_gpy_plugin_cpu import fake_ffi
    from oracle import ref_gpy
    from gpy_b200 import gpy_plugin
    M = ref_gpy.load_models()
    C = ref_gpy.load_combination()
    G = M.G
    ffi = fake_ffi()
    ffi.Engine = MultiFakeEngine
    B = gpy_plugin.make(G.RBF, G.Exponential, G.Matern32, G.Matern52, G.ExactGaussianInference, ffi=ffi, Add=C.Add, Prod=C.Prod)
    X, Y = o.synthetic(90, 4, seed=6)

    def build(K, Add, Prod):
        return Add([Prod([K.RBF(2, variance=1.2, lengthscale=[1.0, 2.0], ARD=True, active_dims=[0, 1]),
                          K.Matern32(2, variance=0.8, lengthscale=1.5, active_dims=[2, 3])]),
                    K.Matern52(4, variance=0.5, lengthscale=2.0), C.White(4, variance=0.05), C.Bias(4, variance=0.3)])

    stock = M.GPRegression(X, Y, build(G, C.Add, C.Prod), noise_var=0.1)
    plug = M.GPRegression(X, Y, build(B, B.Add, B.Prod), noise_var=0.1)
    plug.inference_method = B.ExactGaussianInference()
    plug.parameters_changed()
    eng = plug.inference_method.engine
    assert eng.calls == ["set_data", "exact_eval_multi"]
    assert abs(plug.log_likelihood() - stock.log_likelihood()) < 1e-9
    np.testing.assert_allclose(plug.gradient, stock.gradient, rtol=1e-8, atol=1e-10)
    plug.optimize(max_iters=15)
    stock.optimize(max_iters=15)
    assert abs(plug.log_likelihood() - stock.log_likelihood()) <= 1e-5 * max(1.0, abs(stock.log_likelihood()))
    assert eng.calls.count("set_data") == 1 and "exact_eval" not in eng.calls
