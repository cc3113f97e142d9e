"""CPU tests of the oracle (the checker) — relational tests mirrored from the reference's own suite, since the
reference holds no numeric golden vectors for this path."""
import os

import numpy as np
import pytest
import scipy.linalg

from oracle import gpy_oracle as o

KIND_ARD = [(k, a) for k in o.KINDS for a in (False, True)]


def _theta(D, ARD, seed=0):
    rng = np.random.default_rng(seed)
    ls = rng.uniform(0.8, 2.5, D) if ARD else float(rng.uniform(0.8, 2.5))
    return 0.5 + rng.uniform(), ls, 0.01 + 0.2 * rng.uniform()


@pytest.mark.parametrize("kind,ARD", KIND_ARD)
def test_model_gradient_matches_finite_differences(kind, ARD):
    """GPy/testing/test_model.py:790-916 (TestGradient.check_model -> m.checkgrad()) for GPRegression x
    {rbf, matern52, matern32, exponential} x {iso, ARD}: 40x2 uniform(-3,3) inputs, sin targets + 0.05 noise."""
    rng = np.random.default_rng(1)
    X = rng.uniform(-3, 3, (40, 2))
    Y = np.sin(X[:, :1]) + 0.05 * rng.standard_normal((40, 1))
    var, ls, noise = _theta(2, ARD, 3)
    lml, g, _ = o.eval_lml_grad(X, Y, kind, ARD, var, ls, noise)
    th = np.concatenate([[var], np.atleast_1d(ls), [noise]])

    def f(t):
        return o.eval_lml_grad(X, Y, kind, ARD, t[0], t[1:-1] if ARD else t[1], t[-1])[0]

    for i in range(th.size):
        h = 1e-6 * th[i]
        tp, tm = th.copy(), th.copy()
        tp[i] += h
        tm[i] -= h
        fd = (f(tp) - f(tm)) / (2 * h)
        assert abs(fd - g[i]) <= 1e-5 * max(1.0, abs(fd)), (kind, ARD, i, fd, g[i])


@pytest.mark.parametrize("kind,ARD", KIND_ARD)
def test_kernel_is_psd_and_dK_dtheta(kind, ARD):
    """GPy/testing/test_kernel.py:45-51,57-70: eigenvalues >= -1e-10 and dK/dtheta against finite differences with a
    fixed random dL_dK, X in R^{10x6}, X2 in R^{20x6}."""
    rng = np.random.default_rng(2)
    X, X2 = rng.standard_normal((10, 6)), rng.standard_normal((20, 6))
    var, ls, _ = _theta(6, ARD, 5)
    k = o.StationaryOracle(kind, 6, var, ls, ARD)
    assert np.linalg.eigvalsh(k.K(X)).min() >= -1e-10
    for XX2 in (None, X2):
        dL = rng.standard_normal((10, 10 if XX2 is None else 20))
        dv, dl = k.update_gradients_full(dL, X, XX2)
        th = np.concatenate([[var], np.atleast_1d(ls)])
        an = np.concatenate([[dv], np.atleast_1d(dl)])
        for i in range(th.size):
            h = 1e-6 * th[i]
            vals = []
            for sgn in (1, -1):
                t = th.copy()
                t[i] += sgn * h
                kk = o.StationaryOracle(kind, 6, t[0], t[1:] if ARD else t[1], ARD)
                vals.append(np.sum(kk.K(X, XX2) * dL))
            fd = (vals[0] - vals[1]) / (2 * h)
            assert abs(fd - an[i]) <= 1e-6 * max(1.0, abs(fd))


def test_native_lengthscale_grads_equal_numpy():
    """GPy/testing/test_cython.py:51-98: RBF(10), X 300x10, Z 20x10: native ARD reduction == NumPy reduction, for our C
    restatement of the Cython loop and for the reference's own stationary_utils.c when it was built."""
    rng = np.random.default_rng(3)
    X, Z = rng.standard_normal((300, 10)), rng.standard_normal((20, 10))
    ls = np.ones(10)
    libs = o._load_native()
    if libs["port"] is None:
        pytest.skip("oracle/_build not built")
    for tmp, A, B in ((rng.standard_normal((300, 300)), X, X), (rng.standard_normal((300, 20)), X, Z)):
        g1 = o.lengthscale_grads_pure(tmp, A, B, ls)
        assert np.allclose(g1, o.lengthscale_grads_native(tmp, A, B, ls, "port"))
        if libs["ref"] is not None:
            assert np.allclose(g1, o.lengthscale_grads_native(tmp, A, B, ls, "ref"))


def _corrupt(seed=0):
    """GPy/testing/test_linalg.py:8-18."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((20, 100))
    A = A.dot(A.T)
    vals, vectors = np.linalg.eigh(A)
    vals[vals.argmin()] = 0
    default_jitter = 1e-6 * np.mean(vals)
    vals[vals.argmin()] = -default_jitter * (10 ** 3.5)
    return (vectors * vals).dot(vectors.T)


def test_jitchol_needs_exactly_five_rounds():
    """GPy/testing/test_linalg.py:20-37."""
    A = _corrupt()
    L, jit = o.jitchol(A, maxtries=5)
    diff = L.dot(L.T) - A
    np.testing.assert_allclose(diff, np.eye(20) * np.diag(diff).mean(), atol=1e-13)
    assert jit > 0
    with pytest.raises(scipy.linalg.LinAlgError):
        o.jitchol(A, maxtries=4)


def test_raw_predict_matches_pinv_formula():
    """GPy/testing/test_model.py:83-105."""
    rng = np.random.default_rng(4)
    X, Xn = rng.uniform(-3, 3, (30, 1)), rng.uniform(-3, 3, (12, 1))
    Y = np.sin(X) + 0.05 * rng.standard_normal((30, 1))
    k = o.StationaryOracle("rbf", 1, 1.3, 0.9, False)
    noise = 0.5
    res = o.exact_inference(k, X, Y, noise)
    Kinv = np.linalg.pinv(k.K(X) + np.eye(30) * (noise + 1e-8))
    K_hat = k.K(Xn) - k.K(Xn, X).dot(Kinv).dot(k.K(X, Xn))
    mu_hat = k.K(Xn, X).dot(Kinv).dot(Y)
    mu, cov = o.raw_predict(k, X, res["L"], res["alpha"], Xn, full_cov=True)
    np.testing.assert_almost_equal(K_hat, cov)
    np.testing.assert_almost_equal(mu_hat, mu)
    mu, var = o.raw_predict(k, X, res["L"], res["alpha"], Xn)
    np.testing.assert_almost_equal(np.diag(K_hat)[:, None], var)


def test_golden_fixtures_reproduce():
    """Every committed fixture (tests/golden/*.npz, generated by tests/golden/make_golden.py) must be reproduced by the
    oracle bit-for-bit up to BLAS thread-count reordering (1e-10 relative)."""
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".npz"))
    assert files, "no golden fixtures committed"
    for fn in files:
        z = np.load(os.path.join(gdir, fn))
        kind, ARD = str(z["kind"]), bool(z["ARD"])
        ls = z["lengthscale"] if ARD else float(z["lengthscale"])
        lml, g, _ = o.eval_lml_grad(z["X"], z["Y"], kind, ARD, float(z["variance"]), ls, float(z["noise"]))
        assert abs(lml - float(z["lml"])) <= 1e-9 * max(1.0, abs(float(z["lml"]))), fn
        np.testing.assert_allclose(g, z["grad"], rtol=1e-8, atol=1e-10, err_msg=fn)


def test_golden_heteroscedastic_fixtures_reproduce():
    """tests/golden/het/*.npz (one noise variance per point, produced by the reference's HeteroscedasticGaussian)."""
    gdir = os.path.join(os.path.dirname(__file__), "golden", "het")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".npz"))
    assert files
    for fn in files:
        z = np.load(os.path.join(gdir, fn))
        kind, ARD = str(z["kind"]), bool(z["ARD"])
        ls = z["lengthscale"] if ARD else float(z["lengthscale"])
        lml, g, _ = o.eval_lml_grad(z["X"], z["Y"], kind, ARD, float(z["variance"]), ls, z["noise_variances"])
        assert abs(lml - float(z["lml"])) <= 1e-9 * max(1.0, abs(float(z["lml"]))), fn
        np.testing.assert_allclose(g, z["grad"], rtol=1e-8, atol=1e-10, err_msg=fn)


def test_golden_sparse_fixtures_reproduce():
    """tests/golden/sparse/*.npz (SparseGPRegression / VarDTC numbers produced by the reference's own objects)."""
    gdir = os.path.join(os.path.dirname(__file__), "golden", "sparse")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".npz"))
    assert files
    for fn in files:
        z = np.load(os.path.join(gdir, fn))
        kind, ARD = str(z["kind"]), bool(z["ARD"])
        ls = z["lengthscale"] if ARD else float(z["lengthscale"])
        lml, g, Zg, _ = o.sparse_eval(z["X"], z["Y"], z["Z"], kind, ARD, float(z["variance"]), ls, float(z["noise"]))
        assert abs(lml - float(z["lml"])) <= 1e-8 * max(1.0, abs(float(z["lml"]))), fn
        np.testing.assert_allclose(g, z["grad"], rtol=1e-6, atol=1e-8, err_msg=fn)
        np.testing.assert_allclose(Zg, z["Zgrad"], rtol=1e-6, atol=1e-8, err_msg=fn)


def test_golden_sparse_heteroscedastic_fixtures_reproduce():
    """tests/golden/sparse_het/*.npz: VarDTC with one noise variance per data point (var_dtc.py:127-128,221-227,241-257,
    267-269), numbers produced by the reference's own VarDTC + HeteroscedasticGaussian objects."""
    gdir = os.path.join(os.path.dirname(__file__), "golden", "sparse_het")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".npz"))
    assert files
    for fn in files:
        z = np.load(os.path.join(gdir, fn))
        assert "GPy" in str(z["source"]), fn
        kind, ARD = str(z["kind"]), bool(z["ARD"])
        ls = z["lengthscale"] if ARD else float(z["lengthscale"])
        lml, g, Zg, res = o.sparse_eval(z["X"], z["Y"], z["Z"], kind, ARD, float(z["variance"]), ls, z["noise_variances"])
        assert g.size == z["grad"].size == 1 + np.size(ls) + z["X"].shape[0] * z["Y"].shape[1], fn
        assert abs(lml - float(z["lml"])) <= 1e-8 * max(1.0, abs(float(z["lml"]))), fn
        np.testing.assert_allclose(g, z["grad"], rtol=1e-6, atol=1e-7, err_msg=fn)
        np.testing.assert_allclose(Zg, z["Zgrad"], rtol=1e-6, atol=1e-7, err_msg=fn)
        np.testing.assert_allclose(res["woodbury_vector"], z["woodbury_vector"], rtol=1e-6, atol=1e-8, err_msg=fn)


def test_sparse_heteroscedastic_reduces_to_homoscedastic():
    """A constant per-point noise vector gives the bound and kernel gradients of the scalar-noise branch, and the N
    per-point noise gradients sum to the scalar one (the het_noise branch is the same bound, differentiated per point)."""
    rng = np.random.default_rng(3)
    N, M, D = 120, 17, 2
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) + 0.1 * rng.standard_normal((N, 1))
    Z = X[:M] + 0.01
    l0, g0, Z0, _ = o.sparse_eval(X, Y, Z, "rbf", True, 1.2, np.array([1.1, 1.9]), 0.04)
    l1, g1, Z1, _ = o.sparse_eval(X, Y, Z, "rbf", True, 1.2, np.array([1.1, 1.9]), np.full(N, 0.04))
    assert abs(l0 - l1) <= 1e-9 * abs(l0)
    np.testing.assert_allclose(g1[:3], g0[:3], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(g1[3:].sum(), g0[3], rtol=1e-7)
    np.testing.assert_allclose(Z1, Z0, rtol=1e-7, atol=1e-9)


def test_logexp_roundtrip():
    x = np.linspace(-20, 50, 50)
    np.testing.assert_allclose(o.logexp_finv(o.logexp_f(x)), x, rtol=1e-9, atol=1e-6)
    f = o.logexp_f(x)
    h = 1e-6
    np.testing.assert_allclose(o.logexp_gradfactor(f), (o.logexp_f(x + h) - o.logexp_f(x - h)) / (2 * h), rtol=1e-6, atol=1e-9)
