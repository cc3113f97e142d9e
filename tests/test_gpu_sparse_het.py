"""GPU parity of the heteroscedastic VarDTC evaluation (gpx_sparse_eval_het; the `het_noise` branches of
GPy/inference/latent_function_inference/var_dtc.py:127-128,221-227,241-257,267-269) against the CPU oracle, which is pinned
to the unmodified reference VarDTC + HeteroscedasticGaussian (tests/test_reference_crosscheck.py), and against the
fixtures the reference itself produced (tests/golden/sparse_het/). Tolerances as for the scalar-noise evaluation: 1e-8
relative on the bound, 1e-6 relative on gradients; dL/dZ and the per-point noise gradients dL_dR (sums of terms of size
beta_n^2 that cancel) with an absolute floor relative to their largest entry."""
import os

import numpy as np
import pytest

import gpy_b200
from gpy_b200 import _ffi
from oracle import gpy_oracle as o

pytestmark = pytest.mark.gpu


def _check(lml, g, dZ, dR, lml0, g0, Zg0, tag=""):
    nk = g.size
    assert abs(lml - lml0) <= 1e-8 * max(1.0, abs(lml0)), (tag, lml, lml0)
    np.testing.assert_allclose(g, g0[:nk], rtol=1e-6, atol=1e-8, err_msg=tag)
    np.testing.assert_allclose(dZ, Zg0, rtol=1e-6, atol=1e-7 * np.abs(Zg0).max(), err_msg=tag)
    dR0 = g0[nk:].reshape(dR.shape)
    np.testing.assert_allclose(dR, dR0, rtol=1e-6, atol=1e-7 * np.abs(dR0).max(), err_msg=tag)


@pytest.mark.parametrize("kind,ARD,N,M,D,P", [("rbf", True, 500, 60, 3, 1), ("matern32", False, 1300, 129, 2, 2),
                                              ("exponential", True, 257, 128, 4, 1), ("matern52", True, 2100, 300, 5, 1)])
def test_sparse_heteroscedastic_eval_matches_oracle(kind, ARD, N, M, D, P):
    rng = np.random.default_rng(N + M)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.stack([np.sin(X).sum(1) / np.sqrt(D) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
    Z = X[rng.permutation(N)[:M]].copy() + 0.01 * rng.standard_normal((M, D))
    ls = np.sqrt(D) * rng.uniform(0.7, 1.3, D) if ARD else float(np.sqrt(D) * 0.9)
    nv = rng.uniform(0.01, 0.3, N)
    e = _ffi.Engine(0)
    try:
        e.sparse_set_data(X, Y)
        lml, g, dZ, dR = e.sparse_eval_het(kind, ARD, 1.3, ls, Z, nv)
        lml0, g0, Zg0, res = o.sparse_eval(X, Y, Z, kind, ARD, 1.3, ls, nv)
        _check(lml, g, dZ, dR, lml0, g0, Zg0, "%s N=%d M=%d" % (kind, N, M))
        np.testing.assert_allclose(e.sparse_get("woodbury_vector"), res["woodbury_vector"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(e.sparse_get("woodbury_inv"), res["woodbury_inv"], rtol=1e-4,
                                   atol=1e-6 * np.abs(res["woodbury_inv"]).max())
        # the scalar-noise entry point on the same context afterwards (shared buffers: Yb, dL_dKnm^T) and het again
        lml_s, g_s, dZ_s = e.sparse_eval(kind, ARD, 1.3, ls, Z, 0.05)
        l0, gs0, Zs0, _ = o.sparse_eval(X, Y, Z, kind, ARD, 1.3, ls, 0.05)
        assert abs(lml_s - l0) <= 1e-8 * max(1.0, abs(l0))
        np.testing.assert_allclose(g_s, gs0, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(dZ_s, Zs0, rtol=1e-6, atol=1e-7 * np.abs(Zs0).max())
        lml2, g2, dZ2, dR2 = e.sparse_eval_het(kind, ARD, 1.3, ls, Z, nv)
        assert abs(lml2 - lml) <= 1e-11 * abs(lml)                                      # nothing of the scalar run is left
        np.testing.assert_allclose(g2, g, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(dR2, dR, rtol=1e-10, atol=1e-10 * np.abs(dR).max())
    finally:
        e.close()


def test_sparse_heteroscedastic_constant_vector_equals_scalar_noise():
    """beta_n = beta for all n: the het_noise route must reproduce the scalar route's bound and kernel gradients, and
    its N per-point noise gradients must sum to the scalar noise gradient."""
    rng = np.random.default_rng(77)
    N, M, D = 900, 100, 3
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) + 0.1 * rng.standard_normal((N, 1))
    Z = X[rng.permutation(N)[:M]].copy() + 0.01 * rng.standard_normal((M, D))
    ls = np.array([1.4, 1.9, 2.2])
    e = _ffi.Engine(0)
    try:
        e.sparse_set_data(X, Y)
        l0, g0, Z0 = e.sparse_eval("rbf", True, 1.1, ls, Z, 0.06)
        l1, g1, Z1, dR = e.sparse_eval_het("rbf", True, 1.1, ls, Z, np.full(N, 0.06))
        assert abs(l0 - l1) <= 1e-9 * abs(l0)
        np.testing.assert_allclose(g1, g0[:-1], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(dR.sum(), g0[-1], rtol=1e-6)
        np.testing.assert_allclose(Z1, Z0, rtol=1e-6, atol=1e-8 * np.abs(Z0).max())
    finally:
        e.close()


def test_sparse_heteroscedastic_golden_fixtures_and_model():
    """tests/golden/sparse_het/*.npz (numbers of the reference's own VarDTC + HeteroscedasticGaussian) through the mirror
    model: SparseGPRegression(likelihood=HeteroscedasticGaussian) -> bound, gradients in the reference's parameter order
    [Z, kern.variance, kern.lengthscale, het_Gauss.variance], inducing-input gradients, woodbury vector."""
    gdir = os.path.join(os.path.dirname(__file__), "golden", "sparse_het")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".npz"))
    assert files
    for fn in files:
        z = np.load(os.path.join(gdir, fn))
        kind, ARD, D = str(z["kind"]), bool(z["ARD"]), z["X"].shape[1]
        N = z["X"].shape[0]
        ls = z["lengthscale"] if ARD else float(z["lengthscale"])
        cls = {"rbf": gpy_b200.RBF, "exponential": gpy_b200.Exponential, "matern32": gpy_b200.Matern32,
               "matern52": gpy_b200.Matern52}[kind]
        k = cls(D, variance=float(z["variance"]), lengthscale=ls, ARD=ARD)
        meta = {"output_index": np.arange(N)[:, None]}
        lik = gpy_b200.HeteroscedasticGaussian(meta)
        lik.variance.values[...] = z["noise_variances"]
        m = gpy_b200.SparseGPRegression(z["X"], z["Y"], kernel=k, Z=z["Z"], likelihood=lik, Y_metadata=meta)
        lml0 = float(z["lml"])
        assert abs(m.log_likelihood() - lml0) <= 1e-8 * max(1.0, abs(lml0)), fn
        g = np.concatenate([k.variance.gradient, k.lengthscale.gradient])
        g0 = z["grad"]
        np.testing.assert_allclose(g, g0[:g.size], rtol=1e-6, atol=1e-8, err_msg=fn)
        dR0 = g0[g.size:]
        np.testing.assert_allclose(m.likelihood.variance.gradient, dR0, rtol=1e-6, atol=1e-7 * np.abs(dR0).max(), err_msg=fn)
        np.testing.assert_allclose(m.Z.gradient, z["Zgrad"], rtol=1e-6, atol=1e-7 * np.abs(z["Zgrad"]).max(), err_msg=fn)
        np.testing.assert_allclose(m.posterior.woodbury_vector, z["woodbury_vector"], rtol=1e-6, atol=1e-7, err_msg=fn)
        np.testing.assert_allclose(m.posterior.woodbury_inv, z["woodbury_inv"], rtol=1e-4,
                                   atol=1e-6 * np.abs(z["woodbury_inv"]).max(), err_msg=fn)
        if N <= 300:
            assert m.checkgrad(step=1e-5), fn
