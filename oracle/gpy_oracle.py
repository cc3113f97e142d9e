"""CPU oracle: a line-by-line NumPy/SciPy restatement of the reference's exact-GP hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Every function cites the reference file:line it
follows (paths relative to /root/reference). The restatement issues the SAME LAPACK/BLAS entry points as
the reference (dsyrk, dpotrf, dtrtri, dpotri, dpotrs, dtrtrs) in the same order, including the work the
reference wastes (unused dtrtri, double symmetrify, second exp pass), so that timing it is timing the
reference's operation sequence.
"""
import ctypes
import os

import numpy as np
from scipy import linalg
from scipy.linalg import blas, lapack

KINDS = ("rbf", "exponential", "matern32", "matern52")
LOG_2_PI = np.log(2.0 * np.pi)  # GPy/inference/latent_function_inference/exact_gaussian_inference.py:8
JITTER = 1e-8  # exact_gaussian_inference.py:56 (unconditional)

_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------------------
# GPy/util/linalg.py, GPy/util/diag.py
# ----------------------------------------------------------------------------------------------------------
NATIVE_LINALG = False  # True: mirror `use_linalg_cython` (GPy/util/linalg.py:14-18) with the C loop of oracle_c.c


def symmetrify(A, upper=False):
    """GPy/util/linalg.py:356-379: the Cython double loop (linalg_cython.pyx:9-18, restated in oracle_c.c) when the
    native helpers are enabled, else the NumPy fallback (_symmetrify_numpy, :374-379). In place."""
    if NATIVE_LINALG and A.flags.c_contiguous or (NATIVE_LINALG and A.flags.f_contiguous):
        lib = _load_native()["port"]
        if lib is not None and A.dtype == np.float64 and A.ndim == 2 and A.shape[0] == A.shape[1]:
            lib.oracle_symmetrify.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int]
            lib.oracle_symmetrify.restype = None
            # an F-ordered array is the transpose of a C-ordered one: swap the direction
            up = bool(upper) if A.flags.c_contiguous else not bool(upper)
            lib.oracle_symmetrify(A.shape[0], A.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(up))
            return
    triu = np.triu_indices_from(A, k=1)
    if upper:
        A.T[triu] = A[triu]
    else:
        A[triu] = A.T[triu]


def tdot(mat):
    """GPy/util/linalg.py:299-320 tdot_blas: X X^T via dsyrk(lower=0) then symmetrify(upper=True)."""
    if mat.dtype != np.float64 or mat.ndim != 2:
        return np.dot(mat, mat.T)
    nn = mat.shape[0]
    out = np.zeros((nn, nn))
    mat = np.asfortranarray(mat)
    out = blas.dsyrk(alpha=1.0, a=mat, beta=0.0, c=out, overwrite_c=1, trans=0, lower=0)
    symmetrify(out, upper=True)
    return np.ascontiguousarray(out)


def diag_view(A):
    """GPy/util/diag.py:6-40: strided view of the main diagonal."""
    from numpy.lib.stride_tricks import as_strided
    assert A.ndim == 2 and A.shape[0] == A.shape[1]
    return as_strided(A, shape=(A.shape[0],), strides=((A.shape[0] + 1) * A.itemsize,))


def diag_add(A, b):
    """GPy/util/diag.py:48-52,85-98: A[i,i] += b in place (b squeezed)."""
    b = np.squeeze(b)
    dA = diag_view(A)
    np.add(dA, b, dA)
    return A


def jitchol(A, maxtries=5):
    """GPy/util/linalg.py:56-75: dpotrf(lower=1); on failure the jitter ladder mean(diag)*1e-6*10^k, k<maxtries.

    Returns (L, jitter_used). Raises numpy.linalg.LinAlgError like the reference."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L, 0.0
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            L = linalg.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
            return L, jitter
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtri(L):
    """GPy/util/linalg.py:217-227."""
    return lapack.dtrtri(np.asfortranarray(L), lower=1)[0]


def dpotri(L, lower=1):
    """GPy/util/linalg.py:127-145: dpotri + symmetrify."""
    R, info = lapack.dpotri(np.asfortranarray(L), lower=lower)
    symmetrify(R)
    return R, info


def dpotrs(L, B, lower=1):
    """GPy/util/linalg.py:116-125."""
    return lapack.dpotrs(np.asfortranarray(L), B, lower=lower)


def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
    """GPy/util/linalg.py:95-114."""
    return lapack.dtrtrs(np.asfortranarray(A), B, lower=lower, trans=trans, unitdiag=unitdiag)


def pdinv(A):
    """GPy/util/linalg.py:193-214: (Ai, L, Li, logdet) — including the dtrtri whose result the exact-GP caller
    never uses (:209) and the second symmetrify (:212)."""
    L, jit = jitchol(A)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    Li = dtrtri(L)
    Ai, _ = dpotri(L, lower=1)
    symmetrify(Ai)
    return Ai, L, Li, logdet


# ----------------------------------------------------------------------------------------------------------
# GPy/kern/src/stationary.py, rbf.py
# ----------------------------------------------------------------------------------------------------------
def unscaled_dist(X, X2=None):
    """GPy/kern/src/stationary.py:130-148."""
    if X2 is None:
        Xsq = np.sum(np.square(X), 1)
        r2 = -2.0 * tdot(X) + (Xsq[:, None] + Xsq[None, :])
        diag_view(r2)[:, ] = 0.0  # :138 force diagonal to be zero
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)
    X1sq = np.sum(np.square(X), 1)
    X2sq = np.sum(np.square(X2), 1)
    r2 = -2.0 * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def scaled_dist(X, X2, lengthscale, ARD):
    """GPy/kern/src/stationary.py:151-168: ARD divides X by l BEFORE the expansion, iso divides r AFTER."""
    if ARD:
        if X2 is not None:
            X2 = X2 / lengthscale
        return unscaled_dist(X / lengthscale, X2)
    return unscaled_dist(X, X2) / lengthscale


def K_of_r(kind, variance, r):
    """rbf.py:51-52; stationary.py:382-383 (Exponential), :488-489 (Matern32), :585-586 (Matern52)."""
    if kind == "rbf":
        return variance * np.exp(-0.5 * r ** 2)
    if kind == "exponential":
        return variance * np.exp(-r)
    if kind == "matern32":
        return variance * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)
    if kind == "matern52":
        return variance * (1 + np.sqrt(5.0) * r + 5.0 / 3 * r ** 2) * np.exp(-np.sqrt(5.0) * r)
    raise ValueError(kind)


def dK_dr(kind, variance, r):
    """rbf.py:177-178; stationary.py:385-386, :491-492, :588-589."""
    if kind == "rbf":
        return -r * K_of_r(kind, variance, r)
    if kind == "exponential":
        return -K_of_r(kind, variance, r)
    if kind == "matern32":
        return -3.0 * variance * r * np.exp(-np.sqrt(3.0) * r)
    if kind == "matern52":
        return variance * (10.0 / 3 * r - 5.0 * r - 5.0 * np.sqrt(5.0) / 3 * r ** 2) * np.exp(-np.sqrt(5.0) * r)
    raise ValueError(kind)


def inv_dist(X, X2, lengthscale, ARD):
    """GPy/kern/src/stationary.py:225-232: 1/r with 1/0 := 0."""
    dist = scaled_dist(X, X2, lengthscale, ARD).copy()
    return 1.0 / np.where(dist != 0.0, dist, np.inf)


def lengthscale_grads_pure(tmp, X, X2, lengthscale):
    """GPy/kern/src/stationary.py:234-235."""
    Q = X.shape[1]
    return -np.array([np.sum(tmp * np.square(X[:, q:q + 1] - X2[:, q:q + 1].T)) for q in range(Q)]) / lengthscale ** 3


_native = {}


def _load_native():
    """Optional compiled helpers. `oracle/_ref/libstationary_utils.so` is the reference's own
    GPy/kern/src/stationary_utils.c compiled where it lies (oracle/Makefile); `oracle/_build/liboracle_c.so`
    is our C restatement of the serial Cython loop stationary_cython.pyx:53-62."""
    if _native:
        return _native
    for key, rel in (("ref", "_ref/libstationary_utils.so"), ("port", "_build/liboracle_c.so")):
        p = os.path.join(_HERE, rel)
        _native[key] = ctypes.CDLL(p) if os.path.exists(p) else None
    return _native


def lengthscale_grads_native(tmp, X, X2, lengthscale, which="port"):
    """stationary.py:237-243 -> stationary_cython.pyx:53-62 (serial q,n,m loop). `which='ref'` calls the
    reference's own C `_lengthscale_grads` (stationary_utils.c:34-48, OpenMP over q)."""
    lib = _load_native()[which]
    if lib is None:
        raise RuntimeError("native oracle helper '%s' not built (run `make -C oracle`)" % which)
    N, M = tmp.shape
    Q = X.shape[1]
    X, X2, tmp = np.ascontiguousarray(X), np.ascontiguousarray(X2), np.ascontiguousarray(tmp)
    grads = np.zeros(Q)
    fn = lib._lengthscale_grads if which == "ref" else lib.oracle_lengthscale_grads
    dp = ctypes.POINTER(ctypes.c_double)
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]
    fn.restype = None
    fn(N, M, Q, tmp.ctypes.data_as(dp), X.ctypes.data_as(dp), X2.ctypes.data_as(dp), grads.ctypes.data_as(dp))
    return -grads / lengthscale ** 3


class StationaryOracle(object):
    """Restates GPy.kern.src.stationary.Stationary (stationary.py:23-243) for the four hot-path kernels.
    `lengthscale` is a scalar (iso) or a length-D vector (ARD), as in stationary.py:64-79."""

    def __init__(self, kind, input_dim, variance=1.0, lengthscale=None, ARD=False, native=None):
        assert kind in KINDS
        self.kind, self.input_dim, self.ARD = kind, input_dim, ARD
        self.variance = float(variance)
        if lengthscale is None:
            lengthscale = np.ones(input_dim) if ARD else 1.0
        self.lengthscale = np.asarray(lengthscale, dtype=np.float64).reshape(-1)
        if ARD:
            assert self.lengthscale.size == input_dim
        else:
            assert self.lengthscale.size == 1
        self.native = native  # None -> pure NumPy ARD reduction; "port"/"ref" -> compiled loops
        self.variance_gradient = None
        self.lengthscale_gradient = None
        # paramz `Cache_this(limit=3)` on _scaled_dist / K / dK_dr_via_X (stationary.py:105,117,130,150): results are
        # memoised per (X, X2) object identity while the parameters are unchanged. The parameters of an oracle
        # object never change, so an identity-keyed dict reproduces which passes the reference recomputes.
        self._cache = {}

    def _memo(self, tag, X, X2, fn):
        key = (tag, id(X), id(X2))
        hit = self._cache.get(key)
        if hit is not None and hit[0] is X and hit[1] is X2:
            return hit[2]
        val = fn()
        self._cache[key] = (X, X2, val)
        return val

    def _r(self, X, X2=None):
        return self._memo("r", X, X2, lambda: scaled_dist(X, X2, self.lengthscale, self.ARD))

    def K(self, X, X2=None):
        """stationary.py:105-115 (cached)."""
        return self._memo("K", X, X2, lambda: K_of_r(self.kind, self.variance, self._r(X, X2)))

    def Kdiag(self, X):
        """stationary.py:170-173."""
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret

    def update_gradients_full(self, dL_dK, X, X2=None):
        """stationary.py:193-213."""
        self.variance_gradient = np.sum(self.K(X, X2) * dL_dK) / self.variance
        dL_dr = dK_dr(self.kind, self.variance, self._r(X, X2)) * dL_dK  # dK_dr_via_X: second exp pass, r from cache
        if self.ARD:
            dist = self._r(X, X2).copy()  # _inv_dist, stationary.py:225-232 (r from cache, then copy/where/divide)
            tmp = dL_dr * (1.0 / np.where(dist != 0.0, dist, np.inf))
            if X2 is None:
                X2 = X
            if self.native:
                self.lengthscale_gradient = lengthscale_grads_native(tmp, X, X2, self.lengthscale, self.native)
            else:
                self.lengthscale_gradient = lengthscale_grads_pure(tmp, X, X2, self.lengthscale)
        else:
            r = self._r(X, X2)
            self.lengthscale_gradient = np.atleast_1d(-np.sum(dL_dr * r) / self.lengthscale)
        return self.variance_gradient, self.lengthscale_gradient

    def gradients_X(self, dL_dK, X, X2=None):
        """stationary.py:245-252 -> _gradients_X_pure (:340-352)."""
        invdist = inv_dist(X, X2, self.lengthscale, self.ARD)
        dL_dr = dK_dr(self.kind, self.variance, self._r(X, X2)) * dL_dK
        tmp = invdist * dL_dr
        if X2 is None:
            tmp = tmp + tmp.T
            X2 = X
        grad = np.empty(X.shape, dtype=np.float64)
        for q in range(self.input_dim):
            np.sum(tmp * (X[:, q][:, None] - X2[:, q][None, :]), axis=1, out=grad[:, q])
        return grad / self.lengthscale ** 2

    def update_gradients_diag(self, dL_dKdiag, X):
        """stationary.py:175-183 (reset-style diag gradient: variance only)."""
        self.variance_gradient = np.sum(dL_dKdiag)
        self.lengthscale_gradient = np.zeros_like(self.lengthscale)
        return self.variance_gradient, self.lengthscale_gradient


# ----------------------------------------------------------------------------------------------------------
# GPy/inference/latent_function_inference/exact_gaussian_inference.py, posterior.py, likelihoods/gaussian.py
# ----------------------------------------------------------------------------------------------------------
def exact_inference(kern, X, Y, noise_variance, K=None):
    """exact_gaussian_inference.py:37-74 with mean_function=None, Z_tilde=None.

    Returns dict(L, alpha, K, Wi, logdet, log_marginal, dL_dK, dL_dthetaL, dL_dm)."""
    YYT_factor = Y - 0  # :50
    if K is None:
        K = kern.K(X)  # :53
    Ky = K.copy()  # :55
    diag_add(Ky, noise_variance + JITTER)  # :56
    Wi, LW, LWi, W_logdet = pdinv(Ky)  # :58
    alpha, _ = dpotrs(LW, YYT_factor, lower=1)  # :60
    log_marginal = 0.5 * (-Y.size * LOG_2_PI - Y.shape[1] * W_logdet - np.sum(alpha * YYT_factor))  # :62
    dL_dK = 0.5 * (tdot(alpha) - Y.shape[1] * Wi)  # :70
    if np.ndim(noise_variance) > 0 and np.size(noise_variance) > 1:
        # HeteroscedasticGaussian with output_index = arange(N) (likelihoods/gaussian.py:356-362,
        # models/gp_heteroscedastic_regression.py:26-27): `variance` reaching :56 is the N-vector, :72 keeps the diagonal
        dL_dthetaL = np.diag(dL_dK).copy()  # gaussian.py:358-359
    else:
        dL_dthetaL = float(np.sum(np.diag(dL_dK)))  # :72 -> likelihoods/gaussian.py:78-79
    return dict(L=LW, alpha=alpha, K=K, Wi=Wi, logdet=W_logdet, log_marginal=float(log_marginal), dL_dK=dL_dK,
                dL_dthetaL=dL_dthetaL, dL_dm=alpha)


def eval_lml_grad(X, Y, kind, ARD, variance, lengthscale, noise_variance, native=None):
    """One GP.parameters_changed() (GPy/core/gp.py:269-282): returns (log_marginal, gradient) with the gradient in
    the order paramz exposes it: [kern.variance, kern.lengthscale (1 or D), Gaussian_noise.variance]
    (link order stationary.py:81, gp.py:106-107)."""
    global NATIVE_LINALG
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    kern = StationaryOracle(kind, X.shape[1], variance, lengthscale, ARD, native=native)
    prev = NATIVE_LINALG
    NATIVE_LINALG = bool(native)  # an installed GPy has its Cython helpers built: symmetrify + lengthscale_grads in C
    try:
        res = exact_inference(kern, X, Y, noise_variance)
        dvar, dlen = kern.update_gradients_full(res["dL_dK"], X)  # gp.py:280
    finally:
        NATIVE_LINALG = prev
    grad = np.concatenate([[dvar], np.atleast_1d(dlen), np.atleast_1d(res["dL_dthetaL"])])
    return res["log_marginal"], grad, res


def raw_predict(kern, X, L, alpha, Xnew, full_cov=False):
    """posterior.py:273-302 (PosteriorExact._raw_predict, 2-D woodbury_chol)."""
    Kx = kern.K(X, Xnew)
    mu = np.dot(Kx.T, alpha)
    if mu.ndim == 1:
        mu = mu.reshape(-1, 1)
    tmp = dtrtrs(L, Kx)[0]
    if full_cov:
        var = kern.K(Xnew) - tdot(tmp.T)
    else:
        var = (kern.Kdiag(Xnew) - np.square(tmp).sum(0))[:, None]
    return mu, var


def predict(kern, X, L, alpha, Xnew, noise_variance, full_cov=False):
    """gp.py:290-365 predict with the Gaussian likelihood (likelihoods/gaussian.py:102-110): adds the noise."""
    mu, var = raw_predict(kern, X, L, alpha, Xnew, full_cov)
    if full_cov:
        var = var + np.eye(var.shape[0]) * noise_variance
    else:
        var = var + noise_variance
    return mu, var


# ----------------------------------------------------------------------------------------------------------
# synthetic workload of record and the Logexp transform paramz applies to all three parameters
# ----------------------------------------------------------------------------------------------------------
def synthetic(N, D, seed=0):
    """rng=default_rng(seed); X~U(-3,3); Y=sum(sin X)/sqrt(D)+0.1 eps (mirrors test_model.py:796-807)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    f = np.sin(X).sum(1, keepdims=True) / np.sqrt(D)
    Y = f + 0.1 * rng.standard_normal((N, 1))
    return X, Y


def theta_bench(D, ARD=True):
    """theta_bench of SURVEY.md §8(d): variance 1, lengthscale sqrt(D) (per dim if ARD), noise 0.01."""
    ls = np.full(D, np.sqrt(D)) if ARD else np.sqrt(D)
    return 1.0, ls, 0.01


def logexp_f(x):
    """paramz.transformations.Logexp.f: theta = log(1+exp(x)) (stable form)."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 36.0, x, np.log1p(np.exp(np.clip(x, -np.inf, 36.0))))


def logexp_finv(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f > 36.0, f, np.log(np.expm1(f)))


def logexp_gradfactor(f):
    """d theta / d x for theta = log(1+e^x): 1 - exp(-theta)."""
    f = np.asarray(f, dtype=np.float64)
    return np.where(f > 36.0, 1.0, -np.expm1(-f))


# ----------------------------------------------------------------------------------------------------------
# composite kernels: GPy/kern/src/add.py:60-99, prod.py:59-68,377-396, static.py:63-99 (White), :142-185 (Bias), with
# active_dims slicing of every part (kernel_slice_operations.py:59-79)
# ----------------------------------------------------------------------------------------------------------
class StaticOracle(object):
    """White (static.py:63-99) / Bias (static.py:142-185)."""

    def __init__(self, kind, variance):
        self.kind, self.variance = kind, float(variance)

    def K(self, X, X2=None):
        n = X.shape[0]
        if self.kind == "white":
            return np.eye(n) * self.variance if X2 is None else np.zeros((n, X2.shape[0]))     # static.py:75-79
        return np.full((n, n if X2 is None else X2.shape[0]), self.variance, dtype=np.float64)   # static.py:155-157

    def Kdiag(self, X):
        return np.full(X.shape[0], self.variance)                                                # static.py:30-33

    def update_gradients_full(self, dL_dK, X, X2=None):
        if self.kind == "white":
            return (np.trace(dL_dK) if X2 is None else 0.0), np.zeros(0)                          # static.py:88-92
        return dL_dK.sum(), np.zeros(0)                                                           # static.py:172-173


def composite_parts(parts):
    """parts: list of dicts(kind, term, dims, variance, [lengthscale, ARD]) -> list of (oracle kernel, dims, term)."""
    out = []
    for p in parts:
        if p["kind"] in ("white", "bias"):
            out.append((StaticOracle(p["kind"], p["variance"]), None, p["term"]))
        else:
            out.append((StationaryOracle(p["kind"], len(p["dims"]), p["variance"], p["lengthscale"], p.get("ARD", False)),
                        np.asarray(p["dims"]), p["term"]))
    return out


def _slice(X, dims):
    return X if dims is None else np.ascontiguousarray(X[:, dims])


def composite_K(kparts, X, X2=None):
    """Add.K over the terms (add.py:60-74) of Prod.K over the factors (prod.py:59-68)."""
    terms = {}
    for (k, dims, t) in kparts:
        Kp = k.K(_slice(X, dims), None if X2 is None else _slice(X2, dims))
        terms[t] = Kp if t not in terms else terms[t] * Kp
    return sum(terms[t] for t in sorted(terms))


def composite_Kdiag(kparts, X):
    terms = {}
    for (k, dims, t) in kparts:
        Kd = k.Kdiag(_slice(X, dims))
        terms[t] = Kd if t not in terms else terms[t] * Kd
    return sum(terms[t] for t in sorted(terms))


def composite_eval_lml_grad(X, Y, parts, noise_variance):
    """One GP.parameters_changed() with a composite kernel: -> (lml, grad [per part: variance, lengthscale.. ; noise], res).
    Add.update_gradients_full hands dL_dK to every term (add.py:76-77); inside a product each factor sees dL_dK times the
    other factors' K (prod.py:377-385)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    kparts = composite_parts(parts)
    K = composite_K(kparts, X)
    Ky = K.copy()
    diag_add(Ky, noise_variance + 1e-8)
    Wi, LW, LWi, W_logdet = pdinv(Ky)
    alpha, _ = dpotrs(LW, Y, lower=1)
    lml = 0.5 * (-Y.size * np.log(2 * np.pi) - Y.shape[1] * W_logdet - np.sum(alpha * Y))
    dL_dK = 0.5 * (tdot(alpha) - Y.shape[1] * Wi)
    grad = []
    for i, (k, dims, t) in enumerate(kparts):
        other = None
        for j, (k2, dims2, t2) in enumerate(kparts):
            if j != i and t2 == t:
                Kj = k2.K(_slice(X, dims2))
                other = Kj if other is None else other * Kj
        dv, dl = k.update_gradients_full(dL_dK if other is None else dL_dK * other, _slice(X, dims), None)
        grad.append(np.concatenate([[dv], np.atleast_1d(dl).reshape(-1)]))
    grad.append([np.trace(dL_dK)])
    return float(lml), np.concatenate(grad), dict(L=LW, alpha=alpha, K=K, Wi=Wi, dL_dK=dL_dK, kparts=kparts)


def optimize_lbfgsb(X, Y, kind, ARD, variance, lengthscale, noise_variance, max_iters=1000, gtol=1e-5,
                    ftol=2.220446049250313e-09):
    """The optimisation loop of GP.optimize (GPy/core/gp.py:663-684 -> paramz Model.optimize, default 'lbfgsb' =
    scipy.optimize.fmin_l_bfgs_b) on this oracle's objective: minimise -LML (model.py:97-109) over the Logexp-transformed
    [variance, lengthscale.., noise] (stationary.py:78-79, gaussian.py:43), gradient chain-ruled through d theta/d x
    (model.py:111-128 + paramz _transform_gradients). paramz itself is absent, so the trajectory is *parity unpinned*
    (SURVEY.md §8c ii); this restatement gives the device path a CPU run of the same loop to compare final values with.
    -> (lml_final, theta_final [variance, lengthscale.., noise], n_evals, lml_initial)"""
    from scipy.optimize import fmin_l_bfgs_b
    D = X.shape[1]
    nl = D if ARD else 1
    th0 = np.concatenate([[variance], np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)).reshape(-1), [noise_variance]])
    count = {"n": 0, "first": None}

    def fg(x):
        th = logexp_f(x)
        ls = th[1:1 + nl] if ARD else float(th[1])
        lml, g, _ = eval_lml_grad(X, Y, kind, ARD, float(th[0]), ls, float(th[-1]))
        count["n"] += 1
        if count["first"] is None:
            count["first"] = lml
        return -lml, -g * logexp_gradfactor(th)

    x, f, d = fmin_l_bfgs_b(fg, logexp_finv(th0), maxfun=max_iters, maxiter=max_iters, pgtol=gtol,
                            factr=ftol / np.finfo(float).eps)
    return -f, logexp_f(x), count["n"], count["first"]


# ----------------------------------------------------------------------------------------------------------
# sparse GP regression: GPy/inference/latent_function_inference/var_dtc.py:66-276 (VarDTC, Gaussian likelihood,
# homoscedastic noise, certain inputs, no mean function) + the gradient wiring of GPy/core/sparse_gp.py:108-119
# ----------------------------------------------------------------------------------------------------------
VARDTC_JITTER = 1e-8  # var_dtc.py:24 const_jitter


def backsub_both_sides(L, X, transpose="left"):
    """GPy/util/linalg.py:381-390: L^-T X L^-1 (transpose='left') or L^-1 X L^-T."""
    if transpose == "left":
        tmp, _ = dtrtrs(L, X, lower=1, trans=1)
        return dtrtrs(L, tmp.T, lower=1, trans=1)[0].T
    tmp, _ = dtrtrs(L, X, lower=1, trans=0)
    return dtrtrs(L, tmp.T, lower=1, trans=0)[0].T


def vardtc_inference(kern, X, Z, noise_variance, Y):
    """var_dtc.py:66-215. `noise_variance` scalar (Gaussian) or one value per data point (HeteroscedasticGaussian:
    `het_noise`, :82-84, the branches :127-128, :221-227, :241-257, :267-269). Returns dict(log_marginal, dL_dKmm,
    dL_dKdiag, dL_dKnm, dL_dthetaL, woodbury_vector, woodbury_inv, Lm, Kmm); dL_dthetaL is the scalar sum for a scalar
    noise and the N x P array dL_dR for per-point noise (heteroscedastic_gaussian.py:33-34 indexes it by output_index)."""
    num_data, output_dim = Y.shape
    num_inducing = Z.shape[0]
    precision = 1.0 / np.fmax(noise_variance, VARDTC_JITTER)  # :79-80
    precision = np.asarray(precision, dtype=np.float64)
    if precision.ndim == 1:  # :82-83
        precision = precision[:, None]
    het_noise = precision.size > 1  # :84
    if not het_noise:
        precision = float(precision)
    beta = precision
    VVT_factor = precision * Y  # :89
    trYYT = np.einsum("ij,ij->", Y, Y)  # :37
    Kmm = kern.K(Z).copy()  # :93
    diag_add(Kmm, VARDTC_JITTER)  # :94
    Lm, _ = jitchol(Kmm)  # :95
    psi0 = kern.Kdiag(X)  # :124
    psi1 = kern.K(X, Z)  # :126
    tmp = psi1 * np.sqrt(precision)  # :127-130 (N x M times N x 1 for per-point noise)
    tmp, _ = dtrtrs(Lm, tmp.T, lower=1)  # :131
    A = tdot(tmp)  # :132
    B = np.eye(num_inducing) + A  # :135
    LB, _ = jitchol(B)  # :136
    tmp, _ = dtrtrs(Lm, psi1.T, lower=1, trans=0)  # :139
    _LBi_Lmi_psi1, _ = dtrtrs(LB, tmp, lower=1, trans=0)  # :140
    _LBi_Lmi_psi1Vf = np.dot(_LBi_Lmi_psi1, VVT_factor)  # :141
    tmp, _ = dtrtrs(LB, _LBi_Lmi_psi1Vf, lower=1, trans=1)  # :142
    Cpsi1Vf, _ = dtrtrs(Lm, tmp, lower=1, trans=1)  # :143
    delit = tdot(_LBi_Lmi_psi1Vf)  # :148
    data_fit = np.trace(delit)  # :149
    DBi_plus_BiPBi = backsub_both_sides(LB, output_dim * np.eye(num_inducing) + delit)  # :150
    delit = -0.5 * DBi_plus_BiPBi  # :152
    delit += -0.5 * B * output_dim
    delit += output_dim * np.eye(num_inducing)
    dL_dKmm = backsub_both_sides(Lm, delit)  # :156
    # _compute_dL_dpsi (:217-234), certain inputs
    dL_dpsi0 = -0.5 * output_dim * (beta * np.ones([num_data, 1])).flatten()
    dL_dpsi1 = np.dot(VVT_factor, Cpsi1Vf.T)
    dL_dpsi2_beta = 0.5 * backsub_both_sides(Lm, output_dim * np.eye(num_inducing) - DBi_plus_BiPBi)
    if het_noise:
        dL_dpsi1 += 2.0 * np.dot(dL_dpsi2_beta, (psi1 * beta).T).T  # :226
    else:
        dL_dpsi2 = beta * dL_dpsi2_beta
        dL_dpsi1 += 2.0 * np.dot(psi1, dL_dpsi2)  # :232
    # _compute_log_marginal_likelihood (:265-276)
    if het_noise:
        lik_1 = (-0.5 * num_data * output_dim * np.log(2.0 * np.pi) + 0.5 * output_dim * np.sum(np.log(beta))
                 - 0.5 * np.sum(beta.ravel() * np.square(Y).sum(axis=-1)))  # :268
        lik_2 = -0.5 * output_dim * (np.sum(beta.flatten() * psi0) - np.trace(A))  # :269
    else:
        lik_1 = -0.5 * num_data * output_dim * (np.log(2.0 * np.pi) - np.log(beta)) - 0.5 * beta * trYYT
        lik_2 = -0.5 * output_dim * (np.sum(beta * psi0) - np.trace(A))
    lik_3 = -output_dim * (np.sum(np.log(np.diag(LB))))
    lik_4 = 0.5 * data_fit
    log_marginal = lik_1 + lik_2 + lik_3 + lik_4
    # _compute_dL_dR (:237-263)
    if het_noise:  # :241-257
        LBi, _ = dtrtrs(LB, np.eye(LB.shape[0]))
        Lmi_psi1, _ = dtrtrs(Lm, psi1.T, lower=1, trans=0)
        _LBi_Lmi_psi1, _ = dtrtrs(LB, Lmi_psi1, lower=1, trans=0)
        dL_dR = -0.5 * beta + 0.5 * VVT_factor ** 2
        dL_dR += 0.5 * output_dim * (psi0 - np.sum(Lmi_psi1 ** 2, 0))[:, None] * beta ** 2
        dL_dR += 0.5 * np.sum(np.dot(LBi.T, np.dot(LBi, Lmi_psi1)) * Lmi_psi1, 0)[:, None] * beta ** 2
        dL_dR += -np.dot(_LBi_Lmi_psi1Vf.T, _LBi_Lmi_psi1).T * Y * beta ** 2
        dL_dR += 0.5 * np.dot(_LBi_Lmi_psi1Vf.T, _LBi_Lmi_psi1).T ** 2 * beta ** 2
        dL_dthetaL = dL_dR
    else:
        dL_dR = -0.5 * num_data * output_dim * beta + 0.5 * trYYT * beta ** 2
        dL_dR += 0.5 * output_dim * (psi0.sum() * beta ** 2 - np.trace(A) * beta)
        dL_dR += beta * (0.5 * np.sum(A * DBi_plus_BiPBi) - data_fit)
        dL_dthetaL = float(np.sum(dL_dR))
    # posterior (:201-214)
    Bi = -dpotri(LB, lower=1)[0]
    diag_add(Bi, 1)
    woodbury_inv = backsub_both_sides(Lm, Bi)
    return dict(log_marginal=float(log_marginal), dL_dKmm=dL_dKmm, dL_dKdiag=dL_dpsi0, dL_dKnm=dL_dpsi1,
                dL_dthetaL=dL_dthetaL, woodbury_vector=Cpsi1Vf, woodbury_inv=woodbury_inv, Lm=Lm, Kmm=Kmm)


def sparse_eval(X, Y, Z, kind, ARD, variance, lengthscale, noise_variance):
    """One SparseGP.parameters_changed() (GPy/core/sparse_gp.py:76-119): returns (log_marginal,
    grad [kern.variance, kern.lengthscale.., Gaussian_noise.variance], Z.gradient, res). With one noise variance per data
    point (a vector `noise_variance`) the tail of grad is dL_dR row by row (N*P entries), which
    HeteroscedasticGaussian.exact_inference_gradients hands to its variance parameter (heteroscedastic_gaussian.py:33-34)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    kern = StationaryOracle(kind, X.shape[1], variance, lengthscale, ARD)
    res = vardtc_inference(kern, X, Z, noise_variance, Y)
    dv0, dl0 = kern.update_gradients_diag(res["dL_dKdiag"], X)  # sparse_gp.py:110
    dv1, dl1 = kern.update_gradients_full(res["dL_dKnm"], X, Z)  # :112
    dv2, dl2 = kern.update_gradients_full(res["dL_dKmm"], Z, None)  # :114
    dvar = dv0 + dv1 + dv2
    dlen = np.atleast_1d(dl0) + np.atleast_1d(dl1) + np.atleast_1d(dl2)
    Zgrad = kern.gradients_X(res["dL_dKmm"], Z)  # :117
    Zgrad = Zgrad + kern.gradients_X(res["dL_dKnm"].T, Z, X)  # :118
    grad = np.concatenate([[dvar], dlen, np.asarray(res["dL_dthetaL"], dtype=np.float64).reshape(-1)])
    return res["log_marginal"], grad, Zgrad, res


def sparse_raw_predict(kern, Z, woodbury_vector, woodbury_inv, Xnew):
    """posterior.py:238-270 (Posterior._raw_predict with woodbury_inv, 2-D): mu = Kx^T wv, var = Kxx - sum(Kx*(Wi Kx))."""
    Kx = kern.K(Z, Xnew)
    mu = np.dot(Kx.T, woodbury_vector)
    Kxx = kern.Kdiag(Xnew)
    var = (Kxx - np.sum(np.dot(woodbury_inv.T, Kx) * Kx, 0))[:, None]
    return mu, var
