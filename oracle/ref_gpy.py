"""TEST INFRASTRUCTURE (build container only): run the UNMODIFIED reference from /root/reference.

`load()` makes the reference's own modules importable without executing GPy/__init__.py (which eagerly imports
plotting, examples, every model ...): the packages on the way (GPy, GPy.kern, GPy.kern.src, GPy.core, GPy.util,
GPy.likelihoods, GPy.inference, GPy.inference.latent_function_inference) are registered as empty stub packages whose
__path__ points into /root/reference, so `import GPy.kern.src.rbf` executes the reference's rbf.py, stationary.py,
kern.py, kernel_slice_operations.py, util/linalg.py, util/diag.py ... verbatim. The missing third-party dependency
`paramz` is supplied by the test-only stand-in in oracle/paramz_shim (see its header).

What is NOT the reference's code on this path and is restated here (5 lines in total):
  * GP.parameters_changed's three calls (GPy/core/gp.py:278-280) in `evaluate()` below,
  * the two hook methods of LatentFunctionInference (inference/latent_function_inference/__init__.py:38-49),
  * psi-statistics helpers (kern/src/psi_comp) are replaced by empty classes: they are constructed by RBF.__init__
    (rbf.py:29-32) but never called on the exact-GP path.
Used by tests/golden/make_golden.py and tests/test_reference_crosscheck.py; never by the product or by GPU-side tests
(/root/reference does not exist on the GPU box).
"""
import importlib
import os
import sys
import types

REF = os.environ.get("GPX_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = None


def _stub(name, relpath):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF, relpath)]
    m.__package__ = name
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


def load():
    """-> namespace with RBF, Exponential, Matern32, Matern52, ExactGaussianInference, Gaussian, linalg, version."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(os.path.join(REF, "GPy")):
        raise ImportError("reference tree %s not present" % REF)
    if "GPy" in sys.modules and not getattr(sys.modules["GPy"], "_gpx_stub", False):
        raise ImportError("a real GPy is already imported")
    shim = os.path.join(_HERE, "paramz_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    g = _stub("GPy", "GPy")
    g._gpx_stub = True
    for name, rel in (("GPy.kern", "GPy/kern"), ("GPy.kern.src", "GPy/kern/src"), ("GPy.core", "GPy/core"),
                      ("GPy.util", "GPy/util"), ("GPy.likelihoods", "GPy/likelihoods"),
                      ("GPy.inference", "GPy/inference"),
                      ("GPy.inference.latent_function_inference", "GPy/inference/latent_function_inference")):
        _stub(name, rel)
    # psi-statistics (out of scope) -> empty classes, so that rbf.py / kern.py import and construct
    psi = types.ModuleType("GPy.kern.src.psi_comp")
    for cls in ("PSICOMP_RBF", "PSICOMP_RBF_GPU", "PSICOMP_GH", "PSICOMP_Linear", "PSICOMP_SSRBF"):
        setattr(psi, cls, type(cls, (object,), {"__init__": lambda self, *a, **kw: None}))
    sys.modules["GPy.kern.src.psi_comp"] = psi
    sys.modules["GPy.kern.src"].psi_comp = psi

    class LatentFunctionInference(object):  # inference/latent_function_inference/__init__.py:38-49 (hooks only)
        def on_optimization_start(self):
            pass

        def on_optimization_end(self):
            pass

        def _save_to_input_dict(self):
            return {}

    sys.modules["GPy.inference.latent_function_inference"].LatentFunctionInference = LatentFunctionInference

    for mod in ("GPy.util.config", "GPy.util.diag", "GPy.util.linalg", "GPy.util.misc"):
        importlib.import_module(mod)
    par = importlib.import_module("GPy.core.parameterization")
    sys.modules["GPy.core"].Param = par.Param
    sys.modules["GPy.core"].parameterization = par
    rbf = importlib.import_module("GPy.kern.src.rbf")
    stat = importlib.import_module("GPy.kern.src.stationary")
    egi = importlib.import_module("GPy.inference.latent_function_inference.exact_gaussian_inference")
    gauss = importlib.import_module("GPy.likelihoods.gaussian")
    vdtc = importlib.import_module("GPy.inference.latent_function_inference.var_dtc")
    mixed = importlib.import_module("GPy.likelihoods.mixed_noise")
    ns = types.SimpleNamespace(
        VarDTC=vdtc.VarDTC,
        RBF=rbf.RBF, Exponential=stat.Exponential, Matern32=stat.Matern32, Matern52=stat.Matern52,
        ExactGaussianInference=egi.ExactGaussianInference, Gaussian=gauss.Gaussian,
        HeteroscedasticGaussian=gauss.HeteroscedasticGaussian, MixedNoise=mixed.MixedNoise,
        linalg=sys.modules["GPy.util.linalg"], diag=sys.modules["GPy.util.diag"], stationary=stat,
        __version__=open(os.path.join(REF, "GPy", "__version__.py")).read().split('"')[1])
    _loaded = ns
    return ns


_models = None


def load_models():
    """-> namespace with the reference's own GP, GPRegression and Model classes: GPy/core/model.py, GPy/core/gp.py and
    GPy/models/gp_regression.py executed verbatim on top of `load()`. GPy/core/__init__.py and
    GPy/models/__init__.py are NOT executed (they import every model family); the names gp.py takes from sibling packages
    (`kern.Kern`, `kern.RBF`, `likelihoods.Gaussian/Likelihood/MixedNoise`) are set on the stub packages from the
    reference's own modules. Not the reference's code: `expectation_propagation.EP` is replaced by a class that raises
    (gp.py:100 only reaches it for non-Gaussian likelihoods, outside this path)."""
    global _models
    if _models is not None:
        return _models
    G = load()
    kern_pkg, lik_pkg = sys.modules["GPy.kern"], sys.modules["GPy.likelihoods"]
    kmod = importlib.import_module("GPy.kern.src.kern")
    kern_pkg.Kern = kmod.Kern
    for n in ("RBF", "Exponential", "Matern32", "Matern52"):
        setattr(kern_pkg, n, getattr(G, n))
    lmod = importlib.import_module("GPy.likelihoods.likelihood")
    lik_pkg.Likelihood = lmod.Likelihood
    lik_pkg.Gaussian, lik_pkg.HeteroscedasticGaussian = G.Gaussian, G.HeteroscedasticGaussian
    lik_pkg.MixedNoise = importlib.import_module("GPy.likelihoods.mixed_noise").MixedNoise
    ep = types.ModuleType("GPy.inference.latent_function_inference.expectation_propagation")

    class EP(object):
        def __init__(self, *a, **kw):
            raise NotImplementedError("expectation propagation is outside the exact-GP path (test stand-in)")

    ep.EP = EP
    sys.modules[ep.__name__] = ep
    sys.modules["GPy.inference.latent_function_inference"].expectation_propagation = ep
    sys.modules["GPy.inference.latent_function_inference"].exact_gaussian_inference = sys.modules[
        "GPy.inference.latent_function_inference.exact_gaussian_inference"]
    model = importlib.import_module("GPy.core.model")
    sys.modules["GPy.core"].Model = model.Model
    gp = importlib.import_module("GPy.core.gp")
    sys.modules["GPy.core"].GP = gp.GP
    _stub("GPy.models", "GPy/models")
    reg = importlib.import_module("GPy.models.gp_regression")
    _models = types.SimpleNamespace(GP=gp.GP, GPRegression=reg.GPRegression, Model=model.Model, G=G)
    return _models


_sparse_models = None


def load_sparse_models():
    """-> namespace with the reference's own SparseGP (GPy/core/sparse_gp.py executed verbatim on top of `load_models()`):
    SparseGP.__init__ / parameters_changed / _update_gradients (:41-119) and, through GP, optimize and predict. The model
    class GPy/models/sparse_gp_regression.py derives from SparseGP_MPI, whose module imports the mini-batch / MPI inference
    (var_dtc_parallel.py -> mpi4py optional, VarDTC_minibatch); it loads here too (the one name it takes from the stub
    package, `VarDTC`, is set from the reference's own module); should that chain fail, `SparseGPRegression` is None and the
    tests fall back to SparseGP, which sparse_gp_regression.py:33-59 only wraps."""
    global _sparse_models
    if _sparse_models is not None:
        return _sparse_models
    M = load_models()
    sgp = importlib.import_module("GPy.core.sparse_gp")
    sys.modules["GPy.core"].SparseGP = sgp.SparseGP
    reg = None
    sys.modules["GPy.inference.latent_function_inference"].VarDTC = M.G.VarDTC   # name sparse_gp_regression.py:9 imports
    try:
        importlib.import_module("GPy.core.sparse_gp_mpi")
        reg = importlib.import_module("GPy.models.sparse_gp_regression").SparseGPRegression
    except Exception:  # noqa: BLE001
        reg = None
    _sparse_models = types.SimpleNamespace(SparseGP=sgp.SparseGP, SparseGPRegression=reg, GP=M.GP, G=M.G)
    return _sparse_models


_comb = None


def load_combination():
    """-> namespace with the reference's own Add, Prod, White, Bias (GPy/kern/src/add.py, prod.py, static.py verbatim).
    Add.__init__ imports RBF / Linear / Bias / White from the GPy.kern package (add.py:34): the names are set on the stub
    package from the reference's own modules."""
    global _comb
    if _comb is not None:
        return _comb
    G = load()
    kern_pkg = sys.modules["GPy.kern"]
    static = importlib.import_module("GPy.kern.src.static")
    linear = importlib.import_module("GPy.kern.src.linear")
    kern_pkg.RBF, kern_pkg.Linear, kern_pkg.Bias, kern_pkg.White = G.RBF, linear.Linear, static.Bias, static.White
    add = importlib.import_module("GPy.kern.src.add")
    prod = importlib.import_module("GPy.kern.src.prod")
    _comb = types.SimpleNamespace(Add=add.Add, Prod=prod.Prod, White=static.White, Bias=static.Bias, G=G)
    return _comb


KERNELS = {"rbf": "RBF", "exponential": "Exponential", "matern32": "Matern32", "matern52": "Matern52"}


def evaluate(G, X, Y, kind, ARD, variance, lengthscale, noise, Xnew=None):
    """One GP.parameters_changed() with the reference's own kernel / inference / likelihood objects.
    The three calls are GPy/core/gp.py:278-280."""
    import numpy as np
    D = X.shape[1]
    kern = getattr(G, KERNELS[kind])(D, variance=variance, lengthscale=lengthscale, ARD=ARD)
    lik = G.Gaussian(variance=noise)
    inf = G.ExactGaussianInference()
    posterior, lml, grad_dict = inf.inference(kern, X, lik, Y, None, None)      # gp.py:278
    lik.update_gradients(grad_dict["dL_dthetaL"])                                # gp.py:279
    kern.update_gradients_full(grad_dict["dL_dK"], X)                            # gp.py:280
    grad = np.concatenate([np.atleast_1d(kern.variance.gradient).reshape(-1),
                           np.atleast_1d(kern.lengthscale.gradient).reshape(-1),
                           np.atleast_1d(lik.variance.gradient).reshape(-1)])
    out = dict(lml=float(lml), grad=grad, alpha=np.asarray(posterior.woodbury_vector),
               L=np.asarray(posterior.woodbury_chol), K=np.asarray(kern.K(X)), dL_dK=np.asarray(grad_dict["dL_dK"]))
    if Xnew is not None:
        mu, var = posterior._raw_predict(kern, Xnew, X)                          # posterior.py:273-302
        mu, var = lik.predictive_values(mu, var)                                  # gaussian.py:102-110
        out["mu"], out["var"] = np.asarray(mu), np.asarray(var)
    return out


def evaluate_het(G, X, Y, kind, ARD, variance, lengthscale, noise_vec):
    """GPHeteroscedasticRegression's evaluation (models/gp_heteroscedastic_regression.py:22-37 + core/gp.py:278-280) with
    the reference's own HeteroscedasticGaussian / ExactGaussianInference / kernel objects."""
    import numpy as np
    N, D = X.shape
    Y_metadata = {"output_index": np.arange(N)[:, None]}                          # gp_heteroscedastic_regression.py:26-27
    kern = getattr(G, KERNELS[kind])(D, variance=variance, lengthscale=lengthscale, ARD=ARD)
    lik = G.HeteroscedasticGaussian(Y_metadata)
    lik.variance[:] = np.asarray(noise_vec).reshape(lik.variance.shape)
    inf = G.ExactGaussianInference()
    posterior, lml, grad_dict = inf.inference(kern, X, lik, Y, None, Y_metadata)  # gp.py:278
    lik.update_gradients(grad_dict["dL_dthetaL"])                                  # gp.py:279
    kern.update_gradients_full(grad_dict["dL_dK"], X)                              # gp.py:280
    grad = np.concatenate([np.atleast_1d(kern.variance.gradient).reshape(-1),
                           np.atleast_1d(kern.lengthscale.gradient).reshape(-1),
                           np.asarray(lik.variance.gradient).reshape(-1)])
    return dict(lml=float(lml), grad=grad, alpha=np.asarray(posterior.woodbury_vector))


def evaluate_mixed(G, X, Y, kind, ARD, variance, lengthscale, noise_list, output_index):
    """Exact inference with the reference's own MixedNoise likelihood (likelihoods/mixed_noise.py:14-41): one Gaussian per
    output index; the three calls of GP.parameters_changed (core/gp.py:278-280). The container's gradient assignment
    (`self.gradient = gradients`, mixed_noise.py:35-36) is paramz routing the vector to the leaves in link order; with the
    observer-free stand-in that is done here."""
    import numpy as np
    N, D = X.shape
    Y_metadata = {"output_index": np.asarray(output_index).reshape(-1, 1)}
    kern = getattr(G, KERNELS[kind])(D, variance=variance, lengthscale=lengthscale, ARD=ARD)
    liks = [G.Gaussian(variance=float(v), name="Gaussian_noise_%d" % j) for j, v in enumerate(noise_list)]
    lik = G.MixedNoise(liks)
    inf = G.ExactGaussianInference()
    posterior, lml, grad_dict = inf.inference(kern, X, lik, Y, None, Y_metadata)   # gp.py:278
    kern.update_gradients_full(grad_dict["dL_dK"], X)                               # gp.py:280
    grad = np.concatenate([np.atleast_1d(kern.variance.gradient).reshape(-1),
                           np.atleast_1d(kern.lengthscale.gradient).reshape(-1),
                           np.asarray(grad_dict["dL_dthetaL"]).reshape(-1)])
    return dict(lml=float(lml), grad=grad, alpha=np.asarray(posterior.woodbury_vector),
                variance=np.asarray(lik.gaussian_variance(Y_metadata)), likelihood=lik, Y_metadata=Y_metadata)


def evaluate_sparse(G, X, Y, Z, kind, ARD, variance, lengthscale, noise):
    """One SparseGP.parameters_changed() with the reference's own VarDTC / kernel / likelihood objects; the gradient
    wiring restates GPy/core/sparse_gp.py:108-119 (certain inputs)."""
    import numpy as np
    D = X.shape[1]
    kern = getattr(G, KERNELS[kind])(D, variance=variance, lengthscale=lengthscale, ARD=ARD)
    lik = G.Gaussian(variance=noise)
    inf = G.VarDTC(limit=3)
    post, lml, gd = inf.inference(kern, X, Z, lik, Y)                             # sparse_gp.py:77-80

    def kgrad():
        return np.concatenate([np.atleast_1d(kern.variance.gradient).reshape(-1),
                               np.atleast_1d(kern.lengthscale.gradient).reshape(-1) * np.ones(kern.lengthscale.size)])

    lik.update_gradients(gd["dL_dthetaL"])                                         # :84
    kern.update_gradients_diag(gd["dL_dKdiag"], X)                                 # :110
    kerngrad = kgrad().copy()
    kern.update_gradients_full(gd["dL_dKnm"], X, Z)                                # :112
    kerngrad += kgrad()
    kern.update_gradients_full(gd["dL_dKmm"], Z, None)                             # :114
    kerngrad += kgrad()
    Zgrad = kern.gradients_X(gd["dL_dKmm"], Z)                                     # :117
    Zgrad = Zgrad + kern.gradients_X(gd["dL_dKnm"].T, Z, X)                        # :118
    grad = np.concatenate([kerngrad, np.atleast_1d(lik.variance.gradient).reshape(-1)])
    return dict(lml=float(np.squeeze(lml)), grad=grad, Zgrad=np.asarray(Zgrad), woodbury_vector=np.asarray(post.woodbury_vector),
                woodbury_inv=np.asarray(post.woodbury_inv), dL_dKmm=np.asarray(gd["dL_dKmm"]),
                dL_dKnm=np.asarray(gd["dL_dKnm"]))


def evaluate_sparse_het(G, X, Y, Z, kind, ARD, variance, lengthscale, noise_vec):
    """One SparseGP.parameters_changed() with the reference's own VarDTC and HeteroscedasticGaussian (one noise variance
    per data point: the `het_noise` branches var_dtc.py:127-128,221-227,241-257,267-269); gradient wiring of
    GPy/core/sparse_gp.py:108-119, the likelihood gradient through heteroscedastic_gaussian.py:33-40."""
    import numpy as np
    N, D = X.shape
    Y_metadata = {"output_index": np.arange(N)[:, None]}
    kern = getattr(G, KERNELS[kind])(D, variance=variance, lengthscale=lengthscale, ARD=ARD)
    lik = G.HeteroscedasticGaussian(Y_metadata)
    lik.variance[:] = np.asarray(noise_vec).reshape(lik.variance.shape)
    inf = G.VarDTC(limit=3)
    post, lml, gd = inf.inference(kern, X, Z, lik, Y, Y_metadata)                 # sparse_gp.py:77-80

    def kgrad():
        return np.concatenate([np.atleast_1d(kern.variance.gradient).reshape(-1),
                               np.atleast_1d(kern.lengthscale.gradient).reshape(-1) * np.ones(kern.lengthscale.size)])

    kern.update_gradients_diag(gd["dL_dKdiag"], X)                                 # :110
    kerngrad = kgrad().copy()
    kern.update_gradients_full(gd["dL_dKnm"], X, Z)                                # :112
    kerngrad += kgrad()
    kern.update_gradients_full(gd["dL_dKmm"], Z, None)                             # :114
    kerngrad += kgrad()
    Zgrad = kern.gradients_X(gd["dL_dKmm"], Z)                                     # :117
    Zgrad = Zgrad + kern.gradients_X(gd["dL_dKnm"].T, Z, X)                        # :118
    grad = np.concatenate([kerngrad, np.asarray(gd["dL_dthetaL"], dtype=np.float64).reshape(-1)])
    return dict(lml=float(np.squeeze(lml)), grad=grad, Zgrad=np.asarray(Zgrad), woodbury_vector=np.asarray(post.woodbury_vector),
                woodbury_inv=np.asarray(post.woodbury_inv))
