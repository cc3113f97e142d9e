"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy/SciPy + a little C) of the reference's exact-GP hot path
(`GPRegression` -> `ExactGaussianInference.inference` -> `Kern.K/Kdiag/update_gradients_full`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
anything from this package, and there only as the checker / the timed CPU baseline. The product
(`gpy_b200/`) never imports it and fails loudly when its CUDA library is missing.

Pinning status: the restatement is checked (a) entry by entry against the UNMODIFIED
reference (`/root/reference/GPy`, imported in the build container through the test-only `paramz` stand-in in
`oracle/paramz_shim/`) by `tests/golden/make_golden.py`, whose outputs are committed under `tests/golden/`,
and (b) against the reference's own relational tests (analytic gradient == finite differences, native
helper == NumPy helper, jitchol ladder).
"""
