"""debug: finite-difference agreement of the sparse model in both modes for one config"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpy_b200

kind, ARD, N, M, D, P = "matern32", True, 600, 600, 5, 1
rng = np.random.default_rng(N + M)
X = rng.uniform(-3, 3, (N, D))
Y = np.stack([np.sin(X).sum(1) / np.sqrt(D) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
Z = X[rng.permutation(N)[:M]].copy() + 0.01 * rng.standard_normal((M, D))
ls = np.sqrt(D) * rng.uniform(0.7, 1.3, D)
for mode in (True, False):
    k = gpy_b200.Matern32(D, variance=1.3, lengthscale=ls.copy(), ARD=ARD)
    m = gpy_b200.SparseGPRegression(X, Y, kernel=k, Z=Z.copy(), device_algebra=mode)
    m.likelihood.variance.values[...] = 0.05
    m.parameters_changed()
    x = m.optimizer_array.copy()
    m.optimizer_array = x
    g = m._grads_transformed()
    f0 = m.objective_function()
    reps = []
    for _ in range(3):
        m.optimizer_array = x
        reps.append(m.objective_function())
    print("mode", mode, "f0 %.12f" % f0, "repeat spread", max(reps) - min(reps))
    nz = Z.size
    rs = np.random.default_rng(0)
    idx = np.concatenate([rs.choice(nz, size=12, replace=False), np.arange(nz, x.size)])
    for step in (1e-6, 1e-5):
        bad = 0
        for i in idx:
            xp, xm = x.copy(), x.copy(); xp[i] += step; xm[i] -= step
            m.optimizer_array = xp; fp = m.objective_function()
            m.optimizer_array = xm; fm = m.objective_function()
            num = (fp - fm) / (2 * step)
            ok = abs(num - g[i]) <= 1e-3 * max(abs(num), 1e-2)
            bad += (not ok)
            if not ok or i >= nz:
                print("  step %g idx %d num %.8e ana %.8e %s" % (step, i, num, g[i], "" if ok else "BAD"))
        print("  step", step, "bad", bad)
