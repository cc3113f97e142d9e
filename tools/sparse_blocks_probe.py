"""BASELINE.json configs[4] building blocks: SparseGPRegression RBF N=262144 M=4096 D=16 — Kuf = K(X,Z) build (8.59 GB),
Kuu build and its jitchol (var_dtc.py:93-95,125: Kmm + 1e-8 I -> Lm). Device times only; the VarDTC bound itself is a
'next' row."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    rng = np.random.default_rng(0)
    X = rng.uniform(-3, 3, (N, D))
    Z = X[rng.permutation(N)[:M]].copy()      # sparse_gp_regression.py:41-43
    ls = np.full(D, np.sqrt(D))
    eng = _ffi.Engine(0)
    out = {"config": "RBF ARD N=%d M=%d D=%d" % (N, M, D)}
    for rep in range(2):
        ms, gbs = eng.kern_K_device_only("rbf", True, 1.0, ls, X, Z)
    out["Kuf_build_ms"], out["Kuf_build_GBps"], out["Kuf_bytes"] = ms, gbs, 8.0 * N * M
    ms, gbs = eng.kern_K_device_only("rbf", True, 1.0, ls, Z)
    out["Kuu_build_ms"] = ms
    Kuu = _ffi.kern_K("rbf", True, 1.0, ls, Z)
    Kuu[np.diag_indices(M)] += 1e-8
    t0 = time.time()
    L = _ffi.jitchol(Kuu, engine=eng)
    out["Kuu_jitchol_wall_ms_incl_h2d_d2h"] = (time.time() - t0) * 1e3
    out["Kuu_chol_residual"] = float(np.abs(L.dot(L.T) - Kuu).max())
    print(json.dumps(out))

if __name__ == "__main__":
    main()
