"""Diagnostic (not a test): gpx_sparse_eval_het against the oracle's het_noise VarDTC on several shapes, every error norm
printed (bound, kernel gradients, dL/dZ, per-point noise gradients dL_dR, woodbury vector), plus the device time of one
evaluation at a larger size.   python tools/sparse_het_check.py [big_N big_M]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi
from oracle import gpy_oracle as o


def case(kind, ARD, N, M, D, P, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.stack([np.sin(X).sum(1) / np.sqrt(D) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
    Z = X[rng.permutation(N)[:M]].copy() + 0.01 * rng.standard_normal((M, D))
    ls = np.sqrt(D) * rng.uniform(0.7, 1.3, D) if ARD else float(np.sqrt(D) * 0.9)
    nv = rng.uniform(0.01, 0.3, N)
    return X, Y, Z, ls, nv


def main():
    e = _ffi.Engine(0)
    for (kind, ARD, N, M, D, P) in [("rbf", True, 500, 60, 3, 1), ("matern32", False, 1300, 129, 2, 2),
                                    ("exponential", True, 257, 128, 4, 1), ("matern52", True, 2100, 300, 5, 1),
                                    ("rbf", False, 4000, 512, 2, 1)]:
        X, Y, Z, ls, nv = case(kind, ARD, N, M, D, P, N + M)
        try:
            e.sparse_set_data(X, Y)
            lml, g, dZ, dR = e.sparse_eval_het(kind, ARD, 1.3, ls, Z, nv)
        except Exception as ex:  # noqa: BLE001
            print(kind, N, M, "EXC", repr(ex), flush=True)
            continue
        lml0, g0, Zg0, res = o.sparse_eval(X, Y, Z, kind, ARD, 1.3, ls, nv)
        nk = g.size
        dR0 = g0[nk:].reshape(dR.shape)
        wv = e.sparse_get("woodbury_vector")
        print("%-11s ARD=%d N=%5d M=%4d P=%d | lml %.10f vs %.10f rel %.1e | kern grad rel %.1e | dZ max-rel %.1e | dL_dR "
              "max-rel %.1e (elementwise rel %.1e) | wv %.1e" % (
                  kind, ARD, N, M, P, lml, lml0, abs(lml - lml0) / max(1, abs(lml0)),
                  np.max(np.abs(g - g0[:nk]) / np.abs(g0[:nk])), np.max(np.abs(dZ - Zg0)) / np.abs(Zg0).max(),
                  np.max(np.abs(dR - dR0)) / np.abs(dR0).max(), np.max(np.abs(dR - dR0) / np.maximum(np.abs(dR0), 1e-300)),
                  np.max(np.abs(wv - res["woodbury_vector"])) / np.abs(res["woodbury_vector"]).max()), flush=True)
    if len(sys.argv) > 2:
        N, M = int(sys.argv[1]), int(sys.argv[2])
        X, Y, Z, ls, nv = case("rbf", True, N, M, 16, 1, 1)
        e.sparse_set_data(X, Y)
        for name, f in (("scalar noise", lambda: e.sparse_eval("rbf", True, 1.3, ls, Z, 0.05)),
                        ("per-point noise", lambda: e.sparse_eval_het("rbf", True, 1.3, ls, Z, nv))):
            f()
            t0 = time.perf_counter()
            for _ in range(3):
                f()
            print("N=%d M=%d D=16 VarDTC evaluation, %s: %.1f ms (host clock around the call, 3 calls)" % (
                N, M, name, (time.perf_counter() - t0) / 3 * 1e3), flush=True)
    e.close()


if __name__ == "__main__":
    main()
