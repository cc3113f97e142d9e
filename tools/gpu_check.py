"""Diagnostic (not a test): run the CUDA path against the oracle on a range of sizes and print every error norm."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi
from oracle import gpy_oracle as o

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

def main():
    sizes = [int(s) for s in sys.argv[1].split(",")] if len(sys.argv) > 1 else [100, 128, 300, 640, 1500]
    eng = _ffi.Engine(0)
    for N in sizes:
        for (kind, ARD, D) in [("rbf", True, 8), ("rbf", False, 2), ("matern52", True, 5), ("matern32", False, 3), ("exponential", True, 4)]:
            X, Y = o.synthetic(N, D, seed=N)
            rng = np.random.default_rng(1)
            ls = np.sqrt(D) * rng.uniform(0.7, 1.4, D) if ARD else np.sqrt(D) * 0.9
            var, noise = 1.3, 0.02
            t0 = time.time()
            l0, g0, res = o.eval_lml_grad(X, Y, kind, ARD, var, ls, noise)
            t1 = time.time()
            try:
                eng.set_data(X, Y)
                if os.environ.get("GPX_TEST_NB"): eng.set_option("nb", int(os.environ["GPX_TEST_NB"]))
                l1, g1, jit = eng.exact_eval(kind, ARD, var, ls, noise)
            except Exception as e:
                print(N, kind, ARD, "EXC", repr(e)); continue
            st = eng.stats()
            line = "N=%5d %-11s ARD=%d  lml cpu %.10f gpu %.10f abs %.2e | grad rel %.2e | gpu %.2f ms cpu %.0f ms" % (
                N, kind, ARD, l0, l1, abs(l0 - l1), np.max(np.abs(g1 - g0) / np.abs(g0)), st["total_ms"], (t1 - t0) * 1e3)
            if N <= 1500:
                L = eng.get("L"); al = eng.get("alpha"); Ki = eng.get("Kinv"); K = eng.get("K"); dl = eng.get("dL_dK"); Li = eng.get("Linv")
                line += " | L %.1e alpha %.1e Kinv %.1e K %.1e dLdK %.1e Linv %.1e" % (
                    rel(L, res["L"]), rel(al, res["alpha"]), rel(Ki, res["Wi"]), rel(K, res["K"]), rel(dl, res["dL_dK"]),
                    rel(Li, np.linalg.inv(res["L"])))
            print(line, flush=True)
    # standalone kernel calls
    X, _ = o.synthetic(300, 6, 3); X2, _ = o.synthetic(170, 6, 4)
    for kind in o.KINDS:
        for ARD in (False, True):
            ls = np.linspace(1.5, 3, 6) if ARD else 2.2
            ko = o.StationaryOracle(kind, 6, 0.7, ls, ARD)
            e1 = rel(_ffi.kern_K(kind, ARD, 0.7, ls, X), ko.K(X)); e2 = rel(_ffi.kern_K(kind, ARD, 0.7, ls, X, X2), ko.K(X, X2))
            dL = np.random.default_rng(0).standard_normal((300, 170))
            dv, dl = _ffi.kern_grad_full(kind, ARD, 0.7, ls, X, dL, X2); v0, l0 = ko.update_gradients_full(dL, X, X2)
            dLs = np.random.default_rng(0).standard_normal((300, 300))
            dv2, dl2 = _ffi.kern_grad_full(kind, ARD, 0.7, ls, X, dLs); v02, l02 = ko.update_gradients_full(dLs, X)
            print("kern %-11s ARD=%d K %.1e K(X,X2) %.1e grad(X,X2) %.1e %.1e grad(X) %.1e %.1e" % (kind, ARD, e1, e2,
                  abs(dv - v0) / abs(v0), np.max(np.abs(dl - l0) / np.abs(l0)), abs(dv2 - v02) / abs(v02), np.max(np.abs(dl2 - l02) / np.abs(l02))))
    # predict
    X, Y = o.synthetic(500, 3, 7); Xn, _ = o.synthetic(37, 3, 8)
    eng.set_data(X, Y); l1, g1, _ = eng.exact_eval("rbf", True, 1.1, [1.0, 1.5, 2.0], 0.05)
    ko = o.StationaryOracle("rbf", 3, 1.1, [1.0, 1.5, 2.0], True); res = o.exact_inference(ko, X, Y, 0.05)
    mu0, v0 = o.raw_predict(ko, X, res["L"], res["alpha"], Xn); mu1, v1 = eng.predict(Xn)
    mu0f, v0f = o.raw_predict(ko, X, res["L"], res["alpha"], Xn, full_cov=True); mu1f, v1f = eng.predict(Xn, full_cov=True)
    print("predict mu %.1e var %.1e fullcov %.1e" % (rel(mu1, mu0), rel(v1, v0), rel(v1f, v0f)))
    # non-PD ladder: duplicate points, zero noise
    Xd = np.repeat(o.synthetic(100, 2, 9)[0], 2, axis=0); Yd = np.sin(Xd[:, :1])
    eng.set_data(Xd, Yd)
    try:
        l, g, jit = eng.exact_eval("rbf", False, 1.0, 2.0, 0.0, jitter=0.0)
        print("ladder: lml %.6f jitter_used %.3e tries %d" % (l, jit, eng.stats()["tries"]))
    except Exception as e:
        print("ladder EXC", repr(e))
    print("launches", eng.total_launches())

if __name__ == "__main__":
    main()
