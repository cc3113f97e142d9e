"""GPU bring-up / regression check of the tcgen05 (Ozaki int8) path against the DMMA path and the CPU oracle.

    python tools/ozaki_check.py [quick|full]

For each size: the same evaluation with option ozaki = 0 (fp64 DMMA) and ozaki = 1 (tcgen05 kind::i8 split), LML and gradient
of both against the oracle (N <= 4096), L / K^-1 / alpha of the Ozaki run against the oracle (N <= 1300), device times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpy_b200 import _ffi  # noqa: E402
from oracle import gpy_oracle as o  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    sizes = [(700, 3, "rbf", True), (1300, 5, "matern52", True), (2048, 8, "rbf", True), (4096, 8, "rbf", True)]
    if mode == "full":
        sizes += [(16384, 8, "rbf", True)]
    ok = True
    for (N, D, kind, ARD) in sizes:
        X, Y = o.synthetic(N, D, seed=N)
        var, ls, noise = o.theta_bench(D, ARD)
        ref = None
        if N <= 4096:
            t0 = time.time()
            lml0, g0, res = o.eval_lml_grad(X, Y, kind, ARD, var, ls, noise)
            ref = (lml0, g0, res, time.time() - t0)
        out = {}
        variants = [("fp64 DMMA", dict(ozaki=0)), ("tcgen05 narrow 8/8", dict(ozaki=1, oz_dig_up=8, oz_wide=0)),
                    ("tcgen05 narrow 8/7", dict(ozaki=1, oz_dig_up=7, oz_wide=0)), ("tcgen05 narrow 8/6", dict(ozaki=1, oz_dig_up=6, oz_wide=0)),
                    ("tcgen05 wide 8/8", dict(ozaki=1, oz_dig_up=8, oz_wide=1)), ("tcgen05 wide 8/7", dict(ozaki=1, oz_dig_up=7, oz_wide=1)),
                    ("tcgen05 wide 8/5", dict(ozaki=1, oz_dig_up=5, oz_wide=1))]
        for name, opts in variants:
            e = _ffi.Engine(0)
            for k, v in opts.items():
                e.set_option(k, v)
            e.set_data(X, Y)
            e.exact_eval(kind, ARD, var, ls, noise)           # warm-up (allocations, tile lists)
            lml, g, jit = e.exact_eval(kind, ARD, var, ls, noise)
            st = e.stats()
            out[name] = (lml, g, st)
            msg = "N=%5d %-8s %-20s total %7.2f ms (sweep %.2f, update %.2f, grad %.2f) launches %d" % (
                N, kind, name, st["total_ms"], st["sweep_ms"], st["update_ms"], st["lauum_ms"], st["launches"])
            if ref is not None:
                el, eg = abs(lml - ref[0]), float(np.max(np.abs(g - ref[1]) / np.abs(ref[1])))
                msg += " | vs oracle: lml abs %.2e grad rel %.2e" % (el, eg)
                if not (el <= 1e-8 and eg <= 1e-6) and "8/5" not in name:
                    ok = False
                    msg += "  <-- OUT OF TOLERANCE"
            if opts["ozaki"] and ref is not None and N <= 1300:
                msg += " | L %.1e Kinv %.1e alpha %.1e" % (rel(e.get("L"), ref[2]["L"]), rel(e.get("Kinv"), ref[2]["Wi"]),
                                                          rel(e.get("alpha"), ref[2]["alpha"]))
            print(msg, flush=True)
            e.close()
        for name, _ in variants[1:]:
            d_l = abs(out["fp64 DMMA"][0] - out[name][0])
            d_g = float(np.max(np.abs(out["fp64 DMMA"][1] - out[name][1]) / np.abs(out["fp64 DMMA"][1])))
            print("   %-20s vs DMMA: lml abs %.2e grad rel %.2e ; speed-up %.2fx" % (
                name, d_l, d_g, out["fp64 DMMA"][2]["total_ms"] / out[name][2]["total_ms"]), flush=True)
            if not (d_l <= 1e-8 and d_g <= 1e-6) and "8/5" not in name:
                ok = False
    print("OZAKI_CHECK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
