"""A/B of the tensor-core panel GEMM (option "oz_panel") and of the chain schedule at the large sizes:

    python tools/panel_ab.py [sizes] [reps]

Every variant against oz_panel = 0 on the same inputs (LML absolute, gradient relative difference; oracle for N <= 4096)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpy_b200 import _ffi  # noqa: E402
from oracle import gpy_oracle as o  # noqa: E402

VARIANTS = [("DMMA panel (oz_panel=0)", dict(oz_panel=0)), ("tensor-core panel (default)", dict(oz_panel=1)),
            ("round-2 schedule (chain=0)", dict(chain=0))]


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1300,4096,8192,16384,32768").split(",")]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ok = True
    for N in sizes:
        X, Y = o.synthetic(N, 8, seed=N)
        var, ls, noise = o.theta_bench(8, True)
        ref = o.eval_lml_grad(X, Y, "rbf", True, var, ls, noise)[:2] if N <= 4096 else None
        base = None
        for name, opts in VARIANTS:
            e = _ffi.Engine(0)
            for k, v in opts.items():
                e.set_option(k, v)
            e.set_data(X, Y)
            e.exact_eval("rbf", True, var, ls, noise)
            ts = []
            for _ in range(reps):
                lml, g, _ = e.exact_eval("rbf", True, var, ls, noise)
                ts.append(e.stats()["total_ms"])
            msg = "N=%5d %-30s %9.3f ms (min %9.3f)" % (N, name, float(np.median(ts)), min(ts))
            if base is None:
                base = (lml, g, float(np.median(ts)))
            else:
                dl, dg = abs(lml - base[0]), float(np.max(np.abs(g - base[1]) / np.abs(base[1])))
                msg += " | vs DMMA panel: lml %.1e grad %.1e speed-up %.3fx" % (dl, dg, base[2] / float(np.median(ts)))
                if not (dl <= 1e-8 and dg <= 1e-6):
                    ok = False
                    msg += " <-- OUT OF TOLERANCE"
            if ref is not None:
                el, eg = abs(lml - ref[0]), float(np.max(np.abs(g - ref[1]) / np.abs(ref[1])))
                msg += " | vs oracle: lml %.1e grad %.1e" % (el, eg)
                if not (el <= 1e-8 and eg <= 1e-6):
                    ok = False
                    msg += " <-- OUT OF TOLERANCE"
            if N <= 1300 and ref is not None:
                res = o.eval_lml_grad(X, Y, "rbf", True, var, ls, noise)[2]
                L = e.get("L")
                msg += " | L rel %.1e" % (float(np.max(np.abs(L - res["L"]))) / float(np.max(np.abs(res["L"]))))
            print(msg, flush=True)
            e.close()
    print("PANEL_AB", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
