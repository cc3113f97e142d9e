"""A/B of the serial-chain work of round 2 (session 2): fine DMMA tiles in the diagonal-block chain (option "fine"), the chain
schedule of the tcgen05 sweep (option "chain") and the third-generation base-block kernel (option "base").

    python tools/chain_ab.py [sizes, comma separated] [reps]

Every variant is compared with the round-2 configuration (fine=0, chain=0, base=2) on the same inputs: LML absolute and
gradient relative difference (and against the CPU oracle for N <= 4096), device time per evaluation (median of reps)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpy_b200 import _ffi  # noqa: E402
from oracle import gpy_oracle as o  # noqa: E402

VARIANTS = [
    ("round-2 (fine0 chain0 base2)", dict(fine=0, chain=0, base=2)),
    ("fine1 chain0 base2", dict(fine=1, chain=0, base=2)),
    ("fine1 chain1 base2", dict(fine=1, chain=1, base=2)),
    ("fine1 chain1 base3", dict(fine=1, chain=1, base=3)),
    ("fine1 chain1 base4 (default)", dict(fine=1, chain=1, base=4)),
]
if os.environ.get("CHAIN_AB_SHORT"):
    VARIANTS = [("round-2 (fine0 chain0 base2)", dict(fine=0, chain=0, base=2, base_pdl=0)),
                ("default without base_pdl", dict(fine=1, chain=1, base=4, base_pdl=0)),
                ("default (fine1 chain1 base4 pdl)", dict(fine=1, chain=1, base=4, base_pdl=1))]
NB_SWEEP = {512: (128, 256, 512), 4096: (256, 512, 1024), 8192: (512, 1024), 16384: (512, 1024, 2048)}


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512,1300,4096,8192,16384").split(",")]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    ok = True
    for N in sizes:
        D = 8
        X, Y = o.synthetic(N, D, seed=N)
        var, ls, noise = o.theta_bench(D, True)
        ref = None
        if N <= 4096:
            lml0, g0, _ = o.eval_lml_grad(X, Y, "rbf", True, var, ls, noise)
            ref = (lml0, g0)
        base = None
        for name, opts in VARIANTS:
            e = _ffi.Engine(0)
            for k, v in opts.items():
                e.set_option(k, v)
            e.set_data(X, Y)
            try:
                e.exact_eval("rbf", True, var, ls, noise)
                ts = []
                for _ in range(reps):
                    lml, g, jit = e.exact_eval("rbf", True, var, ls, noise)
                    ts.append(e.stats()["total_ms"])
            except Exception as ex:  # noqa: BLE001
                print("N=%5d %-30s FAILED: %s" % (N, name, ex), flush=True)
                ok = False
                e.close()
                continue
            st = e.stats()
            msg = "N=%5d %-30s %8.3f ms (min %8.3f) sweep %7.3f launches %4d" % (N, name, float(np.median(ts)), min(ts), st["sweep_ms"], st["launches"])
            if base is None:
                base = (lml, g, float(np.median(ts)))
            else:
                dl, dg = abs(lml - base[0]), float(np.max(np.abs(g - base[1]) / np.abs(base[1])))
                msg += " | vs round-2: lml %.1e grad %.1e speed-up %.2fx" % (dl, dg, base[2] / float(np.median(ts)))
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
                L = e.get("L")
                msg += " | L finite %s" % bool(np.all(np.isfinite(L)))
            print(msg, flush=True)
            e.close()
    if len(sys.argv) > 3 and sys.argv[3] == "nb":       # block-size sweep of the default configuration
        for N in sizes:
            if N not in NB_SWEEP:
                continue
            X, Y = o.synthetic(N, 8, seed=N)
            var, ls, noise = o.theta_bench(8, True)
            for nb in NB_SWEEP[N]:
                e = _ffi.Engine(0)
                e.set_option("nb", nb)
                e.set_data(X, Y)
                e.exact_eval("rbf", True, var, ls, noise)
                ts = []
                for _ in range(reps):
                    e.exact_eval("rbf", True, var, ls, noise)
                    ts.append(e.stats()["total_ms"])
                print("N=%5d NB=%4d (default options) %8.3f ms (min %8.3f)" % (N, nb, float(np.median(ts)), min(ts)), flush=True)
                e.close()
    if os.environ.get("CHAIN_AB_PROF"):                  # phase clocks of the base-block kernel (last launch of an N = 4096 evaluation)
        X, Y = o.synthetic(4096, 8, seed=1)
        var, ls, noise = o.theta_bench(8, True)
        e = _ffi.Engine(0)
        e.set_data(X, Y)
        e.exact_eval("rbf", True, var, ls, noise)
        e.set_option("base_pdl", 1)
        e.set_option("base_prof", 1)
        e.exact_eval("rbf", True, var, ls, noise)
        e.set_option("base_prof", 2)
        e.set_option("base_prof", 0)
        e.close()
    print("CHAIN_AB", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
