"""Timing diagnosis of the tcgen05 GEMM (measurement only): which role bounds it. Runs the N=16384 evaluation with the
measurement variants of oz_gemm_kernel (oz_dbg: 1 = no MMA issue, 2 = no TMA loads, 4 = no epilogue work; results invalid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi

def synthetic(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
    return X, Y

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X, Y = synthetic(N, 8)
th = ("rbf", True, 1.0, np.full(8, np.sqrt(8)), 0.01)
e = _ffi.Engine(0)
e.set_data(X, Y)
CASES = [(0, 0, 1, 8, "full"), (0, 0, 0, 8, "full, no look-ahead"), (4, 0, 0, 8, "no epilogue"), (2, 0, 0, 8, "no TMA"),
                              (6, 0, 0, 8, "no TMA, no epilogue (MMA issue only)"), (1, 0, 0, 8, "no MMA"), (5, 0, 0, 8, "TMA only"),
                               (0, 132, 1, 8, "132 CTAs"), (0, 296, 1, 8, "296 CTAs (non-persistent-like)"), (0, 0, 1, 7, "inverse part 7 digits"), (0, 0, 1, 6, "inverse part 6 digits"),
                              (0, 0, 1, 4, "inverse part 4 digits")]
CASES += [(0, -t, 1, 7, "7 digits, %d tiles per CTA" % t) for t in (1, 2, 8, 16)]
CASES += [(1000 + d, c, l, g, "WIDE " + lab) for (d, c, l, g, lab) in CASES if lab in (
    "full", "full, no look-ahead", "no epilogue", "no TMA", "no TMA, no epilogue (MMA issue only)", "no MMA", "TMA only",
    "inverse part 7 digits", "inverse part 6 digits", "7 digits, 2 tiles per CTA", "7 digits, 8 tiles per CTA")]
CASES += [(3000 + r, 0, 1, 7, "WIDE 7 digits, panel on the main stream, %d SMs reserved" % r) for r in (0, 2, 4, 8, 16)]
CASES += [(2000, 0, 1, 7, "narrow 7 digits, panel on the main stream, 4 SMs reserved")]
NB_CASES = []
CASES = [c for c in CASES if c[4].startswith("WIDE") and "panel on the main" not in c[4] and "tiles per CTA" not in c[4]]
CASES += [(1008, 0, 0, 8, "WIDE no look-ahead, epilogue polls without back-off"), (1024, 0, 0, 8, "WIDE no look-ahead, epilogue + producer poll without back-off"),
          (4000, 0, 1, 7, "WIDE 7 digits, without the U0 launch"), (1000, 0, 1, 7, "WIDE 7 digits, with the U0 launch (default)"),
          (4000, 0, 1, 8, "WIDE 8 digits, without the U0 launch"), (1000, 0, 1, 8, "WIDE 8 digits, with the U0 launch (default)")]
for (dbg, ctas, la, dig, label) in CASES:
    e.set_option("oz_u0", 0 if dbg == 4000 else 1)
    if dbg == 4000:
        dbg = 1000
    e.set_option("oz_wide", 1 if (1000 <= dbg < 2000 or dbg >= 3000) else 0)
    e.set_option("oz_sched", 1 if dbg >= 2000 else 0)
    e.set_option("oz_reserve", dbg - 3000 if dbg >= 3000 else 4)
    dbg = dbg % 1000 if dbg < 2000 else 0
    e.set_option("oz_dbg", dbg); e.set_option("oz_ctas", max(ctas, 0)); e.set_option("oz_tpc", max(-ctas, 0))
    e.set_option("lookahead", la); e.set_option("oz_dig_up", dig)
    for rep in range(2):
        try:
            e.exact_eval(*th, max_tries=0)
        except Exception as ex:      # the measurement variants produce garbage: a not-PD report is expected
            pass
    st = e.stats()
    print("%-40s total %7.2f ms  sweep %7.2f  update(sum of launches) %7.2f  grad %5.2f  tries %d" % (
        label, st["total_ms"], st["sweep_ms"], st["update_ms"], st["lauum_ms"], st["tries"]), flush=True)

for (nb, tpc) in NB_CASES:      # a fresh context per block size (the panel buffers are sized by it)
    e2 = _ffi.Engine(0)
    e2.set_option("nb", nb); e2.set_option("oz_tpc", tpc); e2.set_option("oz_dig_up", 7)
    e2.set_data(X, Y)
    for rep in range(2):
        e2.exact_eval(*th)
    st = e2.stats()
    print("%-40s total %7.2f ms  sweep %7.2f  update(sum of launches) %7.2f  grad %5.2f" % (
        "WIDE 7 digits, NB = %d, tpc %d" % (nb, tpc), st["total_ms"], st["sweep_ms"], st["update_ms"], st["lauum_ms"]), flush=True)
    e2.close()
