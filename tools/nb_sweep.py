"""Block-size sweep of the tcgen05 path at the smaller metric sizes + the composite-kernel timing of VERDICT item 6
(sum kernel at N=16384 vs the single kernel): python tools/nb_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi

def synthetic(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
    return X, Y

D = 8
th = ("rbf", True, 1.0, np.full(D, np.sqrt(D)), 0.01)
for N, nbs in ((2048, (256, 512)), (4096, (256, 512, 1024)), (8192, (512, 1024)), (16384, (1024,))):
    X, Y = synthetic(N, D)
    for nb in nbs:
        for oz in (1, 0):
            e = _ffi.Engine(0)
            e.set_option("nb", nb); e.set_option("ozaki", oz)
            e.set_data(X, Y)
            ms = []
            for r in range(4):
                e.exact_eval(*th); ms.append(e.stats()["total_ms"])
            print("N=%5d NB=%4d %-8s %8.3f ms" % (N, nb, "tcgen05" if oz else "DMMA", float(np.median(ms[1:]))), flush=True)
            e.close()
# composite: RBF(ARD) + Matern52(iso) + White at N=16384 vs single RBF
N = 16384
X, Y = synthetic(N, D)
e = _ffi.Engine(0)
e.set_data(X, Y)
ms = []
for r in range(4):
    e.exact_eval(*th); ms.append(e.stats()["total_ms"])
t1 = float(np.median(ms[1:]))
parts = [("rbf", True, 0, list(range(D)), 0.7, np.full(D, np.sqrt(D))), ("matern52", False, 1, list(range(D)), 0.3, 4.0), ("white", False, 2, [], 0.001, None)]
ms = []
for r in range(4):
    lml, g, _ = e.exact_eval_multi(parts, 0.01); ms.append(e.stats()["total_ms"])
t2 = float(np.median(ms[1:]))
st = e.stats()
print("N=16384 single RBF-ARD kernel %.2f ms ; sum kernel RBF-ARD + Matern52 + White (gpx_exact_eval_multi) %.2f ms (+%.1f %%) ; kbuild %.2f ms, gradient passes %.2f ms" % (
    t1, t2, 100 * (t2 / t1 - 1), st["kbuild_ms"], st["lauum_ms"]), flush=True)
