"""Option sweep of the tcgen05 path at one size: python tools/oz_tune.py N "k=v,k=v" "k=v" ...  (each argument one variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi

N = int(sys.argv[1])
D = 8
rng = np.random.default_rng(0)
X = rng.uniform(-3, 3, (N, D))
Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
th = ("rbf", True, 1.0, np.full(D, np.sqrt(D)), 0.01)
base = None
for var in [""] + sys.argv[2:]:
    e = _ffi.Engine(0)
    for kv in filter(None, var.split(",")):
        k, v = kv.split("=")
        e.set_option(k, int(v))
    e.set_data(X, Y)
    ms = []
    for r in range(6):
        lml, g, _ = e.exact_eval(*th)
        ms.append(e.stats()["total_ms"])
    st = e.stats()
    if base is None:
        base = (lml, g)
    print("N=%d %-28s total %8.3f ms (min %8.3f) sweep %7.2f update-sum %7.2f | vs default: lml %.1e grad %.1e" % (
        N, var or "(default)", float(np.median(ms[2:])), min(ms[2:]), st["sweep_ms"], st["update_ms"], abs(lml - base[0]),
        float(np.max(np.abs(g - base[1]) / np.abs(base[1])))), flush=True)
    e.close()
