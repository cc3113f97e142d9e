"""one (or a few) evaluations at size N for profiler captures: python tools/one_eval.py [N] [reps] [ozaki]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpy_b200 import _ffi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
X = rng.uniform(-3, 3, (N, 8))
Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(8) + 0.1 * rng.standard_normal((N, 1))
e = _ffi.Engine(0)
if len(sys.argv) > 3:
    e.set_option("ozaki", int(sys.argv[3]))
e.set_data(X, Y)
for r in range(reps):
    lml, g, _ = e.exact_eval("rbf", True, 1.0, np.full(8, np.sqrt(8)), 0.01)
    print(lml, e.stats()["total_ms"])
