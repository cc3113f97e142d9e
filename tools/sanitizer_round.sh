#!/bin/bash
# compute-sanitizer passes over the exact path (DMMA and tcgen05), the composite-kernel path and the sparse path at small
# sizes; summaries land in gpurun_out/ (copy the ERROR SUMMARY lines into profiles/).
#   gpurun -- tools/sanitizer_round.sh
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import gpy_b200
from gpy_b200 import _ffi
case = sys.argv[1]
rng = np.random.default_rng(0)
def data(N, D):
    X = rng.uniform(-3, 3, (N, D)); Y = np.sin(X).sum(1, keepdims=True) + 0.1 * rng.standard_normal((N, 1)); return X, Y
if case == "exact129":
    X, Y = data(129, 3); e = _ffi.Engine(0); e.set_data(X, Y); print(e.exact_eval("matern32", True, 1.1, np.array([1.2, 1.5, 2.0]), 0.05)[0])
elif case == "exact1300_dmma":
    X, Y = data(1300, 4); e = _ffi.Engine(0); e.set_option("ozaki", 0); e.set_data(X, Y); print(e.exact_eval("rbf", True, 1.1, np.full(4, 2.0), 0.05)[0])
elif case == "exact1300_tcgen05":
    X, Y = data(1300, 4); e = _ffi.Engine(0); e.set_option("ozaki", 1); e.set_data(X, Y); print(e.exact_eval("rbf", True, 1.1, np.full(4, 2.0), 0.05)[0])
elif case == "exact1300_tcgen05_wide":
    X, Y = data(1300, 4); e = _ffi.Engine(0); e.set_option("ozaki", 1); e.set_option("oz_wide", 1); e.set_data(X, Y); print(e.exact_eval("rbf", True, 1.1, np.full(4, 2.0), 0.05)[0])
elif case == "composite700":
    X, Y = data(700, 4)
    k = gpy_b200.Add([gpy_b200.Prod([gpy_b200.RBF(2, active_dims=[0, 1]), gpy_b200.Matern32(2, active_dims=[2, 3])]), gpy_b200.White(4, variance=0.05), gpy_b200.Bias(4, variance=0.3)])
    m = gpy_b200.GPRegression(X, Y, k, noise_var=0.05); print(m.log_likelihood()); print(m.predict(X[:3])[0].ravel())
elif case == "sparse":
    X, Y = data(700, 3); Z = X[:130].copy()
    m = gpy_b200.SparseGPRegression(X, Y, kernel=gpy_b200.Matern32(3, lengthscale=1.5), Z=Z); print(m.log_likelihood())
PY
for tool in memcheck racecheck; do
  for c in exact129 exact1300_dmma exact1300_tcgen05 exact1300_tcgen05_wide composite700 sparse; do
    echo "=== $tool $c" >> gpurun_out/r2_sanitizer.txt
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py $c 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|^-?[0-9]" | head -12 >> gpurun_out/r2_sanitizer.txt
  done
done
cat gpurun_out/r2_sanitizer.txt
