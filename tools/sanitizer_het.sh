#!/bin/bash
# compute-sanitizer memcheck over the heteroscedastic VarDTC evaluation (gpx_sparse_eval_het) at a small size, followed by the
# scalar-noise evaluation on the same context (shared buffers).   gpurun -- bash tools/sanitizer_het.sh
set -u
mkdir -p gpurun_out
cat > /tmp/san_het.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from gpy_b200 import _ffi
rng = np.random.default_rng(0)
N, M, D, P = 700, 130, 3, 2
X = rng.uniform(-3, 3, (N, D)); Y = np.stack([np.sin(X).sum(1) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
Z = X[:M].copy() + 0.01
e = _ffi.Engine(0)
e.sparse_set_data(X, Y)
l, g, dZ, dR = e.sparse_eval_het("matern32", True, 1.2, np.array([1.3, 1.7, 2.1]), Z, rng.uniform(0.01, 0.3, N))
print("het", l, float(np.abs(dR).max()))
print("wv", float(np.abs(e.sparse_get("woodbury_vector")).max()))
l2, g2, dZ2 = e.sparse_eval("matern32", True, 1.2, np.array([1.3, 1.7, 2.1]), Z, 0.05)
print("scalar", l2)
PY
echo "=== memcheck sparse_het (N=700, M=130, P=2: gpx_sparse_eval_het, gpx_sparse_get, gpx_sparse_eval on one context)" > gpurun_out/r2s3_sanitizer_het.txt
timeout 75 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san_het.py 2>&1 | grep -E "ERROR SUMMARY|Error|error|Invalid|^het|^wv|^scalar" | head -14 >> gpurun_out/r2s3_sanitizer_het.txt
cat gpurun_out/r2s3_sanitizer_het.txt
