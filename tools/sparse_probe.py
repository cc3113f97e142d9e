"""BASELINE.json configs[4]: SparseGPRegression RBF N=262144 M=4096 D=16 on one B200 — full VarDTC evaluation (bound +
all gradients) through the mirror. Reports wall time per evaluation and its split (device psi1 statistics / device
gradient reductions / M x M algebra), optionally the CPU oracle on the same inputs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpy_b200
from gpy_b200 import _ffi

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    cpu = "cpu" in sys.argv[4:]
    rng = np.random.default_rng(0)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
    Z = X[rng.permutation(N)[:M]].copy()
    k = gpy_b200.RBF(D, variance=1.0, lengthscale=np.full(D, np.sqrt(D)), ARD=True)
    t0 = time.time()
    m = gpy_b200.SparseGPRegression(X, Y, kernel=k, Z=Z)
    t_first = time.time() - t0
    eng = m.inference_method.engine
    times = []
    for rep in range(3):
        m.likelihood.variance.values[...] = 0.05 * (1 + 0.01 * rep)
        t0 = time.time(); m.parameters_changed(); times.append(time.time() - t0)
    out = {"config": "SparseGPRegression RBF ARD N=%d M=%d D=%d" % (N, M, D), "first_eval_incl_alloc_s": t_first,
           "eval_wall_s": float(np.median(times)), "evals_per_s": 1.0 / float(np.median(times)),
           "mode": "fused device evaluation (gpx_sparse_eval)",
           "flops_NM2": {"tmp = Lm^-1 psi1^T (triangular)": 1.0 * N * M * M, "A = tmp tmp^T (lower)": 1.0 * N * M * M,
                         "dL_dKnm^T = W2 psi1^T": 2.0 * N * M * M}, "lml": m.log_likelihood(),
           "grad_kern_variance": float(k.variance.gradient[0])}
    if cpu:
        from oracle import gpy_oracle as o
        t0 = time.time()
        l0, g0, Zg0, _ = o.sparse_eval(X, Y, Z, "rbf", True, 1.0, np.full(D, np.sqrt(D)), float(m.likelihood.variance[0]))
        out["cpu_oracle_s"] = time.time() - t0
        out["parity_lml_abs"] = abs(l0 - m.log_likelihood())
        g = np.concatenate([k.variance.gradient, k.lengthscale.gradient, m.likelihood.variance.gradient])
        out["parity_grad_rel_max"] = float(np.max(np.abs(g - g0) / np.abs(g0)))
        out["parity_Zgrad_rel"] = float(np.max(np.abs(m.Z.gradient - Zg0)) / np.max(np.abs(Zg0)))
    print(json.dumps(out))

if __name__ == "__main__":
    main()
