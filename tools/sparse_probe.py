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
    split = "split" in sys.argv[4:]
    m = gpy_b200.SparseGPRegression(X, Y, kernel=k, Z=Z, device_algebra=not split)
    t_first = time.time() - t0
    eng = m.inference_method.engine
    times = []
    for rep in range(3):
        m.likelihood.variance.values[...] = 0.05 * (1 + 0.01 * rep)
        t0 = time.time(); m.parameters_changed(); times.append(time.time() - t0)
    # split: time the two device calls alone
    kind, ard, var, ls = k._theta()
    t0 = time.time(); G, pY = eng.sparse_stats(kind, ard, var, ls, Z); t_stats = time.time() - t0
    W2 = np.eye(M) * 1e-3; C = np.zeros((M, 1))
    t0 = time.time(); eng.sparse_grads(W2, C, 20.0); t_grads = time.time() - t0
    out = {"config": "SparseGPRegression RBF ARD N=%d M=%d D=%d" % (N, M, D), "first_eval_incl_alloc_s": t_first,
           "eval_wall_s": float(np.median(times)), "evals_per_s": 1.0 / float(np.median(times)),
           "device_stats_call_s": t_stats, "device_grads_call_s": t_grads,
           "mode": "split (host M x M algebra)" if split else "fused device evaluation (gpx_sparse_eval)",
           "MxM_algebra_s": float(np.median(times)) - t_stats - t_grads,
           "flops_stats": 2.0 * N * M * M / 2, "flops_grads": 2.0 * N * M * M, "lml": m.log_likelihood(),
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
