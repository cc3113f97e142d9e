"""BASELINE.json configs[2]: GPRegression Matern52 N=16384 D=32 fp64, full optimize() loop (L-BFGS-B on the Logexp-
transformed parameters, as paramz does) through the plugin mirror. Reports evaluations, wall time, evals/s, LML.
With a 5th argument `cpu` the CPU oracle (GPy's operation sequence) is evaluated ONCE at the optimizer's final theta and
its log marginal likelihood and gradient are reported beside the device's ("final LML vs CPU").
    python tools/optimize_probe.py 16384 32 40 iso cpu"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpy_b200

def synthetic(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
    return X, Y

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    ard = (sys.argv[4] == "ard") if len(sys.argv) > 4 else False
    X, Y = synthetic(N, D)
    k = gpy_b200.Matern52(D, variance=1.0, lengthscale=np.sqrt(D), ARD=ard)
    t0 = time.time()
    m = gpy_b200.GPRegression(X, Y, k, noise_var=0.1)
    l0 = m.log_likelihood()
    t1 = time.time()
    res = m.optimize(max_iters=iters)
    t2 = time.time()
    out = {"config": "GPRegression Matern52 N=%d D=%d ARD=%s optimize(lbfgsb, max_iters=%d)" % (N, D, ard, iters),
           "lml_initial": l0, "lml_final": m.log_likelihood(), "n_evals": res["n_evals"], "optimize_wall_s": t2 - t1,
           "evals_per_s": res["n_evals"] / (t2 - t1), "first_eval_incl_alloc_s": t1 - t0,
           "theta_final": {"variance": float(k.variance[0]), "lengthscale": k.lengthscale.values.tolist()[:4],
                           "noise": float(m.likelihood.variance[0])}, "warnflag": int(res.get("warnflag", -1)),
           "device_ms_last_eval": m.inference_method.engine.stats()["total_ms"]}
    if len(sys.argv) > 5 and sys.argv[5] == "cpu":
        from oracle import gpy_oracle as o
        ls = k.lengthscale.values.copy() if ard else float(k.lengthscale[0])
        tc = time.time()
        lml_c, g_c, _ = o.eval_lml_grad(X, Y, "matern52", ard, float(k.variance[0]), ls, float(m.likelihood.variance[0]))
        out["cpu_oracle_at_final_theta"] = {
            "lml": float(lml_c), "lml_abs_diff": abs(float(lml_c) - float(m.log_likelihood())),
            "grad_rel_max": float(np.max(np.abs(np.asarray(m.gradient) - g_c) / np.maximum(np.abs(g_c), 1e-300))),
            "grad_abs_max": float(np.max(np.abs(np.asarray(m.gradient) - g_c))), "grad_cpu": np.asarray(g_c).tolist()[:6],
            "cpu_eval_s": time.time() - tc}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
