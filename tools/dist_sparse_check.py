"""torchrun --nproc-per-node G tools/dist_sparse_check.py [N M D] : row-sharded sparse VarDTC (gpx_sparse_eval with a
communicator) vs the oracle at small sizes, and timing at the given size (no CPU work at the large size)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import gpy_b200
from gpy_b200 import _ffi
from gpy_b200 import dist as gdist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = _ffi.Engine(local)
    gdist.init_engine_comm(eng)
    from oracle import gpy_oracle as o
    for (kind, ARD, N, M, D, P) in [("rbf", True, 1501, 130, 3, 1), ("matern32", False, 900, 300, 4, 2)]:
        rng = np.random.default_rng(N + M)
        X = rng.uniform(-3, 3, (N, D))
        Y = np.stack([np.sin(X).sum(1) / np.sqrt(D) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
        Z = X[rng.permutation(N)[:M]].copy() + 0.01 * rng.standard_normal((M, D))
        ls = np.sqrt(D) * rng.uniform(0.7, 1.3, D) if ARD else float(np.sqrt(D) * 0.9)
        cls = {"rbf": gpy_b200.RBF, "matern32": gpy_b200.Matern32}[kind]
        k = cls(D, variance=1.3, lengthscale=ls, ARD=ARD)
        rows = gdist.shard_rows(N, rank, world)
        m = gpy_b200.SparseGPRegression(X[rows], Y[rows], kernel=k, Z=Z, engine=eng)
        m.likelihood.variance.values[...] = 0.05
        m.parameters_changed()
        lml0, g0, Zg0, res = o.sparse_eval(X, Y, Z, kind, ARD, 1.3, ls, 0.05)
        g = np.concatenate([k.variance.gradient, k.lengthscale.gradient, m.likelihood.variance.gradient])
        Xn = rng.uniform(-3, 3, (7, D))
        mu, var = m.predict(Xn, include_likelihood=False)
        ko = o.StationaryOracle(kind, D, 1.3, ls, ARD)
        mu0, var0 = o.sparse_raw_predict(ko, Z, res["woodbury_vector"], res["woodbury_inv"], Xn)
        print("SPARSE G=%d rank %d N=%d M=%d %s lml abs %.2e rel %.2e grad rel %.2e Zgrad rel %.2e predict %.2e" % (
            world, rank, N, M, kind, abs(m.log_likelihood() - lml0), abs(m.log_likelihood() - lml0) / abs(lml0),
            np.max(np.abs(g - g0) / np.abs(g0)), np.max(np.abs(m.Z.gradient - Zg0)) / np.max(np.abs(Zg0)),
            max(np.max(np.abs(mu - mu0)), np.max(np.abs(var - var0)))), flush=True)
    if len(sys.argv) > 3:
        N, M, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
        rng = np.random.default_rng(0)
        X = rng.uniform(-3, 3, (N, D))
        Y = np.sin(X).sum(1, keepdims=True) / np.sqrt(D) + 0.1 * rng.standard_normal((N, 1))
        Z = X[rng.permutation(N)[:M]].copy()
        k = gpy_b200.RBF(D, variance=1.0, lengthscale=np.full(D, np.sqrt(D)), ARD=True)
        rows = gdist.shard_rows(N, rank, world)
        m = gpy_b200.SparseGPRegression(X[rows], Y[rows], kernel=k, Z=Z, engine=eng)
        times = []
        for rep in range(4):
            m.likelihood.variance.values[...] = 0.05 * (1 + 0.01 * rep)
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.time(); m.parameters_changed(); torch.cuda.synchronize(); dist.barrier()
            times.append(time.time() - t0)
        if rank == 0:
            print(json.dumps({"config": "row-sharded SparseGPRegression RBF ARD N=%d M=%d D=%d" % (N, M, D), "n_gpus": world,
                              "eval_wall_s": float(np.median(times[1:])), "evals_per_s": 1.0 / float(np.median(times[1:])),
                              "lml": m.log_likelihood(), "single_gpu_lml_reference": -557210.2538022916}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
